// pcaone_b200 host — option parsing. Table-driven (the reference uses the popl library,
// Cmd.cpp:9-12); flag names, defaults and the post-parse derivations follow Cmd.cpp:141-238.
#include "cmd.hpp"

#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <vector>

#include "../../include/pcaone_b200.h"

namespace pcaone_host {

namespace {

struct Opt {
  const char* sname;  // "" if none
  const char* lname;
  bool takes_value;
  const char* help;
  std::function<void(const std::string&)> set;
  bool seen = false;
};

template <class T>
T parse_num(const std::string& flag, const std::string& s) {
  try {
    size_t pos = 0;
    T v;
    if constexpr (std::is_same_v<T, double>) {
      v = std::stod(s, &pos);
    } else if constexpr (std::is_same_v<T, uint>) {
      if (!s.empty() && s[0] == '-') throw std::invalid_argument("negative");
      v = (uint)std::stoul(s, &pos);
    } else {
      v = (T)std::stol(s, &pos);
    }
    if (pos != s.size()) throw std::invalid_argument("trailing");
    return v;
  } catch (const std::exception&) {
    std::cerr << "Invalid Option Exception: invalid_argument\noption: " << flag << "\nvalue:  " << s << "\n";
    std::exit(EXIT_FAILURE);
  }
}

}  // namespace

int Param::precision_code() const {
  if (precision == "fp64") return PCAONE_PREC_FP64;
  if (precision == "int8x2") return PCAONE_PREC_INT8X2;
  if (precision == "int8x3") return PCAONE_PREC_INT8X3;
  if (precision == "int8x4") return PCAONE_PREC_INT8X4;
  throw std::invalid_argument("--precision must be one of fp64, int8x2, int8x3, int8x4");
}

Param::Param(int argc, char** argv) {
  bool haploid = false, help = false;
  uint svd = 2;
  std::string usvprefix, not_on_path;
  std::vector<Opt> opts;
  auto val = [&](const char* s, const char* l, const char* help_, std::function<void(const std::string&)> f) {
    opts.push_back({s, l, true, help_, std::move(f)});
  };
  auto sw = [&](const char* s, const char* l, const char* help_, bool* flag) {
    opts.push_back({s, l, false, help_, [flag](const std::string&) { *flag = true; }});
  };
  auto off_path = [&](const char* s, const char* l, bool takes) {
    opts.push_back({s, l, takes, nullptr, [&not_on_path, l](const std::string&) { not_on_path = l; }});
  };
#define NUM(T, field) [this](const std::string& v) { field = parse_num<T>(#field, v); }
  sw("h", "help", "print all options", &help);
  val("m", "memory", "RAM usage in GB unit for out-of-core mode. default is in-core mode", NUM(double, memory));
  val("n", "threads", "the number of host threads (accepted; the GPU path does not use them)", NUM(uint, threads));
  val("v", "verbose", "verbosity level for logs. 0: silent; 1: concise; 2: verbose; 3: debug", NUM(uint, verbose));
  val("d", "svd", "SVD method. 1: single-pass RSVD with power iterations (sSVD); 2: window-based RSVD (winSVD, default)",
      [&svd](const std::string& v) { svd = parse_num<uint>("svd", v); });
  val("k", "pc", "top k principal components (PCs) to be calculated", NUM(uint, k));
  val("C", "scale", "-9: standardize genetic data by sqrt(ploidy*f*(1-f)); 0: do nothing", NUM(int, scale));
  val("", "maxp", "maximum number of power iterations for RSVD algorithm.", NUM(uint, maxp));
  sw("S", "no-shuffle", "do not shuffle columns of data for --svd 2 (if not locally correlated).", &noshuffle);
  val("w", "batches", "the number of mini-batches used by --svd 2.", NUM(uint, bands));
  val("", "seed", "seeds for reproducing results.", NUM(int, seed));
  sw("", "emu", "use EMU algorithm for genotype input with missingness.", &emu);
  val("", "M", "the number of features (eg. SNPs) if already known.", NUM(uint, nsnps));
  val("", "N", "the number of samples if already known.", NUM(uint, nsamples));
  val("", "buffer", "memory buffer in GB unit for permuting the data.", NUM(uint, buffer));
  val("", "oversamples", "the number of oversampling columns for RSVD.", NUM(uint, oversamples));
  val("", "rand", "the random matrix type. 0: uniform; 1: guassian.", NUM(uint, rand));
  val("", "maxiter", "maximum number of EM iterations.", NUM(uint, maxiter));
  val("", "tol-rsvd", "tolerance for RSVD algorithm.", NUM(double, tol));
  val("", "tol-em", "tolerance for EMU algorithm.", NUM(double, tolem));
  val("b", "bfile", "prefix of PLINK .bed/.bim/.fam files.", [this](const std::string& v) {
    filein = v;
    file_t = FileType::PLINK;
  });
  sw("", "haploid", "the plink format represents haploid data.", &haploid);
  val("F", "match-bim", "the .mbim file to be matched, where the 7th column is allele frequency.",
      [this](const std::string& v) { filebim = v; });
  val("P", "USV", "prefix of PCAone .eigvecs/.sigvals/.loadings/.mbim.", [&usvprefix](const std::string& v) { usvprefix = v; });
  val("o", "out", "prefix of output files. default [pcaone].", [this](const std::string& v) { fileout = v; });
  sw("V", "printv", "output the right eigenvectors with suffix .loadings.", &printv);
  sw("D", "ld", "output a binary matrix for downstream LD related analysis.", &ld);
  sw("R", "print-r2", "print LD R2 to *.ld.gz file for pairwise SNPs within a window controlled by --ld-bp.", &print_r2);
  val("", "ld-bp", "physical distance threshold in bases for LD window.", NUM(uint, ld_bp));
  val("", "ld-stats", "0: the ancestry adjusted LD; 1: the standard LD.", NUM(int, ld_stats));
  val("", "device", "[GPU] CUDA ordinal of the (first) GPU to use.", NUM(int, device));
  val("", "gpus", "[GPU] shard the SNPs over this many GPUs of the box.", NUM(int, gpus));
  val("", "precision", "[GPU] GEMM arithmetic: fp64 | int8x2 | int8x3 (default) | int8x4.",
      [this](const std::string& v) { precision = v; });
#undef NUM
  // reference flags whose subsystems are outside the GPU hot path (SURVEY §8 "out of scope")
  off_path("p", "pgen", true);
  val("B", "binary", "path of binary file (the .residuals written by --ld); LD options only.", [this](const std::string& v) {
    filein = v;
    file_t = FileType::BINARY;
  });
  val("c", "csv", "path of comma seperated CSV file compressed by zstd.", [this](const std::string& v) {
    filein = v;
    file_t = FileType::CSV;
  });
  val("g", "bgen", "path of BGEN file compressed by gzip/zstd.", [this](const std::string& v) {
    filein = v;
    file_t = FileType::BGEN;
  });
  val("G", "beagle", "path of BEAGLE file compressed by gzip (genotype likelihoods, PCAngsd algorithm).",
      [this](const std::string& v) {
        filein = v;
        file_t = FileType::BEAGLE;
      });
  sw("", "pcangsd", "use PCAngsd algorithm for genotype likelihood input.", &pcangsd);
  val("", "tol-maf", "tolerance for MAF estimation by EM.", [this](const std::string& v) { tolmaf = std::stod(v); });
  off_path("", "hardcall", false);
  off_path("", "maf", true);
  val("", "project", "project the new samples onto the existing PCs (--USV). 1: by multiplying the loadings; 2: by solving g = V x per sample over its called SNPs (takes missing genotypes).",
      [this](const std::string& v) { project = std::stoi(v); });
  off_path("", "project-bootstrap", true);
  off_path("", "project-bootstrap-save", false);
  off_path("", "inbreed", true);
  off_path("", "selection", true);
  val("", "ld-r2", "r2 cutoff for LD-based pruning (usually 0.2).", [this](const std::string& v) { ld_r2 = std::stod(v); });
  val("", "clump", "assoc-like file with target variants and pvalues for clumping.", [this](const std::string& v) { clump = v; });
  val("", "clump-names", "column names in assoc-like file for locating chr, pos and pvalue.",
      [this](const std::string& v) { assoc_colnames = v; });
  val("", "clump-p1", "significance threshold for index SNPs.", [this](const std::string& v) { clump_p1 = std::stod(v); });
  val("", "clump-p2", "secondary significance threshold for clumped SNPs.", [this](const std::string& v) { clump_p2 = std::stod(v); });
  val("", "clump-r2", "r2 cutoff for LD-based clumping.", [this](const std::string& v) { clump_r2 = std::stod(v); });
  val("", "clump-bp", "physical distance threshold in bases for clumping.",
      [this](const std::string& v) { clump_bp = (uint)std::stoul(v); });
  off_path("", "scale-factor", true);
  off_path("", "imaxiter", true);
  off_path("", "itol", true);
  off_path("", "ncv", true);
  off_path("", "tol-maf", true);
  off_path("", "read-U", true);
  off_path("", "read-V", true);
  off_path("", "read-S", true);

  ss << "PCAone-b200 (B200-native randomized-SVD path of PCAone)\nOptions in effect:\n";
  for (int i = 0; i < argc; ++i) ss << argv[i] << ' ';

  auto usage = [&]() {
    std::cout << "Usage: PCAone-b200 -b plink_prefix [-k 10] [-d 1|2] [-m GB] [-o out] ...\n\n";
    for (const auto& o : opts) {
      if (!o.help) continue;
      std::string names = (o.sname[0] ? std::string("-") + o.sname + ", " : std::string("    ")) + "--" + o.lname;
      if (o.takes_value) names += " arg";
      std::cout << "  " << names << std::string(names.size() < 26 ? 26 - names.size() : 1, ' ') << o.help << "\n";
    }
  };

  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], value;
    bool has_inline = false;
    Opt* hit = nullptr;
    if (a.rfind("--", 0) == 0) {
      std::string name = a.substr(2);
      auto eq = name.find('=');
      if (eq != std::string::npos) {
        value = name.substr(eq + 1);
        name = name.substr(0, eq);
        has_inline = true;
      }
      for (auto& o : opts)
        if (name == o.lname) hit = &o;
    } else if (a.size() >= 2 && a[0] == '-') {
      std::string name = a.substr(1, 1);
      for (auto& o : opts)
        if (o.sname[0] && name == o.sname) hit = &o;
      if (hit && a.size() > 2) {
        if (!hit->takes_value) hit = nullptr;  // grouped switches are not supported
        else {
          value = a.substr(2);
          has_inline = true;
        }
      }
    }
    if (!hit) {
      std::cerr << "unknown option: " << a << "\n";
      std::exit(EXIT_FAILURE);
    }
    if (hit->takes_value && !has_inline) {
      if (i + 1 >= argc) {
        std::cerr << "Invalid Option Exception: missing_argument\noption: " << a << "\n";
        std::exit(EXIT_FAILURE);
      }
      value = argv[++i];
    }
    hit->seen = true;
    hit->set(value);
  }
  if (help || argc == 1) {
    usage();
    std::exit(EXIT_SUCCESS);
  }
  try {
    if (!not_on_path.empty())
      throw std::invalid_argument("--" + not_on_path +
                                  " belongs to a PCAone subsystem outside the B200 randomized-SVD path; use the reference "
                                  "PCAone binary for it");
    if (svd == 1)
      svd_t = SvdType::PCAoneAlg1;
    else if (svd == 2)
      svd_t = SvdType::PCAoneAlg2;
    else if (svd == 3)
      svd_t = SvdType::FULL;  // exact PCA: covariance GEMM + eigen-decomposition on the device (Main.cpp:180-217)
    else
      throw std::invalid_argument("--svd 0 (IRAM: the Spectra driver) is outside the B200 path; use --svd 1, 2 or 3");
    if (file_t != FileType::PLINK && file_t != FileType::BEAGLE && file_t != FileType::BINARY && file_t != FileType::BGEN && file_t != FileType::CSV)
      throw std::invalid_argument("please give the PLINK prefix with -b/--bfile, a BGEN file with -g/--bgen, a BEAGLE file with -G/--beagle, or -B residuals for LD");
    genetic = file_t != FileType::CSV;
    if (!usvprefix.empty()) {
      fileU = usvprefix + ".eigvecs";
      fileE = usvprefix + ".eigvals";
      fileS = usvprefix + ".sigvals";
      fileV = usvprefix + ".loadings";
      if (filebim.empty()) filebim = usvprefix + ".mbim";
    }
    if (project != 0) {  // Cmd.cpp:187-194
      if (project < 1 || project > 2)
        throw std::invalid_argument("--project supports 1 or 2 on the B200 path (3, the GL-aware EM projection, is not built)");
      if (fileV.empty() || fileS.empty()) throw std::invalid_argument("please use --USV together with --project");
      if (file_t != FileType::PLINK) throw std::invalid_argument("--project needs PLINK input on the B200 path");
      if (gpus > 1) throw std::invalid_argument("--project runs on one GPU");
      dopca = false, out_of_core = false;
      memory = 0;
    }
    if (print_r2 || ld_r2 > 0 || !clump.empty()) {  // Cmd.cpp:181-184
      dopca = false;
      memory /= 2.0;
    }
    oversamples = oversamples > k ? oversamples : k;  // Cmd.cpp:216
    if (haploid && genetic) ploidy = 1;
    if (memory > 0) out_of_core = true;
    if (dopca && file_t == FileType::BEAGLE) pcangsd = true;  // Cmd.cpp:227
    if (emu || pcangsd)
      missme = true;
    else if (dopca)
      maxiter = 0;
    if (file_t == FileType::BEAGLE) {
      if (out_of_core) throw std::invalid_argument("not supporting -m option (out-of-core) for PCAngsd and BEAGLE input yet!");
      if (emu) throw std::invalid_argument("--emu does not apply to BEAGLE input (PCAngsd is used)");
      if (print_r2 || ld) throw std::invalid_argument("LD options need PLINK input on the B200 path");
      if (gpus > 1) throw std::invalid_argument("--gpus > 1 is not available for BEAGLE input");
      precision = "fp64";  // genotype likelihoods run on the FP64 kernels
    }
    if (file_t == FileType::CSV) {
      if (out_of_core) throw std::invalid_argument("not supporting -m option (out-of-core) for CSV input on the B200 path");
      if (emu || pcangsd) throw std::invalid_argument("--emu / --pcangsd do not apply to CSV input");
      if (print_r2 || ld || ld_r2 > 0 || !clump.empty()) throw std::invalid_argument("LD options need PLINK input on the B200 path");
      if (gpus > 1) throw std::invalid_argument("--gpus > 1 is not available for CSV input");
      if (svd_t == SvdType::FULL) throw std::invalid_argument("--svd 3 needs PLINK input on the B200 path");
      precision = "fp64";  // a dense FP64 matrix runs on the FP64 kernels
    }
    if (file_t == FileType::BGEN) {
      if (out_of_core) throw std::invalid_argument("not supporting -m option (out-of-core) for BGEN input on the B200 path");
      if (emu) throw std::invalid_argument("--emu is not available for BGEN input");
      if (print_r2 || ld || ld_r2 > 0 || !clump.empty()) throw std::invalid_argument("LD options need PLINK input on the B200 path");
      if (gpus > 1) throw std::invalid_argument("--gpus > 1 is not available for BGEN input");
      if (svd_t == SvdType::FULL) throw std::invalid_argument("--svd 3 needs PLINK input on the B200 path");
      precision = "fp64";  // float dosages run on the FP64 kernels
    }
    if (bands < 4 || bands % 2 != 0)
      throw std::invalid_argument("the -w/--batches must be a power of 2 and the minimun is 4.");
    if (svd_t == SvdType::PCAoneAlg2 && !noshuffle) perm = true;
    if (gpus < 1) throw std::invalid_argument("--gpus must be >= 1");
    (void)precision_code();
  } catch (const std::exception& e) {
    std::cerr << "Exception: " << e.what() << "\n";
    std::exit(EXIT_FAILURE);
  }
}

}  // namespace pcaone_host
