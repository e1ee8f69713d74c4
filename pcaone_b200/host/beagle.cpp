#include "beagle.hpp"

#include <sstream>

#include <zlib.h>

#include <cstdlib>
#include <cstring>

namespace pcaone_host {

namespace {

// one text line of a gz stream of any length; false at end of file
bool gz_line(gzFile fp, std::string& line) {
  line.clear();
  char buf[1 << 16];
  while (gzgets(fp, buf, sizeof buf)) {
    line.append(buf);
    if (!line.empty() && line.back() == '\n') {
      line.pop_back();
      return true;
    }
  }
  return !line.empty();
}

uint64 count_fields(const std::string& s) {
  uint64 n = 0;
  bool in = false;
  for (char c : s) {
    const bool sep = c == '\t' || c == ' ' || c == '\r';
    if (!sep && !in) ++n;
    in = !sep;
  }
  return n;
}

}  // namespace

FileBeagle::FileBeagle(const Param& p) : Data(p) {
  cao.print(tick.date(), "start parsing BEAGLE format");
  p_miss = 1.0;  // enable EM-PCA (FileBeagle.hpp:14)
  if (params.nsnps > 0 && params.nsamples > 0) {
    cao.print(tick.date(), "use nsamples and nsnps given by user");
    nsamples = params.nsamples;
    nsnps = params.nsnps;
  } else {
    gzFile fp = gzopen(params.filein.c_str(), "r");
    if (!fp) cao.error("can not open " + params.filein);
    std::string line;
    if (!gz_line(fp, line)) cao.error("empty BEAGLE file " + params.filein);
    const uint64 ncol = count_fields(line);
    if (ncol % 3) cao.error("Number of columns should be a multiple of 3.");
    nsamples = ncol / 3 - 1;
    nsnps = 0;
    while (gz_line(fp, line)) nsnps++;
    gzclose(fp);
  }
  cao.print(tick.date(), "N (# samples):", nsamples, ", M (# SNPs):", nsnps);
  // in-core winSVD shuffles the SNP order (Halko.cpp:183-186): the map is fixed here and applied
  // while the likelihoods are parsed, like FileBed does at read time
  if (p.perm && p.svd_t == SvdType::PCAoneAlg2) {
    perm.resize(nsnps);
    pcaone_shuffle_indices(nsnps, perm.data());
  }
}

void FileBeagle::read_all() {
  // parse_beagle_file (Utils.cpp): per site, skip marker / allele1 / allele2, then per sample the
  // likelihoods of genotype 0 and 1 (the third is implied) -> P(2i, j), P(2i+1, j)
  tick.clock();
  std::vector<double> P((size_t)2 * nsamples * nsnps);
  gzFile fp = gzopen(params.filein.c_str(), "r");
  if (!fp) cao.error("can not open " + params.filein);
  std::string line;
  gz_line(fp, line);  // header: marker allele1 allele2, then three columns per sample
  {
    samples.clear();
    std::istringstream hs(line);
    std::string tok;
    for (uint64 col = 0; hs >> tok; ++col)
      if (col >= 3 && col % 3 == 0) samples.push_back(tok);
  }
  std::vector<uint64> logical_of(nsnps);  // perm[logical] = original
  for (uint64 l = 0; l < nsnps; ++l) logical_of[perm.empty() ? l : perm[l]] = l;
  for (uint64 j = 0; j < nsnps; ++j) {
    if (!gz_line(fp, line)) cao.error("BEAGLE file has fewer sites than expected");
    char* s = line.data();
    char* save = nullptr;
    const char* delims = "\t \r";
    char* tok = strtok_r(s, delims, &save);
    tok = strtok_r(nullptr, delims, &save);
    tok = strtok_r(nullptr, delims, &save);
    double* col = P.data() + (size_t)2 * nsamples * logical_of[j];
    for (uint64 i = 0; i < nsamples; ++i) {
      tok = strtok_r(nullptr, delims, &save);
      if (!tok) cao.error("BEAGLE line with too few columns at site", j + 1);
      col[2 * i] = std::strtod(tok, nullptr);
      tok = strtok_r(nullptr, delims, &save);
      if (!tok) cao.error("BEAGLE line with too few columns at site", j + 1);
      col[2 * i + 1] = std::strtod(tok, nullptr);
      tok = strtok_r(nullptr, delims, &save);
    }
  }
  gzclose(fp);
  readtime += tick.reltime();
  check(pcaone_upload_gl(ctx, P.data(), nsnps, 0));
  cao.print(tick.date(), "begin to estimate allele frequencies");
  int iters = 0;
  check(pcaone_gl_em_maf(ctx, params.maxiter, tolmaf, &iters));
  if (iters < (int)params.maxiter || params.maxiter == 0)
    cao.print(tick.date(), "EM (MAF) converged at iteration:", iters);
  else
    cao.print(tick.date(), "EM (MAF) did not converge");
  F.resize(nsnps);
  check(pcaone_get_F(ctx, F.data()));
}

void FileBeagle::read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) {
  const uint64 B = stop_idx - start_idx + 1;
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  check(pcaone_decode_block(ctx, start_idx, stop_idx, standardize ? 1 : 0, 0, G.data()));
}

void FileBeagle::read_block_update(uint64 start_idx, uint64 stop_idx, const Mat2D& U, const Mat1D& svals, const Mat2D& VT,
                                   bool standardize) {
  const uint64 B = stop_idx - start_idx + 1;
  Mat2D V(VT.cols(), VT.rows());
  for (uint64 i = 0; i < VT.rows(); ++i)
    for (uint64 j = 0; j < VT.cols(); ++j) V(j, i) = VT(i, j);
  check(pcaone_set_usv(ctx, U.data(), svals.data(), V.data()));
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  check(pcaone_decode_block(ctx, start_idx, stop_idx, standardize ? 1 : 0, 1, G.data()));
}

void FileBeagle::attach_stream_source() { cao.error("doesn't support out-of-core PCAngsd algorithm"); }

}  // namespace pcaone_host
