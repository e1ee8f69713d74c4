#include "ld.hpp"

#include <zlib.h>

#include <algorithm>
#include <sstream>

namespace pcaone_host {

static std::vector<std::string> tokens_of(const std::string& line) {
  std::vector<std::string> t;
  std::istringstream is(line);
  std::string w;
  while (is >> w) t.push_back(w);
  return t;
}

void get_snp_pos_bim(SNPld& snp, const std::string& filebim) {
  std::ifstream fin(filebim);
  if (!fin.is_open()) cao.error("can not open " + filebim);
  std::string line, prev;
  int i = 0;
  while (std::getline(fin, line)) {
    if (line.empty() || line[0] == '#') continue;
    auto t = tokens_of(line);
    if (t.size() < 4) cao.error("the input variant file is not valid!\n => " + filebim);
    if (prev.empty()) snp.chr.push_back(t[0]);
    if (!prev.empty() && prev != t[0]) {  // a new chromosome starts
      snp.end_pos.push_back(i - 1);
      snp.chr.push_back(t[0]);
    }
    prev = t[0];
    if (t.size() == 7) snp.af.push_back(std::stod(t[6]));  // LD.cpp:90
    snp.pos.push_back(std::stoi(t[3]));
    ++i;
  }
  snp.end_pos.push_back(i - 1);
}

void divide_pos_by_window(SNPld& snp, int ld_window_bp) {
  // LD.cpp:154-168: every SNP but the last of its chromosome leads a window of the following
  // SNPs within ld_window_bp. Positions ascend inside a chromosome, so the window end only
  // moves forward: a two-pointer sweep gives the reference's (ws, we) in O(M).
  const int n = (int)snp.pos.size();
  int c = 0, j = 0;
  for (int i = 0; i < n; ++i) {
    if (snp.pos[i] == snp.pos[snp.end_pos[c]]) {
      ++c;
      continue;
    }
    const int e = snp.end_pos[c];
    if (j < i) j = i;
    while (j <= e && (long long)snp.pos[j] - snp.pos[i] <= ld_window_bp) ++j;
    snp.ws.push_back(i);
    snp.we.push_back(j - i);
  }
}

std::vector<std::string> read_variant_labels(const std::string& filebim) {
  std::ifstream fin(filebim);
  if (!fin.is_open()) cao.error("can not open " + filebim);
  std::vector<std::string> labels;
  std::string line;
  while (std::getline(fin, line)) {
    if (line.empty() || line[0] == '#') continue;
    auto t = tokens_of(line);
    if (t.size() < 6) cao.error("the input variant file is not valid!\n => " + filebim);
    labels.push_back(t[0] + "\t" + t[3] + "\t" + t[1]);
  }
  return labels;
}

void run_ld_stuff(Data* data, const Param& params) {
  cao.print(tick.date(), "run LD stuff");
  // operand of the r2 tiles: the float rows of a `-B` residual file (FileBin::read_all, FileBinary.cpp:21-30) or the
  // centred genotypes of the resident bed
  pcaone_ld_source src{};
  if (auto* fb = dynamic_cast<FileBin*>(data)) {
    fb->prepare_ld();
    src.kind = PCAONE_LD_RESID_F32;
    src.data = fb->rows;
    if (params.filebim.empty()) cao.error("-B needs -F/--match-bim (the .mbim written next to the residuals)");
  } else {
    data->prepare();
    src.kind = PCAONE_LD_PACKED;
  }
  SNPld snp;
  const std::string filebim = params.filebim.empty() ? params.filein + ".bim" : params.filebim;
  get_snp_pos_bim(snp, filebim);
  if ((uint64)snp.pos.size() != data->nsnps) cao.error("the number of SNPs in " + filebim + " does not match the bed");
  divide_pos_by_window(snp, (int)params.ld_bp);
  if (!params.print_r2) {
    // ld_prune_big (LD.cpp:240-268): greedy pruning on the device, then write_pruned_snp_ids (:170-190)
    if (!(params.ld_r2 > 0)) cao.error("give --print-r2 or --ld-r2 <cutoff>");
    const bool pick_random_one = snp.af.empty();
    cao.print(tick.date(), "LD pruning, choose sites to be kept randomly or with high MAF? 1(random) : 0(high MAF). =>",
              pick_random_one);
    if (!snp.af.empty() && snp.af.size() != data->nsnps) cao.error("the 7th column (allele frequency) must be on every line");
    std::vector<uint8_t> keep(data->nsnps, 1);
    data->check(pcaone_ld_r2_ex(data->ctx, &src, data->nsnps, snp.ws.data(), snp.we.data(), snp.ws.size(), nullptr,
                                snp.af.empty() ? nullptr : snp.af.data(), params.ld_r2, keep.data()));
    uint64 nkeep = 0;
    for (uint8_t k : keep) nkeep += k;
    cao.print(tick.date(), nkeep, " sites will be kept");
    std::ifstream fin(filebim);
    std::ofstream ofs_out(params.fileout + ".ld.prune.out"), ofs_in(params.fileout + ".ld.prune.in");
    std::string line;
    uint64 i = 0;
    while (std::getline(fin, line)) {
      if (line.empty() || line[0] == '#') continue;
      auto t = tokens_of(line);
      std::ofstream& o = keep[i] ? ofs_in : ofs_out;
      o << t[0] << "\t" << t[1] << "\t" << t[2] << "\t" << t[3] << "\t" << t[4] << "\t" << t[5] << std::endl;
      ++i;
    }
    return;
  }
  uint64 npairs = 0;
  for (int w : snp.we) npairs += (uint64)(w - 1);
  cao.print(tick.date(), "LD windows:", snp.ws.size(), ", pairs:", npairs);
  std::vector<double> r2(npairs);
  // centred genotypes of the resident packed shard (G == NULL), one banded Gram on the GPU
  data->check(pcaone_ld_r2_ex(data->ctx, &src, data->nsnps, snp.ws.data(), snp.we.data(), snp.ws.size(), r2.data(), nullptr,
                              0.0, nullptr));
  cao.print(tick.date(), "r2 computed on the device; writing", params.fileout + ".ld.gz");
  auto bims = read_variant_labels(filebim);
  gzFile gz = gzopen((params.fileout + ".ld.gz").c_str(), "wb1");
  if (!gz) cao.error("can not open " + params.fileout + ".ld.gz");
  gzbuffer(gz, 1 << 20);
  std::string out = "CHR_A\tBP_A\tSNP_A\tCHR_B\tBP_B\tSNP_B\tR2\n";
  uint64 p = 0;
  for (size_t w = 0; w < snp.ws.size(); ++w) {
    const int i = snp.ws[w];
    for (int j = 1; j < snp.we[w]; ++j, ++p) {
      out += bims[i];
      out += '\t';
      out += bims[i + j];
      out += '\t';
      out += std::to_string(r2[p]);  // 6 decimals, like LD.cpp:467
      out += '\n';
    }
    if (out.size() > (1u << 22)) {
      if (gzwrite(gz, out.data(), (unsigned)out.size()) != (int)out.size()) cao.error("failed to write data to ld.gz file");
      out.clear();
    }
  }
  if (!out.empty() && gzwrite(gz, out.data(), (unsigned)out.size()) != (int)out.size())
    cao.error("failed to write data to ld.gz file");
  gzclose(gz);
}

}  // namespace pcaone_host
