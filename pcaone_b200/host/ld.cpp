#include "ld.hpp"

#include <zlib.h>

#include <algorithm>
#include <map>
#include <sstream>
#include <unordered_map>

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

// ---------------------------------------------------------------- LD-based clumping
// Same selection as LD.cpp:323-401 (ld_clump_single_pheno) with the bookkeeping of :270-321 and :105-152
// (valid_assoc_file, map_index_snps, get_target_snp_idx): per chromosome run of the association file the
// variants with P <= p2 are candidates; index variants (P <= p1) are visited by ascending P and absorb every
// remaining candidate within clump_bp whose r2 with them reaches clump_r2. The r2 values are NOT computed pair by
// pair on the host: every SNP that can appear in a pair leads ONE forward window that reaches its farthest partner,
// and all windows go through the banded tile Gram on the device in one call (pcaone_ld_r2_ex).
namespace {

std::vector<std::string> split_tabs(const std::string& line) {
  std::vector<std::string> t;
  size_t b = 0;
  while (true) {
    const size_t e = line.find('\t', b);
    t.push_back(line.substr(b, e == std::string::npos ? e : e - b));
    if (e == std::string::npos) break;
    b = e + 1;
  }
  return t;
}

struct AssocLine {
  std::string chr, text;
  int bp;
  double p;
};

struct AssocRun {                                       // one run of equal CHR values in the file
  std::vector<int> bp, idx;                             // matched variants: position, SNP index of the LD operand
  std::unordered_map<int, int> first_at;                // position -> first slot in bp / idx
  std::unordered_map<int, std::pair<double, int>> cand; // position -> (P, line number) for P <= p2
};

}  // namespace

static void run_ld_clump(Data* data, const Param& params, const pcaone_ld_source& src, const SNPld& snp) {
  std::vector<std::string> files;
  {
    std::istringstream is(params.clump);
    std::string f;
    while (std::getline(is, f, ',')) files.push_back(f);
  }
  std::vector<std::string> names{"CHR", "BP", "P"};
  if (!params.assoc_colnames.empty()) {
    names.clear();
    std::istringstream is(params.assoc_colnames);
    std::string f;
    while (std::getline(is, f, ',')) names.push_back(f);
    if (names.size() != 3) cao.error("--clump-names takes three comma separated column names (chr, pos, pvalue)");
  }
  for (size_t fi = 0; fi < files.size(); ++fi) {
    cao.print(tick.date(), "LD-based clumping for associated file:", files[fi]);
    std::ifstream fin(files[fi]);
    if (!fin.is_open()) cao.error("can not open " + files[fi]);
    std::string head, line;
    std::getline(fin, head);
    int col[3] = {-1, -1, -1};
    {
      const auto h = split_tabs(head);
      for (size_t j = 0; j < h.size(); ++j)
        for (int q = 0; q < 3; ++q)
          if (h[j] == names[q]) col[q] = (int)j;
      for (int q = 0; q < 3; ++q)
        if (col[q] < 0) cao.error("the assoc-like file has no " + names[q] + " column");
    }
    std::vector<AssocLine> lines;
    while (std::getline(fin, line)) {
      if (line.empty()) continue;
      const auto t = split_tabs(line);
      if ((int)t.size() <= std::max(col[0], std::max(col[1], col[2]))) cao.error("short line in " + files[fi]);
      lines.push_back({t[col[0]], line, std::stoi(t[col[1]]), std::stod(t[col[2]])});
    }
    // runs of equal CHR; the candidate of a line goes to the run that was open BEFORE the line was looked at
    // (map_index_snps inserts, then notices the chromosome change), and a run's variant list starts at the last
    // line of the previous run (get_target_snp_idx starts at end_pos[tc - 1]): both kept, outputs are compared
    std::vector<AssocRun> runs;
    std::vector<std::string> run_chr;
    std::vector<int> run_last;
    for (size_t i = 0; i < lines.size(); ++i) {
      if (i == 0 || lines[i].chr != lines[i - 1].chr) {
        if (i) run_last.push_back((int)i - 1);
        run_chr.push_back(lines[i].chr);
      }
    }
    if (lines.empty()) cao.error("no variants in " + files[fi]);
    run_last.push_back((int)lines.size() - 1);
    runs.resize(run_chr.size());
    {
      size_t r = 0;
      for (size_t i = 0; i < lines.size(); ++i) {
        if (lines[i].p <= params.clump_p2) runs[r].cand.insert({lines[i].bp, {lines[i].p, (int)i}});
        if (i && lines[i].chr != lines[i - 1].chr) ++r;
      }
    }
    {  // the association file must list the chromosomes in the order of the LD operand's variants
      int prev = -1;
      for (size_t c = 0; c < snp.chr.size(); ++c) {
        int i = 0;
        while (i < (int)run_chr.size() && run_chr[i] != snp.chr[c]) ++i;
        if (i < prev) cao.error("the association file may be not sorted, hence not matching the bim file");
        prev = i;
      }
    }
    cao.print(tick.date(), "try to match target SNPs to the SNPs in LD matrix");
    for (size_t tc = 0; tc < runs.size(); ++tc) {
      size_t c = 0;
      while (c < snp.chr.size() && snp.chr[c] != run_chr[tc]) ++c;
      if (c == snp.chr.size()) cao.error("chromosome " + run_chr[tc] + " of " + files[fi] + " is not among the LD variants");
      std::unordered_map<int, int> at;
      for (int i = c > 0 ? snp.end_pos[c - 1] : 0; i <= snp.end_pos[c]; ++i) at[snp.pos[i]] = i;
      for (int i = tc > 0 ? run_last[tc - 1] : 0; i <= run_last[tc]; ++i) {
        auto it = at.find(lines[i].bp);
        if (it == at.end()) continue;
        runs[tc].first_at.insert({lines[i].bp, (int)runs[tc].bp.size()});
        runs[tc].bp.push_back(lines[i].bp);
        runs[tc].idx.push_back(it->second);
      }
    }
    // ---- one forward window per SNP that can appear in a pair: the walk of the greedy pass below, run once for every
    // possible index variant without the "already absorbed" test, names every pair (j, k) it can ask for; SNP
    // min(j, k) then leads a window that reaches max(j, k)
    std::map<int, int> reach;  // lead SNP -> last SNP it has to reach
    for (auto& r : runs) {
      const int n = (int)r.bp.size();
      for (auto& kv : r.cand) {
        if (kv.second.first > params.clump_p1) continue;
        auto at = r.first_at.find(kv.first);
        if (at == r.first_at.end()) continue;
        const int j = at->second, p = kv.first;
        auto need = [&](int k) {
          if (!r.cand.count(r.bp[k]) || r.idx[k] == r.idx[j]) return;
          const int a = std::min(r.idx[j], r.idx[k]), b = std::max(r.idx[j], r.idx[k]);
          auto it = reach.find(a);
          if (it == reach.end()) reach[a] = b;
          else it->second = std::max(it->second, b);
        };
        for (int k = j - 1; k >= 0 && (long long)r.bp[k] >= (long long)p - (long long)params.clump_bp; --k) need(k);
        for (int k = j + 1; k < n && (long long)r.bp[k] <= (long long)p + (long long)params.clump_bp; ++k) need(k);
      }
    }
    std::vector<int> ws, we;
    std::vector<uint64> offs{0};
    std::unordered_map<int, int> win_of;
    for (auto& kv : reach) {
      win_of[kv.first] = (int)ws.size();
      ws.push_back(kv.first);
      we.push_back(kv.second - kv.first + 1);
      offs.push_back(offs.back() + (uint64)(kv.second - kv.first));
    }
    std::vector<double> r2(offs.back());
    cao.print(tick.date(), "clumping candidates:", ws.size(), "windows,", offs.back(), "r2 values on the device");
    if (!ws.empty())
      data->check(pcaone_ld_r2_ex(data->ctx, &src, data->nsnps, ws.data(), we.data(), ws.size(), r2.data(), nullptr, 0.0, nullptr));
    auto r2_of = [&](int a, int b) -> double {
      if (a == b) return 1.0;
      if (a > b) std::swap(a, b);
      auto it = win_of.find(a);
      if (it == win_of.end() || b - a >= we[it->second]) cao.error("clumping: pair outside the planned windows");
      return r2[offs[it->second] + (uint64)(b - a - 1)];
    };
    // ---- greedy clumping per run
    std::ofstream ofs(params.fileout + ".p" + std::to_string(fi) + ".clump");
    ofs << head << "\tSP2" << std::endl;
    for (auto& r : runs) {
      auto left = r.cand;  // candidates not absorbed yet
      std::vector<std::pair<double, int>> index;  // (P, position) of the index variants
      for (auto& kv : r.cand)
        if (kv.second.first <= params.clump_p1) index.push_back({kv.second.first, kv.first});
      std::sort(index.begin(), index.end());
      const int n = (int)r.bp.size();
      for (auto& ip : index) {
        const int p = ip.second;
        if (!left.count(p)) continue;  // absorbed by a stronger index variant
        auto at = r.first_at.find(p);
        if (at == r.first_at.end()) continue;
        const int j = at->second;
        std::vector<int> clumped;
        auto visit = [&](int k) {
          const int p2 = r.bp[k];
          if (!left.count(p2)) return;
          if (r2_of(r.idx[j], r.idx[k]) >= params.clump_r2) {
            clumped.push_back(p2);
            left.erase(p2);
          }
        };
        for (int k = j - 1; k >= 0 && (long long)r.bp[k] >= (long long)p - (long long)params.clump_bp; --k) visit(k);
        for (int k = j + 1; k < n && (long long)r.bp[k] <= (long long)p + (long long)params.clump_bp; ++k) visit(k);
        ofs << lines[r.cand.at(p).second].text << "\t";
        if (clumped.empty()) {
          ofs << "NONE";
        } else {
          std::vector<std::pair<double, size_t>> byp;
          for (size_t q = 0; q < clumped.size(); ++q) byp.push_back({r.cand.at(clumped[q]).first, q});
          std::sort(byp.begin(), byp.end());
          for (size_t q = 0; q < byp.size(); ++q) ofs << (q ? "," : "") << clumped[byp[q].second];
        }
        ofs << std::endl;
      }
    }
  }
}

// U of --USV (<prefix>.eigvecs: one row per sample, k tab separated columns) -> column-major N x k
static std::vector<double> read_eigvecs(const std::string& path, uint64 nsamples, uint32_t* ncols) {
  std::ifstream fin(path);
  if (!fin.is_open()) cao.error("can not open " + path);
  std::vector<std::vector<double>> rows;
  std::string line;
  while (std::getline(fin, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream is(line);
    std::vector<double> v;
    double x;
    while (is >> x) v.push_back(x);
    rows.push_back(v);
  }
  if (rows.size() != nsamples) cao.error("the number of rows of " + path + " does not match the samples");
  const size_t k = rows[0].size();
  std::vector<double> U(nsamples * k);
  for (uint64 i = 0; i < nsamples; ++i) {
    if (rows[i].size() != k) cao.error("ragged rows in " + path);
    for (size_t x = 0; x < k; ++x) U[x * nsamples + i] = rows[i][x];
  }
  *ncols = (uint32_t)k;
  return U;
}

void run_ld_stuff(Data* data, const Param& params) {
  cao.print(tick.date(), "run LD stuff");
  // operand of the r2 tiles: the float rows of a `-B` residual file (FileBin::read_all, FileBinary.cpp:21-30) or the
  // centred genotypes of the resident bed
  pcaone_ld_source src{};
  std::vector<double> Uproj;
  if (auto* fb = dynamic_cast<FileBin*>(data)) {
    fb->prepare_ld();
    src.kind = PCAONE_LD_RESID_F32;
    src.data = fb->rows;
    if (params.filebim.empty()) cao.error("-B needs -F/--match-bim (the .mbim written next to the residuals)");
  } else {
    data->prepare();
    src.kind = PCAONE_LD_PACKED;
    if (!params.fileU.empty()) {  // --USV with a bed: LD of (I - U U^T) G (LD.cpp:491-496), applied on the device
      Uproj = read_eigvecs(params.fileU, data->nsamples, &src.ncols);
      src.kind = PCAONE_LD_PACKED_PROJECT;
      src.data = Uproj.data();
    }
  }
  SNPld snp;
  const std::string filebim = params.filebim.empty() ? params.filein + ".bim" : params.filebim;
  get_snp_pos_bim(snp, filebim);
  if ((uint64)snp.pos.size() != data->nsnps) cao.error("the number of SNPs in " + filebim + " does not match the bed");
  if (!params.clump.empty()) {
    run_ld_clump(data, params, src, snp);
    return;
  }
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
