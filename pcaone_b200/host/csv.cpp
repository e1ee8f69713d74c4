#include "csv.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>

// libzstd is on the box as a shared object without headers: the streaming entry points the reader needs
extern "C" {
struct PcaZstdIn {
  const void* src;
  size_t size, pos;
};
struct PcaZstdOut {
  void* dst;
  size_t size, pos;
};
void* ZSTD_createDStream(void);
size_t ZSTD_freeDStream(void* zds);
size_t ZSTD_decompressStream(void* zds, PcaZstdOut* output, PcaZstdIn* input);
unsigned ZSTD_isError(size_t code);
}

namespace pcaone_host {

static std::string zstd_file_to_string(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) cao.error("can not open " + path);
  std::vector<char> in((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (in.empty()) cao.error("input file is empty.");
  void* ds = ZSTD_createDStream();
  if (!ds) cao.error("ZSTD_createDStream failed");
  std::string out;
  std::vector<char> buf(1 << 22);
  PcaZstdIn input{in.data(), in.size(), 0};
  size_t last = 1;
  while (input.pos < input.size) {
    PcaZstdOut output{buf.data(), buf.size(), 0};
    last = ZSTD_decompressStream(ds, &output, &input);
    if (ZSTD_isError(last)) cao.error("Error: ZSTD decompression failed");
    out.append(buf.data(), output.pos);
  }
  ZSTD_freeDStream(ds);
  if (last != 0) cao.error("EOF before end of ZSTD_decompressStream.");
  return out;
}

FileCsv::FileCsv(const Param& p) : Data(p) {
  cao.print(tick.date(), "start parsing CSV format compressed by ZSTD");
  if (params.scale == 3 || params.scale == 4)
    cao.error("--scale 3 / 4 divide by library sizes the reference only collects for --scale 2; not available here");
  tick.clock();
  const std::string text = zstd_file_to_string(params.filein);
  // ---- lines = features, columns = samples (parse_csvzstd, FileCsv.cpp:96-146); a last line without '\n' is dropped
  std::vector<size_t> line_start;
  for (size_t b = 0, e; (e = text.find('\n', b)) != std::string::npos; b = e + 1) line_start.push_back(b);
  nsnps = line_start.size();
  if (nsnps == 0) cao.error("error when parsing csv file");
  auto line_end = [&](uint64 j) { return text.find('\n', line_start[j]); };
  nsamples = 1 + (uint64)std::count(text.begin() + line_start[0], text.begin() + line_end(0), ',');
  cao.print(tick.date(), "shape of input matrix (features x samples) is", nsnps, " x", nsamples);
  X.resize(nsamples, nsnps);
  std::vector<long long> libsize(nsamples, 0);
  for (uint64 j = 0; j < nsnps; ++j) {
    const char* s = text.data() + line_start[j];
    const char* e = text.data() + line_end(j);
    uint64 i = 0;
    while (s <= e && i < nsamples) {
      char* next = nullptr;
      const float entry = std::strtof(s, &next);  // std::stof of the reference
      if (next == s) cao.error("error when parsing csv file");
      X(i, j) = entry;
      if (params.scale == 2) libsize[i] += std::strtol(s, nullptr, 10);  // std::stoi of the same field
      ++i;
      s = next;
      while (s < e && *s != ',') ++s;
      if (s >= e) break;
      ++s;
    }
    if (i != nsamples) cao.error("the csv file has unaligned columns");
  }
  if (params.scale == 2) {
    // counts per median library size, log10(x + 1) (FileCsv.cpp:41-42); the median of an even number of ints is their
    // integer mean, as get_median<int> computes it
    std::vector<int> ls(libsize.begin(), libsize.end());
    std::sort(ls.begin(), ls.end());
    const size_t n = ls.size();
    const double median = n % 2 == 0 ? (double)((ls[n / 2 - 1] + ls[n / 2]) / 2) : (double)ls[n / 2];
    for (uint64 j = 0; j < nsnps; ++j)
      for (uint64 i = 0; i < nsamples; ++i) X(i, j) = std::log10((double)(float)X(i, j) * median / (double)(int)libsize[i] + 1);
  }
  if (params.scale >= 1) {  // standardize(G), Utils.cpp:36-45: centre, divide by the sample sd when it exceeds the tolerance
    const double sqrt_rdf = std::sqrt((double)nsamples - 1.0);
    for (uint64 j = 0; j < nsnps; ++j) {
      double mean = 0;
      for (uint64 i = 0; i < nsamples; ++i) mean += X(i, j);
      mean /= (double)nsamples;
      double ss = 0;
      for (uint64 i = 0; i < nsamples; ++i) {
        X(i, j) -= mean;
        ss += X(i, j) * X(i, j);
      }
      const double sd = std::sqrt(ss) / sqrt_rdf;
      if (sd > 1e-10)
        for (uint64 i = 0; i < nsamples; ++i) X(i, j) /= sd;
    }
  }
  readtime += tick.reltime();
  // in-core winSVD shuffles the feature order (Halko.cpp:183-186): fixed here, applied before the upload
  if (p.perm && p.svd_t == SvdType::PCAoneAlg2) {
    perm.resize(nsnps);
    pcaone_shuffle_indices(nsnps, perm.data());
    Mat2D Y(nsamples, nsnps);
    for (uint64 l = 0; l < nsnps; ++l)
      std::copy(X.v.begin() + (size_t)perm[l] * nsamples, X.v.begin() + (size_t)(perm[l] + 1) * nsamples,
                Y.v.begin() + (size_t)l * nsamples);
    X.v.swap(Y.v);
  }
}

void FileCsv::read_all() { check(pcaone_upload_dense_data(ctx, X.data())); }

void FileCsv::read_block_initial(uint64 start_idx, uint64 stop_idx, bool) {
  const uint64 B = stop_idx - start_idx + 1;
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  std::copy(X.v.begin() + (size_t)start_idx * nsamples, X.v.begin() + (size_t)(stop_idx + 1) * nsamples, G.v.begin());
}

}  // namespace pcaone_host
