// Test helper (CPU only): parse a zstd-compressed CSV with the front-end's own reader (host/csv.cpp) and dump the
// normalised matrix it hands to the device: <out>.matrix = [uint64 nsamples][uint64 nsnps][nsamples x nsnps float64,
// column-major]. Same flags as the front-end (`--csv file -C scale -S -o prefix`). Used by tests/test_host_cpu.py against
// the data->G of the reference's FileCsv::read_all on the same file.
#include <cstdio>

#include "../csv.hpp"

namespace pcaone_host {
Logger cao;
Timer tick;
}  // namespace pcaone_host

using namespace pcaone_host;

int main(int argc, char* argv[]) {
  Param params(argc, argv);
  try {
    FileCsv d(params);
    FILE* f = std::fopen((params.fileout + ".matrix").c_str(), "wb");
    if (!f) return 2;
    const uint64_t n = d.nsamples, m = d.nsnps;
    std::fwrite(&n, 8, 1, f);
    std::fwrite(&m, 8, 1, f);
    std::fwrite(d.matrix().data(), sizeof(double), (size_t)n * m, f);
    std::fclose(f);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
