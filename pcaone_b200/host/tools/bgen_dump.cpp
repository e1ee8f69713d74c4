// Test helper (CPU only): parse a .bgen with the front-end's own reader (host/bgen.cpp) and dump the minor-allele
// dosages it hands to the device: <out>.dosages = [uint64 nsnps][uint64 nsamples][nsnps x nsamples float32].
// Same flags as the front-end (`--bgen file -o prefix [--maf x]`). Used by tests/test_host_cpu.py against the dosages
// the reference's vendored bgen reader returns for the same file.
#include <cstdio>

#include "../bgen.hpp"

namespace pcaone_host {
Logger cao;
Timer tick;
}  // namespace pcaone_host

using namespace pcaone_host;

int main(int argc, char* argv[]) {
  Param params(argc, argv);
  try {
    FileBgen d(params);
    FILE* f = std::fopen((params.fileout + ".dosages").c_str(), "wb");
    if (!f) return 2;
    const uint64_t m = d.nsnps, n = d.nsamples;
    std::fwrite(&m, 8, 1, f);
    std::fwrite(&n, 8, 1, f);
    std::fwrite(d.dosages.data(), sizeof(float), d.dosages.size(), f);
    std::fclose(f);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
