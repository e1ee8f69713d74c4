// pcaone_b200 host — `FileBgen`: BGEN v1.1 / v1.2 input (`-g/--bgen`), in core. Mirrors /root/reference/src/FileBgen.hpp and
// FileBgen.cpp:15-72 (read_all): per variant the minor-allele dosage of every sample (NaN = missing), allele frequency
// over the non-missing samples, `af > --maf` filter, then centring / mean imputation / scaling — here fused into the
// operand load on the device (pcaone_upload_dosage). The container itself (header, sample block, zlib / zstd variant
// blocks, layout 1 and layout 2 probability packing) is parsed by the code in bgen.cpp, written against the BGEN
// specification; dosage = (2 P(AA) + P(AB)) of the first allele, swapped to 2 - dosage when the first allele is the
// major one by the same sampled-frequency rule as the reference's bgen library (external/bgen/genotypes.cpp:531-555).
#pragma once
#include "data.hpp"

namespace pcaone_host {

class FileBgen : public Data {
 public:
  explicit FileBgen(const Param& p);
  ~FileBgen() override = default;

  void read_all() override;  // FileBgen.cpp:15-72: dosages -> device, allele frequencies on the device
  void check_file_offset_first_var() override {}
  void read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) override;
  void read_block_update(uint64, uint64, const Mat2D&, const Mat1D&, const Mat2D&, bool) override {
    cao.error("FileBgen: --emu is not available for BGEN input");
  }
  void attach_stream_source() override { cao.error("not supporting -m (out-of-core) for BGEN input on the B200 path"); }

  std::vector<float> dosages;  // [nsnps][nsamples] after the --maf filter, logical (permuted) order
  uint64 nvariants_file = 0;
};

}  // namespace pcaone_host
