// pcaone_b200 host — `FileCsv`: a zstd-compressed CSV matrix (`-c/--csv`, one feature per line, one column per sample),
// in core. Mirrors /root/reference/src/FileCsv.hpp / FileCsv.cpp:10-62 (read_all) and :96-146 (parse_csvzstd): the
// `-C/--scale` normalisations (2: log10 of counts per median library size + 1; 1: plain standardisation) and the
// per-feature standardisation of Utils.cpp:36-45 are host arithmetic on the parsed numbers, exactly as in the
// reference; the resulting dense N x M matrix is the operand of the FP64 products on the device
// (pcaone_upload_dense_data). The out-of-core route of the reference (shuffle_csvzstd_to_bin + FileBin) is not built.
#pragma once
#include "data.hpp"

namespace pcaone_host {

class FileCsv : public Data {
 public:
  explicit FileCsv(const Param& p);
  ~FileCsv() override = default;

  void read_all() override;
  void check_file_offset_first_var() override {}
  void read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) override;
  void read_block_update(uint64, uint64, const Mat2D&, const Mat1D&, const Mat2D&, bool) override {}
  void attach_stream_source() override { cao.error("not supporting -m (out-of-core) for CSV input on the B200 path"); }
  const Mat2D& matrix() const { return X; }  // what read_all uploads (tests)

 private:
  Mat2D X;  // nsamples x nsnps after normalisation, logical (permuted) feature order
};

}  // namespace pcaone_host
