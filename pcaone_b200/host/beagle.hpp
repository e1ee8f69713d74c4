// pcaone_b200 host — `FileBeagle`: genotype likelihoods in BEAGLE text format (gz) for the PCAngsd
// algorithm. Mirrors /root/reference/src/FileBeagle.hpp / FileBeagle.cpp:14-68: the constructor sizes
// the data set from the header line and the line count, read_all() parses the likelihoods into the
// 2N x M matrix P and estimates the allele frequencies by EM. Here the parsing is host code (zlib) and
// everything numerical — emMAF_with_GL, the expected genotypes, the EM-PCA loop — runs on the device
// through pcaone_upload_gl / pcaone_gl_em_maf / pcaone_run_em.
#pragma once
#include "data.hpp"

namespace pcaone_host {

class FileBeagle : public Data {
 public:
  explicit FileBeagle(const Param& p);
  ~FileBeagle() override = default;

  void read_all() override;                    // FileBeagle.cpp:14-68
  void check_file_offset_first_var() override {}
  void read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) override;
  void read_block_update(uint64 start_idx, uint64 stop_idx, const Mat2D& U, const Mat1D& svals, const Mat2D& VT,
                         bool standardize) override;
  void attach_stream_source() override;        // out-of-core PCAngsd does not exist in the reference either

  double tolmaf = 1e-6;  // --tol-maf
  std::vector<std::string> samples;  // first name of every sample triple of the header (parse_beagle_samples, Utils.cpp:652-669)
};

}  // namespace pcaone_host
