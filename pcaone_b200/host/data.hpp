// pcaone_b200 host — `Data` and `FileBed`, the genotype side of the drop-in boundary.
// Mirrors /root/reference/src/Data.hpp:9-59 and FilePlink.hpp:8-47: same class names, the same
// four virtuals (`read_all`, `check_file_offset_first_var`, `read_block_initial`,
// `read_block_update`), the same public planning fields (`start`, `stop`, `nblocks`,
// `blocksize`, `bandFactor`, `F`, `perm`, `readtime`). What changes is where the numbers live:
// the dense N x M `G` of the reference is never materialised — the packed 2-bit rows go to HBM
// (in-core) or are streamed from pinned host buffers (out-of-core) through the C-ABI of
// include/pcaone_b200.h, and `G` only holds a decoded block when a caller asks for one
// through `read_block_initial/update` (the reference's own contract for those calls).
#pragma once
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pcaone_b200.h"
#include "cmd.hpp"
#include "common.hpp"

namespace pcaone_host {

uint64 count_lines(const std::string& path);  // Utils.cpp count_lines

// Data::prepare out-of-core block plan (Data.cpp:42-84). Throws through cao.error.
struct BlockPlan {
  uint64 blocksize = 0;
  uint nblocks = 1, bandFactor = 1;
  std::vector<uint64> start, stop;  // inclusive SNP ranges
};
BlockPlan ooc_block_plan(uint64 N, uint64 M, uint l, double memory_gb, bool winsvd, uint bands);

// PermMat of permute_plink (FilePlink.cpp:303-408): indices[new position] = original SNP.
std::vector<uint32_t> permute_plink_indices(uint64 M, uint64 N, uint bands, uint gb);

// shard of a SNP-sharded job: sSVD = contiguous slice, winSVD = 1/world of every
// window (in the permuted order when there is one)
struct Shard {
  int rank = 0, world = 1;
  std::vector<uint64> snps;         // global (logical) SNP index of every local row, ascending per window
  std::vector<uint64> start, stop;  // local block / window ranges (empty: derive in the library)
};

class Data {
 public:
  explicit Data(const Param& p) : params(p) {}
  virtual ~Data();

  virtual void read_all() = 0;
  virtual void check_file_offset_first_var() = 0;
  virtual void read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) = 0;
  virtual void read_block_update(uint64 start_idx, uint64 stop_idx, const Mat2D& U, const Mat1D& svals,
                                 const Mat2D& VT, bool standardize) = 0;

  // out-of-core: install this object's block reader in the device context
  virtual void attach_stream_source() = 0;

  void prepare();  // Data.cpp:14-85
  // Data.cpp:211-240 (.eigvals .sigvals .eigvecs [.loadings .mbim])
  void write_eigs_files(const Mat1D& E, const Mat1D& S, const Mat2D& U, const Mat2D& V);
  void write_residuals(const Mat1D& S, const Mat2D& U, const Mat2D& VT);  // Data.cpp:242-291
  void save_snps_in_mbim();                                               // Data.cpp:110-170

  // device context of this data set for one GPU; created by prepare()
  pcaone_ctx* ctx = nullptr;
  void check(int rc) const;  // non-zero C status -> cao.error (throws std::runtime_error)

  const Param& params;
  double readtime = 0;
  uint64 nsamples = 0, nsnps = 0;  // nsnps is the whole job's M
  uint64 nsnps_local = 0;          // rows owned by this context (== nsnps on one GPU)
  uint nblocks = 1, bandFactor = 1;
  uint64 blocksize = 0;
  std::vector<uint64> start, stop;
  Mat2D G;                     // a decoded block (read_block_*) — never the whole matrix
  Mat1D F;                     // allele frequencies in the current (possibly permuted) order
  std::vector<uint32_t> perm;  // perm[logical] = original SNP (empty: identity)
  double p_miss = 0.0;
  Shard shard;
  uint svd_code = PCAONE_SVD_WINSVD;

 protected:
  void create_context();
};

class FileBed : public Data {
 public:
  explicit FileBed(const Param& p, int rank = 0, int world = 1);
  ~FileBed() override;

  void read_all() override;                                                    // FilePlink.cpp:26-120
  void check_file_offset_first_var() override;                                 // FilePlink.cpp:17-24
  void read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) override;  // :122-218
  void read_block_update(uint64 start_idx, uint64 stop_idx, const Mat2D& U, const Mat1D& svals, const Mat2D& VT,
                         bool standardize) override;                           // :220-298

  // pcaone_read_block_fn for the out-of-core streamer: gathers the (logically permuted) rows
  static int read_block_cb(void* user, uint64_t s, uint64_t e, uint8_t* dst);
  void attach_stream_source() override;

 private:
  uint64 bed_bytes_per_snp = 0;
  int fd = -1;                 // the .bed, read with pread (no shared file position)
  uint8_t* pinned = nullptr;   // in-core: the packed shard in page-locked memory
};

// `-B file.residuals` (FileBinary.hpp / FileBinary.cpp:21-46): [uint32 M][uint32 N][M x N float32], the LD input the
// reference writes with --ld. The floats stay in the (memory-mapped) file; the device converts, centres and tiles
// them chunk by chunk (pcaone_ld_r2_ex, PCAONE_LD_RESID_F32).
class FileBin : public Data {
 public:
  explicit FileBin(const Param& p);
  ~FileBin() override;
  void read_all() override {}
  void check_file_offset_first_var() override {}
  void read_block_initial(uint64, uint64, bool) override { cao.error("FileBin: the device reads the float rows itself"); }
  void read_block_update(uint64, uint64, const Mat2D&, const Mat1D&, const Mat2D&, bool) override {
    cao.error("FileBin: no EMU on residual input");
  }
  void attach_stream_source() override {}
  void prepare_ld();            // sizes + a device context; no genotype source
  const float* rows = nullptr;  // M x N floats, SNP-major

 private:
  void* map = nullptr;
  size_t map_bytes = 0;
};

}  // namespace pcaone_host
