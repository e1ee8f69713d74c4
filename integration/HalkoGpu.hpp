// HalkoGpu.hpp — the ~120 lines a PCAone maintainer adds to run the randomized-SVD hot path on
// libpcaone_b200.so (include/pcaone_b200.h). Everything it derives from is the UNMODIFIED reference:
//
//   GpuFileBed            : Data            (src/Data.hpp:9-59)    — keeps nsamples / nsnps / the block
//                           plan of Data::prepare (src/Data.cpp:14-85); the genotypes stay packed
//                           (2 bits) on the device instead of read_all()'s dense N x M doubles
//   GpuRsvdOpData         : RsvdOpData      (src/Halko.hpp:6-42)   — overrides the pure virtual
//                           computeGandH (src/Halko.hpp:27); RsvdOpData::computeUSV, initOmg and
//                           computeU (src/Halko.cpp:15-97) run as they are on the G / H handed back
//   GpuNormalRsvdOpData / GpuFancyRsvdOpData   the two variants run_pca_with_halko picks between
//                           (src/Halko.cpp:276-282)
//
// It is compiled against /root/reference/src by oracle/Makefile (target refgpu) together with the
// reference's own objects, and tests/test_gpu_dropin.py runs RsvdOpData::computeUSV through it.
#ifndef PCAONE_HALKO_GPU_
#define PCAONE_HALKO_GPU_

#include <fstream>
#include <vector>

#include "Halko.hpp"
#include "Utils.hpp"
#include "pcaone_b200.h"

// PLINK input whose genotypes live on the GPU. Only the sizes are read on the host.
class GpuFileBed : public Data {
 public:
  explicit GpuFileBed(const Param& params_) : Data(params_) {
    nsamples = count_lines(params.filein + ".fam");   // FilePlink.hpp:14-17
    nsnps = count_lines(params.filein + ".bim");
    snpmajor = true;
    if (params.dopca) F = Mat1D::Zero(nsnps);
  }
  void read_all() final {}                                       // the device does it (pcaone_upload_bed)
  void check_file_offset_first_var() final {}
  void read_block_initial(uint64, uint64, bool) final { cao.error("GpuFileBed: blocks are decoded on the device"); }
  void read_block_update(uint64, uint64, const Mat2D&, const Mat1D&, const Mat2D&, bool) final {
    cao.error("GpuFileBed: blocks are decoded on the device");
  }
};

class GpuRsvdOpData : public RsvdOpData {
 protected:
  pcaone_ctx* ctx = nullptr;
  const Index nk, os;
  bool permuted = false;
  void check(int rc) {
    if (rc) cao.error(pcaone_last_error(ctx));   // -> std::runtime_error, Logger.hpp:85-94
  }

 public:
  GpuRsvdOpData(Data* d, int k, int os_, uint32_t svd, int precision) : RsvdOpData(d), nk(k), os(os_) {
    const Param& p = d->params;
    pcaone_config c{};
    c.nsamples = d->nsamples;
    c.nsnps = c.nsnps_total = d->nsnps;
    c.k = k;
    c.oversamples = os_;
    c.svd = svd;
    c.bands = p.bands;
    c.maxp = p.maxp;
    c.tol = p.tol;
    c.ploidy = p.ploidy;
    c.scale = p.scale;
    c.emu = p.emu;
    c.out_of_core = p.out_of_core;
    c.precision = precision;
    c.device = 0;
    c.rank = 0;
    c.world = 1;
    c.maxiter = p.maxiter;
    c.tolem = p.tolem;
    if (pcaone_create(&c, &ctx)) cao.error(pcaone_last_error(nullptr));
    if (p.out_of_core) {   // FileBed::read_block_initial's ifstream.read (FilePlink.cpp:125-136)
      check(pcaone_open_bed(ctx, (p.filein + ".bed").c_str(), 0));
      std::vector<uint64_t> s(d->start.begin(), d->start.end()), e(d->stop.begin(), d->stop.end());
      check(pcaone_set_blocks(ctx, s.data(), e.data(), d->nblocks, d->bandFactor));   // Data.cpp:78-84
    } else {               // FileBed::read_all (FilePlink.cpp:26-120): packed bytes, not doubles
      std::ifstream f(p.filein + ".bed", std::ios::binary);
      const size_t nbytes = (size_t)((d->nsamples + 3) >> 2) * d->nsnps;
      std::vector<uint8_t> bed(nbytes);
      f.seekg(3);
      f.read(reinterpret_cast<char*>(bed.data()), nbytes);
      if ((size_t)f.gcount() != nbytes) cao.error("Cannot read the bed file.");
      check(pcaone_upload_bed(ctx, bed.data(), d->nsnps, 0));
      check(pcaone_allele_freq(ctx));
      check(pcaone_get_F(ctx, d->F.data()));
    }
    initOmg();   // unchanged host RNG (Halko.cpp:15-23)
    check(pcaone_set_omega(ctx, Omg.data()));
  }
  ~GpuRsvdOpData() override { pcaone_destroy(ctx); }
  Index rows() const override { return data->nsnps; }
  Index cols() const override { return data->nsamples; }
  Index ranks() const override { return nk; }
  Index oversamples() const override { return os; }

  // Halko.hpp:27 — one pass; G (M x l) and H (N x l) are Eigen column-major, exactly the ABI layout
  void computeGandH(Mat2D& G, Mat2D& H, int pi) override {
    check(pcaone_set_flags(ctx, update, standardize));
    if (update) check(pcaone_set_usv(ctx, U.data(), S.data(), V.data()));
    if (pi == 0 && data->params.perm && !data->params.out_of_core && !permuted) {   // Halko.cpp:183-186
      std::vector<uint32_t> idx(rows());
      pcaone_shuffle_indices(rows(), idx.data());
      check(pcaone_permute_resident(ctx, idx.data()));
      Eigen::VectorXi pi_(rows());
      for (Index i = 0; i < rows(); ++i) pi_(i) = (int)idx[i];
      data->perm = PermMat(pi_);
      permuted = true;
    }
    if (pi == 0) {         // Halko.cpp:105,160: initOmg() at the start of every computeUSV
      initOmg();
      check(pcaone_set_omega(ctx, Omg.data()));
    }
    check(pcaone_compute_gandh(ctx, pi));
    check(pcaone_get_GH(ctx, G.data(), H.data()));   // the host computeUSV keeps working unchanged
  }

  // optional fast path: the whole epoch loop on the device (what computeUSV does, Halko.cpp:46-97)
  void computeUSVonDevice(int p, double tol) {
    check(pcaone_set_flags(ctx, update, standardize));
    check(pcaone_compute_usv(ctx, p, tol, nullptr, nullptr));
    U.resize(cols(), nk);
    V.resize(rows(), nk);
    S.resize(nk);
    check(pcaone_get_usv(ctx, U.data(), S.data(), V.data()));
  }
};

struct GpuNormalRsvdOpData : GpuRsvdOpData {
  GpuNormalRsvdOpData(Data* d, int k, int os, int precision = PCAONE_PREC_INT8X3)
      : GpuRsvdOpData(d, k, os, PCAONE_SVD_SSVD, precision) {}
};
struct GpuFancyRsvdOpData : GpuRsvdOpData {
  GpuFancyRsvdOpData(Data* d, int k, int os, int precision = PCAONE_PREC_INT8X3)
      : GpuRsvdOpData(d, k, os, PCAONE_SVD_WINSVD, precision) {}
};

#endif  // PCAONE_HALKO_GPU_
