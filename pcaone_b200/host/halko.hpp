// pcaone_b200 host — the operator side of the drop-in boundary: `RsvdOpData`,
// `NormalRsvdOpData`, `FancyRsvdOpData`, `run_pca_with_halko`, with the member names and
// call order of /root/reference/src/Halko.hpp:6-93 and Halko.cpp:271-345. The state the
// reference keeps in Eigen matrices (Omg, H1, H2, G, H, U, S, V) lives in HBM inside the
// pcaone_ctx owned by `data`; the host copies below are filled when a caller reads them.
#pragma once
#include <functional>

#include "data.hpp"

namespace pcaone_host {

using Index = long long;  // Eigen::Index

class RsvdOpData {
 public:
  Data* data;
  bool update = false, standardize = false;
  Mat2D U, Omg, Omg2;  // nsamples x nk ; Omg is the seeded start matrix (initOmg)
  Mat2D V;             // nsnps (local rows of this shard) x nk
  Mat1D S;             // nk
  double diff = 0;     // 1 - mev(U_cur, U_pre) of the last epoch
  int epochs = 0;      // power iterations the last computeUSV ran

  explicit RsvdOpData(Data* data_) : data(data_) {}
  virtual ~RsvdOpData() {}

  virtual Index rows() const = 0;
  virtual Index cols() const = 0;
  virtual Index ranks() const = 0;
  virtual Index oversamples() const = 0;
  inline Index size() const { return ranks() + oversamples(); }

  // One power-iteration pass for epoch pi on the device (Halko.cpp:99-269); G (M x l) and
  // H (N x l) are copied back for the caller like the reference's out-parameters.
  virtual void computeGandH(Mat2D& G, Mat2D& H, int pi) = 0;

  void setFlags(bool is_update, bool is_standardize);  // Halko.hpp:32-35
  void computeUSV(int p, double tol);                  // Halko.cpp:46-97, whole loop on the device
  void initOmg();                                      // Halko.cpp:15-23
  int runEM();                                         // EM driver of Halko.cpp:290-319 on the device
  void fetchUSV();

 protected:
  void gandh_device(Mat2D& G, Mat2D& H, int pi);
};

class NormalRsvdOpData : public RsvdOpData {
  const Index nk, os;

 public:
  NormalRsvdOpData(Data* data_, int k_, int os_ = 10) : RsvdOpData(data_), nk(k_), os(os_) { initOmg(); }
  Index rows() const override { return (Index)data->nsnps_local; }
  Index cols() const override { return (Index)data->nsamples; }
  Index ranks() const override { return nk; }
  Index oversamples() const override { return os; }
  void computeGandH(Mat2D& G, Mat2D& H, int pi = 0) override { gandh_device(G, H, pi); }
};

class FancyRsvdOpData : public RsvdOpData {
  const Index nk, os;

 public:
  FancyRsvdOpData(Data* data_, int k_, int os_ = 10) : RsvdOpData(data_), nk(k_), os(os_) { initOmg(); }
  Index rows() const override { return (Index)data->nsnps_local; }
  Index cols() const override { return (Index)data->nsamples; }
  Index ranks() const override { return nk; }
  Index oversamples() const override { return os; }
  void computeGandH(Mat2D& G, Mat2D& H, int pi = 0) override { gandh_device(G, H, pi); }
};

// Halko.cpp:271-345. `gather_v` (multi-GPU jobs) turns this shard's V rows into the whole
// job's V on the writing rank; single-GPU callers leave it empty.
void run_pca_full(Data* data, const Param& params);  // --svd 3 (Main.cpp:180-217)
void run_projection(Data* data, const Param& params);  // --project 1 | 2 (Projection.cpp:188-246, :305-308)
void run_pcangsd_grm(Data* data, const Param& params, const std::vector<std::string>& samples);  // Halko.cpp:320-334
void run_pca_with_halko(Data* data, const Param& params,
                        const std::function<void(RsvdOpData*)>& before_write = nullptr);

// PLINK2-style .eigvecs2 (Utils.cpp:225-238)
void make_plink2_eigenvec_file(int K, const std::string& fout, const std::string& fin, const std::string& fam);

// one process, `params.gpus` GPUs: one host thread + one context per GPU, SNP-sharded, the
// N x l partial products all-reduced with NCCL over NVLink at every Omega update
void run_pca_sharded(const Param& params);

}  // namespace pcaone_host
