// pcaone_b200 host — RsvdOpData family and run_pca_with_halko on top of the C-ABI
// (include/pcaone_b200.h). Control flow, log lines and the order of setFlags / computeUSV /
// flip_UV / writers follow /root/reference/src/Halko.cpp:271-345; the passes themselves are
// pcaone_compute_usv / pcaone_run_em on the device.
#include "halko.hpp"

#include <cmath>
#include <iomanip>

#include <atomic>
#include <condition_variable>
#include <thread>

#ifdef PCAONE_WITH_NCCL
#include <nccl.h>
#endif

namespace pcaone_host {

void RsvdOpData::setFlags(bool is_update, bool is_standardize) {
  update = is_update;
  standardize = is_standardize;
  data->check(pcaone_set_flags(data->ctx, update ? 1 : 0, standardize ? 1 : 0));
}

void RsvdOpData::initOmg() {
  // the reference's seeded libstdc++ stream (RSVD.hpp:20-59), generated on the host by the
  // library's helper so that Omega is identical by construction, then kept in HBM
  Omg.resize(cols(), size());
  if (pcaone_init_omega((uint64)cols(), (uint32_t)size(), data->params.seed, data->params.rand ? 1 : 0, Omg.data()))
    cao.error("initOmg failed");
  Omg2 = Omg;
  data->check(pcaone_set_omega(data->ctx, Omg.data()));
}

void RsvdOpData::gandh_device(Mat2D& G, Mat2D& H, int pi) {
  data->check(pcaone_compute_gandh(data->ctx, pi));
  if (G.rows() != (uint64)rows() || G.cols() != (uint64)size()) G.resize(rows(), size());
  if (H.rows() != (uint64)cols() || H.cols() != (uint64)size()) H.resize(cols(), size());
  data->check(pcaone_get_GH(data->ctx, G.data(), H.data()));
}

void RsvdOpData::fetchUSV() {
  U.resize(cols(), ranks());
  V.resize(rows(), ranks());
  S.resize(ranks());
  data->check(pcaone_get_usv(data->ctx, U.data(), S.data(), V.data()));
}

void RsvdOpData::computeUSV(int p, double tol) {
  data->check(pcaone_compute_usv(data->ctx, p, tol, &diff, &epochs));
  if (data->params.verbose > 1) cao.print(tick.date(), "running of epoch =", epochs - 1, ", diff =", diff);
  cao.print(tick.date(), "stops at epoch =", epochs);
  fetchUSV();
}

int RsvdOpData::runEM() {
  int iters = 0;
  data->check(pcaone_run_em(data->ctx, &iters));
  fetchUSV();
  return iters;
}

static RsvdOpData* compute_pca(Data* data, const Param& params) {
  RsvdOpData* rsvd;
  if (params.svd_t == SvdType::PCAoneAlg2) {
    if (params.ld) cao.warn("You are recommended to use --svd 1 for outputting the LD residual matrix");
    cao.print(tick.date(), "initialize window-based RSVD (winSVD) with",
              params.out_of_core ? "out-of-core" : "in-core");
    rsvd = new FancyRsvdOpData(data, params.k, params.oversamples);
  } else {
    cao.print(tick.date(), "initialize single-pass RSVD (sSVD) with", params.out_of_core ? "out-of-core" : "in-core");
    rsvd = new NormalRsvdOpData(data, params.k, params.oversamples);
  }
  if (!params.missme) {
    rsvd->setFlags(false, params.genetic ? !params.ld : false);
    rsvd->computeUSV(params.maxp, params.tol);
  } else {
    if (data->p_miss == 0.0 && !params.out_of_core) cao.warn("there is no missing values");
    cao.print(tick.date(), "run EM-PCA. maxiter =", params.maxiter);
    const int iters = rsvd->runEM();  // Halko.cpp:290-319 incl. the final standardised EMU pass
    cao.print(tick.date(), "individual allele frequencies estimated, EM iterations =", iters);
  }
  return rsvd;
}

static void write_pca(Data* data, RsvdOpData* rsvd, const Param& params) {
  Mat2D VT;
  if (params.ld) {
    VT.resize(rsvd->V.cols(), rsvd->V.rows());
    for (uint64 i = 0; i < rsvd->V.rows(); ++i)
      for (uint64 j = 0; j < rsvd->V.cols(); ++j) VT(j, i) = rsvd->V(i, j);
    data->write_residuals(rsvd->S, rsvd->U, VT);
  }
  Mat1D E(rsvd->S.size());
  for (uint64 i = 0; i < E.size(); ++i) E(i) = rsvd->S(i) * rsvd->S(i) / (double)data->nsnps;
  data->write_eigs_files(E, rsvd->S, rsvd->U, rsvd->V);
}

void run_pca_with_halko(Data* data, const Param& params, const std::function<void(RsvdOpData*)>& before_write) {
  RsvdOpData* rsvd = compute_pca(data, params);
  if (before_write) before_write(rsvd);
  if (data->F.size() != data->nsnps && params.out_of_core && data->shard.world == 1) {
    data->F.resize(data->nsnps);  // out-of-core: frequencies were computed during the first pass
    data->check(pcaone_get_F(data->ctx, data->F.data()));
  }
  if (data->shard.rank == 0) write_pca(data, rsvd, params);
  delete rsvd;
  cao.print(tick.date(), "PCAone - Randomized SVD done");
}

// --svd 3, Main.cpp:180-217: exact PCA through the sample covariance. K = G G^T / nsnps on the standardised genotypes
// (pcaone_sample_covariance: FP64 GEMM panels on the device), its eigen-decomposition (pcaone_sym_svd: one-sided
// Jacobi on the device), V = G^T U / sqrt(eval * nsnps) (pcaone_xt_times), flip_UV by the largest |U| entry.
// The reference switches to the SNP x SNP covariance when nsamples > nsnps; the non-zero spectrum and its vectors are
// the same, and only the N x N form is built here.
void run_pca_full(Data* data, const Param& params) {
  cao.print(tick.date(), "running exact PCA with in-core eigendecomposition (PLINK-like).");
  const uint64 N = data->nsamples, M = data->nsnps;
  const uint64 ncomp = std::min<uint64>(params.k, std::min(N, M));
  data->check(pcaone_set_flags(data->ctx, 0, 1));  // standardize_E
  Mat2D K(N, N), Uall(N, N);
  Mat1D Sall(N);
  data->check(pcaone_sample_covariance(data->ctx, K.data()));
  for (auto& x : K.v) x /= (double)M;
  int sweeps = 0;
  data->check(pcaone_sym_svd(data->ctx, K.data(), N, Uall.data(), Sall.data(), &sweeps));
  cao.print(tick.date(), "eigendecomposition of the", N, "x", N, "sample covariance on the device, Jacobi sweeps =", sweeps);
  Mat1D evals(ncomp), svals(ncomp);
  Mat2D U(N, ncomp), V(M, ncomp);
  for (uint64 i = 0; i < ncomp; ++i) {
    evals(i) = std::max(0.0, Sall(i));
    svals(i) = std::sqrt(evals(i) * (double)M);
    for (uint64 r = 0; r < N; ++r) U(r, i) = Uall(r, i);
  }
  data->check(pcaone_xt_times(data->ctx, U.data(), (uint32_t)ncomp, V.data(), nullptr));
  for (uint64 i = 0; i < ncomp; ++i) {
    if (svals(i) > 0)
      for (uint64 r = 0; r < M; ++r) V(r, i) /= svals(i);
    uint64 x = 0;  // flip_UV (Utils.cpp:118-133): the largest |U| entry of every component is positive
    for (uint64 r = 1; r < N; ++r)
      if (std::fabs(U(r, i)) > std::fabs(U(x, i))) x = r;
    if (U(x, i) < 0) {
      for (uint64 r = 0; r < N; ++r) U(r, i) = -U(r, i);
      for (uint64 r = 0; r < M; ++r) V(r, i) = -V(r, i);
    }
  }
  if (data->F.size() != M) {
    data->F.resize(M);
    data->check(pcaone_get_F(data->ctx, data->F.data()));
  }
  data->write_eigs_files(evals, svals, U, V);
}

// --project 1 | 2 (Projection.cpp:188-246): new samples onto the PCs of a reference panel (--USV: .sigvals, .loadings,
// .mbim). Allele frequencies come from the panel's .mbim (Data.cpp:17-31), the genotypes are standardised with them,
//   1: U = G (V S^-1)                                                  one product on the device
//   2: per sample, least squares of g_i = (V S) x over the SNPs it was called at (solve_projection_scores, :159-179):
//      normal equations A_i x = b_i, b = G W on the device, A_i = W^T W - (missing-call indicator) x (products of the
//      columns of W) on the device (pcaone_mask_times), the N small K x K solves (Cholesky) on the host.
// The SNPs of the target must be those of the .mbim, in order (the reference also matches subsets and flipped
// alleles; that bookkeeping is not built).
void run_projection(Data* data, const Param& params) {
  cao.print(tick.date(), "run projection");
  const uint64 N = data->nsamples, M = data->nsnps;
  {  // identical SNP sets, allele frequencies of the panel
    std::ifstream fb(params.filein + ".bim"), fm(params.filebim);
    if (!fb.is_open()) cao.error("can not open " + params.filein + ".bim");
    if (!fm.is_open()) cao.error("can not open " + params.filebim);
    cao.print(tick.date(), "read allele frequency from .mbim file: " + params.filebim);
    Mat1D F(M);
    std::string lb, lm;
    uint64 j = 0;
    while (std::getline(fm, lm)) {
      if (!std::getline(fb, lb) || j >= M) cao.error("the .bim and the .mbim list different SNPs: only identical SNP sets are supported by --project here");
      std::istringstream ib(lb), im(lm);
      std::string tb[6], tm[7];
      for (auto& t : tb) ib >> t;
      for (auto& t : tm) im >> t;
      if (tm[6].empty()) cao.error("the input file is not valid!\n => " + params.filebim);
      if (tb[0] != tm[0] || tb[3] != tm[3] || tb[4] != tm[4] || tb[5] != tm[5])
        cao.error("the .bim and the .mbim list different SNPs: only identical SNP sets are supported by --project here");
      F(j++) = std::stod(tm[6]);
    }
    if (j != M || std::getline(fb, lb)) cao.error("the .bim and the .mbim list different SNPs: only identical SNP sets are supported by --project here");
    data->check(pcaone_set_F(data->ctx, F.data()));
  }
  data->check(pcaone_set_flags(data->ctx, 0, 1));  // standardize_E (Projection.cpp:216)
  cao.print(tick.date(), "start parsing V:", params.fileV, ", S:", params.fileS);
  std::vector<double> S;
  {
    std::ifstream fs(params.fileS);
    if (!fs.is_open()) cao.error("can not open " + params.fileS);
    std::string line;
    while (std::getline(fs, line))
      if (!line.empty() && line[0] != '#') S.push_back(std::stod(line));
  }
  const uint64 K = std::min<uint64>(S.size(), params.k);
  if (K == 0) cao.error("no singular values in " + params.fileS);
  Mat2D V(M, K);
  {
    std::ifstream fv(params.fileV);
    if (!fv.is_open()) cao.error("can not open " + params.fileV);
    std::string line;
    for (uint64 j = 0; j < M; ++j) {
      if (!std::getline(fv, line)) cao.error("the number of rows of " + params.fileV + " does not match the SNPs");
      std::istringstream is(line);
      for (uint64 x = 0; x < K; ++x)
        if (!(is >> V(j, x))) cao.error("too few columns in " + params.fileV);
    }
  }
  Mat2D U(N, K);
  uint64 nmiss = 0;
  data->check(pcaone_missing_count(data->ctx, &nmiss));
  if (params.project == 1) {
    if (nmiss > 0) cao.warn("there are missing genotypes. recommend using --project 2 or 3.");
    for (uint64 x = 0; x < K; ++x)
      for (uint64 j = 0; j < M; ++j) V(j, x) /= S[x];
    data->check(pcaone_x_times(data->ctx, V.data(), (uint32_t)K, U.data()));
  } else {
    if (nmiss == 0) cao.warn("there is no missing genotypes");
    for (uint64 x = 0; x < K; ++x)
      for (uint64 j = 0; j < M; ++j) V(j, x) *= S[x];
    Mat2D b(N, K);
    data->check(pcaone_x_times(data->ctx, V.data(), (uint32_t)K, b.data()));
    // the K (K + 1) / 2 column products of W through the missing-call indicator, in panels of k + oversamples columns
    const uint64 T = K * (K + 1) / 2, panel = params.k + params.oversamples;
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    for (uint32_t a = 0; a < K; ++a)
      for (uint32_t c = a; c < K; ++c) pairs.push_back({a, c});
    Mat2D Mz(N, T);
    for (uint64 t0 = 0; t0 < T; t0 += panel) {
      const uint64 nc = std::min<uint64>(panel, T - t0);
      Mat2D Z(M, nc), out(N, nc);
      for (uint64 t = 0; t < nc; ++t)
        for (uint64 j = 0; j < M; ++j) Z(j, t) = V(j, pairs[t0 + t].first) * V(j, pairs[t0 + t].second);
      data->check(pcaone_mask_times(data->ctx, Z.data(), (uint32_t)nc, out.data()));
      std::copy(out.v.begin(), out.v.end(), Mz.v.begin() + (size_t)t0 * N);
    }
    std::vector<double> A0(T, 0.0);
    for (uint64 t = 0; t < T; ++t)
      for (uint64 j = 0; j < M; ++j) A0[t] += V(j, pairs[t].first) * V(j, pairs[t].second);
    std::vector<double> A(K * K), y(K);
    for (uint64 i = 0; i < N; ++i) {
      for (uint64 t = 0; t < T; ++t) {
        const double v = A0[t] - Mz(i, t);
        A[pairs[t].first * K + pairs[t].second] = A[pairs[t].second * K + pairs[t].first] = v;
      }
      // Cholesky A = L L^T in place (lower), then the two triangular solves
      for (uint64 c = 0; c < K; ++c) {
        double d = A[c * K + c];
        for (uint64 q = 0; q < c; ++q) d -= A[c * K + q] * A[c * K + q];
        if (!(d > 0)) cao.error("--project 2: a sample has too few called SNPs for the requested number of PCs");
        d = std::sqrt(d);
        A[c * K + c] = d;
        for (uint64 r = c + 1; r < K; ++r) {
          double v = A[r * K + c];
          for (uint64 q = 0; q < c; ++q) v -= A[r * K + q] * A[c * K + q];
          A[r * K + c] = v / d;
        }
      }
      for (uint64 r = 0; r < K; ++r) {
        double v = b(i, r);
        for (uint64 q = 0; q < r; ++q) v -= A[r * K + q] * y[q];
        y[r] = v / A[r * K + r];
      }
      for (uint64 r = K; r-- > 0;) {
        double v = y[r];
        for (uint64 q = r + 1; q < K; ++q) v -= A[q * K + r] * U(i, q);
        U(i, r) = v / A[r * K + r];
      }
    }
  }
  std::ofstream outu(params.fileout + ".eigvecs");
  outu << std::setprecision(6);
  for (uint64 i = 0; i < N; ++i) {
    for (uint64 x = 0; x < K; ++x) outu << (x ? "\t" : "") << U(i, x);
    outu << "\n";
  }
}

// The GRM step of PCAngsd (Halko.cpp:320-334) after the EM loop: covariance of the re-standardised expected
// genotypes with its Dc diagonal (pcaone_gl_grm) -> <out>.cov; its SVD (pcaone_sym_svd, on the device) -> all N
// eigenvectors in <out>.eigvecs2 with the sample names of the BEAGLE header (write_eigvecs2_beagle, Utils.cpp:671-683).
void run_pcangsd_grm(Data* data, const Param& params, const std::vector<std::string>& samples) {
  cao.print(tick.date(), "estimate GRM for pcangsd");
  const uint64 N = data->nsamples;
  Mat2D C(N, N), U2(N, N);
  Mat1D S2(N);
  data->check(pcaone_gl_grm(data->ctx, C.data(), nullptr));
  {
    std::ofstream fcov(params.fileout + ".cov");
    if (!fcov.is_open()) cao.error("can not open " + params.fileout + ".cov");
    for (uint64 i = 0; i < N; ++i) {
      for (uint64 j = 0; j < N; ++j) fcov << (j ? " " : "") << C(i, j);
      fcov << "\n";
    }
  }
  int sweeps = 0;
  data->check(pcaone_sym_svd(data->ctx, C.data(), N, U2.data(), S2.data(), &sweeps));
  std::ofstream feig2(params.fileout + ".eigvecs2");
  if (!feig2.is_open()) cao.error("can not open " + params.fileout + ".eigvecs2");
  feig2 << "#FID\tIID";
  for (uint64 i = 0; i < N; ++i) feig2 << "\tPC" << i + 1;
  feig2 << "\n";
  for (uint64 i = 0; i < N; ++i) {
    const std::string name = i < samples.size() ? samples[i] : "Ind" + std::to_string(i);
    feig2 << name << "\t" << name << "\t";
    for (uint64 j = 0; j < N; ++j) feig2 << (j ? " " : "") << U2(i, j);
    feig2 << "\n";
  }
}

void make_plink2_eigenvec_file(int K, const std::string& fout, const std::string& fin, const std::string& fam) {
  std::ifstream ifam(fam), ifin(fin);
  std::ofstream ofs(fout);
  ofs << "#FID\tIID";
  for (int i = 0; i < K; i++) ofs << "\tPC" << i + 1;
  ofs << "\n";
  std::string line1, line2;
  while (std::getline(ifam, line1)) {
    std::istringstream is(line1);
    std::string fid, iid;
    is >> fid >> iid;
    std::getline(ifin, line2);
    ofs << fid << "\t" << iid << "\t" << line2 << std::endl;
  }
}

// ---------------------------------------------------------------------------- multi-GPU
#ifdef PCAONE_WITH_NCCL
namespace {
// breakable barrier: a worker that throws releases the others instead of leaving them waiting
class Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count, waiting = 0, gen = 0;
  bool broken = false;

 public:
  explicit Barrier(int n) : count(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    if (broken) throw std::runtime_error("another GPU worker failed");
    const int g = gen;
    if (++waiting == count) {
      waiting = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return g != gen || broken; });
      if (broken) throw std::runtime_error("another GPU worker failed");
    }
  }
  void brk() {
    std::lock_guard<std::mutex> lk(m);
    broken = true;
    cv.notify_all();
  }
};
}  // namespace
#endif

void run_pca_sharded(const Param& params) {
#ifndef PCAONE_WITH_NCCL
  (void)params;
  cao.error("this binary was built without NCCL; --gpus > 1 is unavailable");
#else
  const int g = params.gpus;
  if (pcaone_device_count() < params.device + g) cao.error("--gpus exceeds the number of CUDA devices");
  if (params.missme) cao.error("--emu with --gpus > 1 is not supported yet (flip_UV across SNP shards)");
  std::vector<int> devs(g);
  for (int i = 0; i < g; ++i) devs[i] = params.device + i;
  std::vector<ncclComm_t> comms(g);
  if (ncclCommInitAll(comms.data(), g, devs.data()) != ncclSuccess) cao.error("ncclCommInitAll failed");
  Mat2D Vfull;
  Mat1D Ffull;
  Barrier bar(g);
  std::vector<std::string> errors(g);
  std::atomic<bool> failed{false};
  std::mutex abort_m;
  bool aborted = false;
  auto abort_all = [&]() {  // peers blocked inside a collective return with an error instead of hanging
    std::lock_guard<std::mutex> lk(abort_m);
    if (aborted) return;
    aborted = true;
    for (auto& cm : comms) ncclCommAbort(cm);
  };
  auto worker = [&](int rank) {
    Logger::muted = rank != 0;
    try {
      FileBed data(params, rank, g);
      data.prepare();
      // the library enqueues its collectives on this communicator itself (include/pcaone_b200.h)
      data.check(pcaone_comm_attach(data.ctx, (void*)comms[rank]));
      run_pca_with_halko(&data, params, [&](RsvdOpData* op) {
        if (rank == 0) {
          Vfull.resize(data.nsnps, op->ranks());
          Ffull.resize(data.nsnps);
        }
        bar.wait();
        Mat1D Floc(data.nsnps_local);
        data.check(pcaone_get_F(data.ctx, Floc.data()));
        for (uint64 r = 0; r < data.nsnps_local; ++r) {
          const uint64 j = data.shard.snps[r];
          Ffull(j) = Floc(r);
          for (Index c = 0; c < op->ranks(); ++c) Vfull(j, c) = op->V(r, c);
        }
        bar.wait();
        if (rank == 0) {
          op->V = Vfull;
          data.F = Ffull;
        }
      });
      if (rank == 0) cao.print(tick.date(), "total elapsed reading time: ", data.readtime, " seconds");
    } catch (const std::exception& e) {
      errors[rank] = e.what();
      failed = true;
      bar.brk();
      abort_all();
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < g; ++r) th.emplace_back(worker, r);
  for (auto& t : th) t.join();
  if (!aborted)
    for (auto& c : comms) ncclCommDestroy(c);
  if (failed)
    for (auto& e : errors)
      if (!e.empty()) cao.error(e);
#endif
}

}  // namespace pcaone_host
