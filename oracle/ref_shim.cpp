// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// C-ABI shim around the UNMODIFIED PCAone reference classes so that tests and
// bench.py's cpu_baseline / --impl reference legs can drive the reference's own
// CPU implementation of the randomized-SVD hot path and read its in-memory
// doubles (the reference's text writers keep only 6 significant digits,
// src/Data.cpp:215).
//
// The reference sources are compiled where they lie (see oracle/Makefile); this
// file only includes their headers. Nothing here is copied from the reference:
// it is a driver that calls the public members of
//   Param            (src/Cmd.hpp:16-98)
//   Data / FileBed   (src/Data.hpp:9-59, src/FilePlink.hpp:8-47)
//   RsvdOpData, NormalRsvdOpData, FancyRsvdOpData (src/Halko.hpp:6-93)
//   permute_plink    (src/FilePlink.cpp:303)
//   flip_UV / mev    (src/Utils.cpp:118,194)
//   PCAone::flipOmg  (src/RSVD.hpp:80)
//   calc_sds / divide_pos_by_window (src/LD.cpp:48,154)
//   PCAone::RsvdOne  (src/RSVD.hpp:327-362, the dense-matrix front-end PCAoneR binds)
//   ArnoldiOpData::perform_op (src/Arnoldi.cpp:18-46, the IRAM operator)
//   FileBeagle       (src/FileBeagle.cpp:14-68; PCAngsd EM through the same ref_run_em)
#define _DECLARE_TOOLBOX_HERE
#include <omp.h>

#include <chrono>
#include <fstream>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "Cmd.hpp"
#include "Common.hpp"
#include "Data.hpp"
#include "FileBeagle.hpp"
#include "FileBinary.hpp"
#include "FileBgen.hpp"
#include "FileCsv.hpp"
#include "Projection.hpp"
#include "bgen/writer.h"
#include "FilePlink.hpp"
#include "Arnoldi.hpp"
#include "Halko.hpp"
#include "LD.hpp"
#include "RSVD.hpp"
#include "Utils.hpp"

namespace {

struct RefCtx {
  Param* params = nullptr;
  Data* data = nullptr;
  RsvdOpData* op = nullptr;
  Mat2D G, H;  // the locals of RsvdOpData::computeUSV (Halko.cpp:50), kept across epochs
  std::string err;
};

thread_local std::string g_err;

// counts the computeGandH calls (= epochs) of the last computeUSV; everything else is the reference's
template <class Base>
struct Counting : Base {
  using Base::Base;
  int calls = 0;
  void computeGandH(Mat2D& G, Mat2D& H, int pi) override {
    ++calls;
    Base::computeGandH(G, H, pi);
  }
};
int* g_calls(RsvdOpData* op) {
  if (auto* f = dynamic_cast<Counting<FancyRsvdOpData>*>(op)) return &f->calls;
  if (auto* n = dynamic_cast<Counting<NormalRsvdOpData>*>(op)) return &n->calls;
  return nullptr;
}

std::vector<std::string> split_ws(const char* s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string t;
  while (is >> t) out.push_back(t);
  return out;
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_threads(int n) { omp_set_num_threads(n); Eigen::setNbThreads(n); }
int ref_get_threads() { return omp_get_max_threads(); }

// Mirrors the PLINK branch of main(): Main.cpp:48-57 (Param, logger), :121-153
// (out-of-core permutation or plain FileBed), :168 (Data::prepare).
void* ref_open(const char* cmdline) {
  RefCtx* c = new RefCtx();
  int rc = guarded([&] {
    auto toks = split_ws(cmdline);
    std::vector<char*> argv;
    for (auto& t : toks) argv.push_back(const_cast<char*>(t.c_str()));
    c->params = new Param((int)argv.size(), argv.data());
    Param& params = *c->params;
    if (cao.cao.is_open()) cao.cao.close();
    cao.cao.open(params.fileout + ".log");
    cao.is_screen = false;
    const bool ooc_permutation = params.perm && params.out_of_core;
    if (params.file_t == FileType::BINARY) {
      c->data = new FileBin(params);
    } else if (params.file_t == FileType::PLINK) {
      if (ooc_permutation) {
        auto perm = permute_plink(params.filein, params.fileout, params.buffer, params.bands);
        c->data = new FileBed(params);
        c->data->perm = perm;
      } else {
        c->data = new FileBed(params);
      }
    } else if (params.file_t == FileType::BEAGLE) {
      c->data = new FileBeagle(params);   // genotype likelihoods; read_all runs emMAF_with_GL and builds E
    } else if (params.file_t == FileType::CSV) {
      c->data = new FileCsv(params);      // zstd-compressed CSV, in core (Main.cpp:160-161)
    } else if (params.file_t == FileType::BGEN) {
      c->data = new FileBgen(params);     // the reference's own BGEN reader (external/bgen)
    } else {
      throw std::runtime_error("ref_shim: only --bfile / --binary / --beagle / --bgen / --csv inputs are driven by the oracle");
    }
    c->data->prepare();
  });
  if (rc) {
    delete c;
    return nullptr;
  }
  return c;
}

void ref_close(void* h) {
  RefCtx* c = (RefCtx*)h;
  if (!c) return;
  delete c->op;
  delete c->data;
  delete c->params;
  delete c;
}

// Mirrors Halko.cpp:271-288: choose the op and set the initial flags.
int ref_new_op(void* h) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    const Param& p = *c->params;
    delete c->op;
    if (p.svd_t == SvdType::PCAoneAlg2)
      c->op = new Counting<FancyRsvdOpData>(c->data, p.k, p.oversamples);
    else
      c->op = new Counting<NormalRsvdOpData>(c->data, p.k, p.oversamples);
    if (p.genetic)
      c->op->setFlags(false, p.ld ? false : true);
    else
      c->op->setFlags(false, false);
    c->G = Mat2D::Zero(c->op->rows(), c->op->size());
    c->H = Mat2D::Zero(c->op->cols(), c->op->size());
  });
}

// dims: [nsamples, nsnps, k, l, nblocks, blocksize, bandFactor, bands, out_of_core, perm]
void ref_dims(void* h, long long* d) {
  RefCtx* c = (RefCtx*)h;
  const Param& p = *c->params;
  d[0] = c->data->nsamples;
  d[1] = c->data->nsnps;
  d[2] = p.k;
  d[3] = p.k + p.oversamples;
  d[4] = c->data->nblocks;
  d[5] = c->data->blocksize;
  d[6] = c->data->bandFactor;
  d[7] = p.bands;
  d[8] = p.out_of_core;
  d[9] = p.perm;
}

void ref_block_plan(void* h, unsigned* start, unsigned* stop) {
  RefCtx* c = (RefCtx*)h;
  for (size_t i = 0; i < c->data->start.size(); ++i) {
    start[i] = c->data->start[i];
    stop[i] = c->data->stop[i];
  }
}

void ref_get_F(void* h, double* out) {
  RefCtx* c = (RefCtx*)h;
  std::memcpy(out, c->data->F.data(), sizeof(double) * c->data->F.size());
}

void ref_get_lookup(void* h, double* out) {  // 4 x M, column-major (Arr2D)
  RefCtx* c = (RefCtx*)h;
  std::memcpy(out, c->data->centered_geno_lookup.data(), sizeof(double) * c->data->centered_geno_lookup.size());
}

long long ref_dataG_cols(void* h) { return ((RefCtx*)h)->data->G.cols(); }
void ref_get_dataG(void* h, double* out) {
  RefCtx* c = (RefCtx*)h;
  std::memcpy(out, c->data->G.data(), sizeof(double) * c->data->G.size());
}

long long ref_perm_size(void* h) { return ((RefCtx*)h)->data->perm.indices().size(); }
void ref_get_perm(void* h, int* out) {
  RefCtx* c = (RefCtx*)h;
  for (Eigen::Index i = 0; i < c->data->perm.indices().size(); ++i) out[i] = c->data->perm.indices()[i];
}

long long ref_missing_count(void* h) { return ((RefCtx*)h)->data->C.count(); }
void ref_get_mask(void* h, unsigned char* out) {
  RefCtx* c = (RefCtx*)h;
  for (Eigen::Index i = 0; i < c->data->C.size(); ++i) out[i] = c->data->C[i];
}

void ref_set_flags(void* h, int update, int standardize) { ((RefCtx*)h)->op->setFlags(update, standardize); }

void ref_get_omg(void* h, double* out) {
  RefCtx* c = (RefCtx*)h;
  std::memcpy(out, c->op->Omg.data(), sizeof(double) * c->op->Omg.size());
}

// One call of the pure virtual on the path: RsvdOpData::computeGandH (Halko.hpp:27).
int ref_gandh(void* h, int pi, double* G_out, double* H_out) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    c->op->computeGandH(c->G, c->H, pi);
    if (G_out) std::memcpy(G_out, c->G.data(), sizeof(double) * c->G.size());
    if (H_out) std::memcpy(H_out, c->H.data(), sizeof(double) * c->H.size());
  });
}

// seconds spent in one computeGandH call (the "power-iteration pass" of the metric)
double ref_time_gandh(void* h, int pi) {
  RefCtx* c = (RefCtx*)h;
  auto t0 = std::chrono::high_resolution_clock::now();
  int rc = guarded([&] { c->op->computeGandH(c->G, c->H, pi); });
  auto t1 = std::chrono::high_resolution_clock::now();
  if (rc) return -1.0;
  return std::chrono::duration<double>(t1 - t0).count();
}

// epochs (computeGandH calls) of the last ref_compute_usv
int ref_last_epochs(void* h) {
  int* n = g_calls(((RefCtx*)h)->op);
  return n ? *n : -1;
}

// RsvdOpData::computeUSV (Halko.cpp:46-97); returns U (N x k), S (k), V (M x k).
int ref_compute_usv(void* h, int maxp, double tol, double* U, double* S, double* V) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    if (int* n = g_calls(c->op)) *n = 0;
    c->op->computeUSV(maxp, tol);
    if (U) std::memcpy(U, c->op->U.data(), sizeof(double) * c->op->U.size());
    if (S) std::memcpy(S, c->op->S.data(), sizeof(double) * c->op->S.size());
    if (V) std::memcpy(V, c->op->V.data(), sizeof(double) * c->op->V.size());
  });
}

// The small dense stage alone (Halko.cpp:55-70) through the reference's computeU
// on caller-supplied G (M x l) and H (N x l): returns V_B[:, :k] (N x k).
int ref_compute_u(void* h, const double* G, const double* H, double* Uout) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    Eigen::Map<const Mat2D> Gm(G, c->op->rows(), c->op->size());
    Eigen::Map<const Mat2D> Hm(H, c->op->cols(), c->op->size());
    Mat2D U = c->op->computeU(Gm, Hm);
    std::memcpy(Uout, U.data(), sizeof(double) * U.size());
  });
}

// The EM driver of run_pca_with_halko (Halko.cpp:290-319) restated on the public
// members so that U,S,V stay in memory. Returns the number of EM iterations run.
int ref_run_em(void* h, double* U, double* S, double* V, int* iters) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    const Param& params = *c->params;
    RsvdOpData* rsvd = c->op;
    Mat2D Vpre;
    rsvd->setFlags(false, false);
    rsvd->computeUSV(params.maxp, params.tol);
    flip_UV(rsvd->U, rsvd->V, false);
    double diff;
    int it = 0;
    for (uint i = 0; i < params.maxiter; ++i) {
      rsvd->setFlags(true, false);
      Vpre = rsvd->V;
      rsvd->computeUSV(params.maxp, params.tol);
      flip_UV(rsvd->U, rsvd->V, false);
      diff = 1.0 - mev(rsvd->V, Vpre);
      it = i + 1;
      if (diff < params.tolem) break;
    }
    if (params.emu) {
      rsvd->setFlags(true, true);
      rsvd->computeUSV(params.maxp, params.tol);
      flip_UV(rsvd->U, rsvd->V, false);
    }
    if (iters) *iters = it;
    if (U) std::memcpy(U, rsvd->U.data(), sizeof(double) * rsvd->U.size());
    if (S) std::memcpy(S, rsvd->S.data(), sizeof(double) * rsvd->S.size());
    if (V) std::memcpy(V, rsvd->V.data(), sizeof(double) * rsvd->V.size());
  });
}

void ref_get_usv(void* h, double* U, double* S, double* V) {
  RefCtx* c = (RefCtx*)h;
  if (U) std::memcpy(U, c->op->U.data(), sizeof(double) * c->op->U.size());
  if (S) std::memcpy(S, c->op->S.data(), sizeof(double) * c->op->S.size());
  if (V) std::memcpy(V, c->op->V.data(), sizeof(double) * c->op->V.size());
}
void ref_set_usv(void* h, const double* U, const double* S, const double* V) {
  RefCtx* c = (RefCtx*)h;
  const long long N = c->op->cols(), M = c->op->rows(), k = c->op->ranks();
  c->op->U = Eigen::Map<const Mat2D>(U, N, k);
  c->op->S = Eigen::Map<const Mat1D>(S, k);
  c->op->V = Eigen::Map<const Mat2D>(V, M, k);
}

// Data virtuals on the path (Data.hpp:16-21), out-of-core mode.
int ref_read_block_initial(void* h, unsigned long long start, unsigned long long stop, int standardize,
                           double* out) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    if (start == c->data->start[0]) c->data->check_file_offset_first_var();
    c->data->read_block_initial(start, stop, standardize);
    if (out) std::memcpy(out, c->data->G.data(), sizeof(double) * c->data->nsamples * (stop - start + 1));
  });
}

int ref_read_block_update(void* h, unsigned long long start, unsigned long long stop, const double* U,
                          const double* S, const double* V, int k, int standardize, double* out) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    Eigen::Map<const Mat2D> Um(U, c->data->nsamples, k);
    Eigen::Map<const Mat1D> Sm(S, k);
    Eigen::Map<const Mat2D> Vm(V, c->data->nsnps, k);
    if (start == c->data->start[0]) c->data->check_file_offset_first_var();
    c->data->read_block_update(start, stop, Um, Sm, Vm.transpose(), standardize);
    if (out) std::memcpy(out, c->data->G.data(), sizeof(double) * c->data->nsamples * (stop - start + 1));
  });
}

// Data::fit_with_pi (Data.cpp:293-349) + standardize_E (:351-362), in-core.
int ref_fit_with_pi(void* h, const double* U, const double* S, const double* V, int k) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    Eigen::Map<const Mat2D> Um(U, c->data->nsamples, k);
    Eigen::Map<const Mat1D> Sm(S, k);
    Eigen::Map<const Mat2D> Vm(V, c->data->nsnps, k);
    c->data->fit_with_pi(Um, Sm, Vm.transpose());
  });
}
int ref_standardize_E(void* h) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] { c->data->standardize_E(); });
}

// helpers on the path: Utils.cpp:194 (mev), :118 (flip_UV), RSVD.hpp:80 (flipOmg)
double ref_mev(const double* X, const double* Y, long long rows, long long cols) {
  Eigen::Map<const Mat2D> Xm(X, rows, cols), Ym(Y, rows, cols);
  return mev(Xm, Ym);
}
void ref_flip_uv(double* U, long long urows, double* V, long long vrows, long long k) {
  Mat2D Um = Eigen::Map<Mat2D>(U, urows, k), Vm = Eigen::Map<Mat2D>(V, vrows, k);
  flip_UV(Um, Vm, false);
  std::memcpy(U, Um.data(), sizeof(double) * Um.size());
  std::memcpy(V, Vm.data(), sizeof(double) * Vm.size());
}
void ref_flip_omg(double* Omg2, double* Omg, long long rows, long long cols) {
  Mat2D A = Eigen::Map<Mat2D>(Omg2, rows, cols), B = Eigen::Map<Mat2D>(Omg, rows, cols);
  PCAone::flipOmg(A, B);
  std::memcpy(Omg2, A.data(), sizeof(double) * A.size());
  std::memcpy(Omg, B.data(), sizeof(double) * B.size());
}
// thin Q of HouseholderQR as used for Omega (Halko.cpp:121-122)
void ref_householder_q(const double* A, long long rows, long long cols, double* Q) {
  Mat2D Am = Eigen::Map<const Mat2D>(A, rows, cols);
  Eigen::HouseholderQR<Mat2D> qr(Am);
  Mat2D Qm = qr.householderQ() * Mat2D::Identity(rows, cols);
  std::memcpy(Q, Qm.data(), sizeof(double) * Qm.size());
}

// Omega exactly as RsvdOpData::initOmg draws it (Halko.cpp:15-23, RSVD.hpp:46-59)
void ref_init_omega(long long rows, long long cols, int seed, int gaussian, double* out) {
  auto rng = std::default_random_engine{};
  rng.seed(seed);
  Mat2D O;
  if (gaussian)
    O = PCAone::StandardNormalRandom<Mat2D, std::default_random_engine>(rows, cols, rng);
  else
    O = PCAone::UniformRandom<Mat2D, std::default_random_engine>(rows, cols, rng);
  std::memcpy(out, O.data(), sizeof(double) * O.size());
}
// the in-core column permutation of winSVD (RSVD.hpp:61-78), indices only
void ref_permute_indices(long long n, int* out) {
  PermMat P(n);
  P.setIdentity();
  auto rng = std::default_random_engine{};
  std::shuffle(P.indices().data(), P.indices().data() + P.indices().size(), rng);
  for (long long i = 0; i < n; ++i) out[i] = P.indices()[i];
}

// LD r2 (LD.cpp:450-473 ld_r2_big) without the gz text writer: windows from
// divide_pos_by_window (LD.cpp:154-168); r2 values appended in output order.
// Returns number of pairs, or -1 on error. Pass out=nullptr to count only.
long long ref_ld_r2(void* h, const char* filebim, int ld_bp, double* out, long long cap, int* ws_out,
                    int* we_out, long long* nwin) {
  RefCtx* c = (RefCtx*)h;
  long long n = 0;
  int rc = guarded([&] {
    SNPld snp;
    get_snp_pos_bim(snp, filebim);
    divide_pos_by_window(snp, ld_bp);
    const Mat2D& G = c->data->G;
    Arr1D sds = 1.0 / calc_sds(G);
    const double df = 1.0 / (G.rows() - 1);
    if (nwin) *nwin = (long long)snp.ws.size();
    for (int w = 0; w < (int)snp.ws.size(); w++) {
      int i = snp.ws[w];
      if (ws_out) ws_out[w] = snp.ws[w];
      if (we_out) we_out[w] = snp.we[w];
      for (int j = 1; j < snp.we[w]; j++) {
        int k = i + j;
        if (out && n < cap) {
          double r = G.col(i).dot(G.col(k)) * (sds(i) * sds(k) * df);
          out[n] = r * r;
        }
        n++;
      }
    }
  });
  return rc ? -1 : n;
}

// ld_prune_big (LD.cpp:240-268) on data->G with the windows of divide_pos_by_window; the keep mask
// is recovered from the reference's own writer output (<fileout>.ld.prune.in holds the kept lines
// of the bim in order): keep_out[i] = 1 iff line i of filebim was written there.
int ref_ld_prune(void* h, const char* filebim, int ld_bp, double r2_tol, const char* fileout, unsigned char* keep_out,
                 long long nsnps) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    SNPld snp;
    get_snp_pos_bim(snp, filebim);
    divide_pos_by_window(snp, ld_bp);
    ld_prune_big(c->data->G, snp, r2_tol, fileout, filebim);
    std::ifstream fin(filebim), fkept(std::string(fileout) + ".ld.prune.in");
    std::string line, kept;
    bool have = (bool)std::getline(fkept, kept);
    long long i = 0;
    while (std::getline(fin, line) && i < nsnps) {
      if (line.empty()) continue;
      // compare the first six tab/space separated fields
      auto norm = [](const std::string& s) {
        std::string o;
        int f = 0;
        bool in = false;
        for (char ch : s) {
          if (ch == ' ' || ch == '\t') {
            if (in) { ++f; in = false; if (f == 6) break; o.push_back('\t'); }
          } else { o.push_back(ch); in = true; }
        }
        while (!o.empty() && o.back() == '\t') o.pop_back();
        return o;
      };
      if (have && norm(line) == norm(kept)) {
        keep_out[i] = 1;
        have = (bool)std::getline(fkept, kept);
      } else {
        keep_out[i] = 0;
      }
      ++i;
    }
    if (have) throw std::runtime_error("ref_ld_prune: kept list does not align with the bim");
  });
}

// LD-based clumping of one association file on data->G (the clump branch of run_ld_stuff, LD.cpp:505-540):
// the reference's own valid_assoc_file / get_snp_pos_bim / map_index_snps / get_target_snp_idx /
// ld_clump_single_pheno; the result is the text file the reference writes.
int ref_ld_clump(void* h, const char* filebim, const char* assoc, const char* colnames, int clump_bp, double clump_r2,
                 double p1, double p2, const char* fileout) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    SNPld snp;
    get_snp_pos_bim(snp, filebim);
    const auto colidx = valid_assoc_file(assoc, colnames);
    SNPld snp_t;
    const std::string head = get_snp_pos_bim(snp_t, assoc, true, colidx);
    const auto pvals = map_index_snps(assoc, colidx, p2);
    Int2D idx_per_chr, bp_per_chr;
    std::tie(idx_per_chr, bp_per_chr) = get_target_snp_idx(snp_t, snp);
    ld_clump_single_pheno(fileout, head, clump_bp, clump_r2, p1, p2, c->data->G, idx_per_chr, bp_per_chr, pvals);
  });
}

// The FULL branch of main() (Main.cpp:180-217, `--svd 3`) for nsamples <= nsnps, on this run's in-core data:
// standardize_E, K = G G^T / nsnps, SelfAdjointEigenSolver, V = G^T U / svals, flip_UV — the reference's own
// Data / Eigen / flip_UV calls (main() itself cannot be linked). U: N x k, V: M x k, column-major.
int ref_full_pca(void* h, int k, double* U_out, double* svals_out, double* V_out, double* evals_out) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    Data* data = c->data;
    if (data->nsamples > data->nsnps) throw std::runtime_error("ref_full_pca: the sample-covariance branch only");
    data->standardize_E();
    const Eigen::Index ncomp = std::min<Eigen::Index>(k, std::min<Eigen::Index>(data->G.rows(), data->G.cols()));
    Mat2D K = (data->G * data->G.transpose()) / data->nsnps;
    Eigen::SelfAdjointEigenSolver<Mat2D> eig(K);
    if (eig.info() != Eigen::Success) throw std::runtime_error("eigendecomposition failed");
    Mat1D evals(ncomp), svals(ncomp);
    Mat2D U(data->nsamples, ncomp), V(data->nsnps, ncomp);
    for (Eigen::Index i = 0; i < ncomp; ++i) {
      const Eigen::Index idx = eig.eigenvalues().size() - 1 - i;
      evals(i) = std::max(0.0, eig.eigenvalues()(idx));
      U.col(i) = eig.eigenvectors().col(idx);
    }
    svals = (evals.array() * data->nsnps).sqrt();
    V.noalias() = data->G.transpose() * U;
    for (Eigen::Index i = 0; i < ncomp; ++i)
      if (svals(i) > 0) V.col(i) /= svals(i);
    flip_UV(U, V);
    std::memcpy(U_out, U.data(), sizeof(double) * U.size());
    std::memcpy(V_out, V.data(), sizeof(double) * V.size());
    std::memcpy(svals_out, svals.data(), sizeof(double) * ncomp);
    std::memcpy(evals_out, evals.data(), sizeof(double) * ncomp);
  });
}

// The GRM step of run_pca_with_halko for PCAngsd (Halko.cpp:320-334) after ref_run_em on a Beagle run: the
// reference's own pcangsd_standardize_E, the covariance with its Dc diagonal, Eigen's JacobiSVD. C, U2: N x N.
int ref_pcangsd_grm(void* h, double* C_out, double* U2_out, double* S2_out, double* Dc_out) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    Data* data = c->data;
    data->pcangsd_standardize_E(c->op->U, c->op->S, c->op->V.transpose());
    Mat2D C = data->G * data->G.transpose();
    C.array() /= (double)data->nsnps;
    C.diagonal() = data->Dc.array() / (double)data->nsnps;
    Eigen::JacobiSVD<Mat2D> svd(C, Eigen::ComputeThinU | Eigen::ComputeThinV);
    std::memcpy(C_out, C.data(), sizeof(double) * C.size());
    Mat2D U2 = svd.matrixU();
    std::memcpy(U2_out, U2.data(), sizeof(double) * U2.size());
    Mat1D S2 = svd.singularValues();
    std::memcpy(S2_out, S2.data(), sizeof(double) * S2.size());
    std::memcpy(Dc_out, data->Dc.data(), sizeof(double) * data->Dc.size());
  });
}

// A BGEN file (layout 1 or 2; compression 0 none / 1 zlib / 2 zstd; unphased diploid, bit_depth bits per probability) written with the writer of the
// reference's vendored library: probs = [nsnps][nsamples][3] genotype probabilities, NaN triples = missing.
int ref_write_bgen(const char* path, long long nsamples, long long nsnps, const double* probs, int bit_depth, int compression,
                   int layout) {
  return guarded([&] {
    std::string p(path), free_data;
    std::vector<std::string> samples;
    for (long long i = 0; i < nsamples; ++i) samples.push_back("s" + std::to_string(i));
    bgen::CppBgenWriter w(p, (std::uint32_t)nsamples, free_data, (uint32_t)compression, (uint32_t)layout, samples);
    std::vector<std::string> alleles{"A", "C"};
    std::vector<double> g((size_t)3 * nsamples);
    for (long long j = 0; j < nsnps; ++j) {
      std::string varid = "v" + std::to_string(j), rsid = "rs" + std::to_string(j), chrom = "1";
      std::uint32_t pos = (std::uint32_t)(100 * (j + 1));
      w.write_variant_header(varid, rsid, chrom, pos, alleles, (std::uint32_t)nsamples);
      std::copy(probs + (size_t)3 * nsamples * j, probs + (size_t)3 * nsamples * (j + 1), g.begin());
      w.add_genotype_data(2, g.data(), (std::uint32_t)g.size(), (std::uint8_t)2, false, (std::uint8_t)bit_depth);
    }
  });
}

// The dosages FileBgen::read_all sees (FileBgen.cpp:24-26: next_var().minor_allele_dosage), SNP-major floats,
// NaN = missing: read with the same library calls.
int ref_bgen_dosages(const char* path, float* out, long long nsamples, long long nsnps) {
  return guarded([&] {
    bgen::CppBgenReader bg(path, "", true);
    if ((long long)bg.header.nsamples != nsamples || (long long)bg.header.nvariants != nsnps)
      throw std::runtime_error("ref_bgen_dosages: size mismatch");
    for (long long j = 0; j < nsnps; ++j) {
      auto var = bg.next_var();
      var.minor_allele_dosage(out + (size_t)j * nsamples);
    }
  });
}

// The projection branch of main() (Main.cpp:95-107) for PLINK input: Param from the command line, FileBed,
// the reference's own run_projection (Projection.cpp:188-309), which writes <out>.eigvecs.
int ref_run_projection(const char* cmdline) {
  return guarded([&] {
    auto toks = split_ws(cmdline);
    std::vector<char*> argv;
    for (auto& t : toks) argv.push_back(const_cast<char*>(t.c_str()));
    Param params((int)argv.size(), argv.data());
    if (cao.cao.is_open()) cao.cao.close();
    cao.cao.open(params.fileout + ".log");
    cao.is_screen = false;
    if (params.file_t != FileType::PLINK || params.project < 1) throw std::runtime_error("ref_run_projection: -b ... --project 1|2");
    Data* data = new FileBed(params);
    run_projection(data, params);
    delete data;
  });
}

// Beagle input: the parsed likelihood matrix P (2N x M, column-major) as FileBeagle::read_all left it.
long long ref_get_P(void* h, double* out) {
  RefCtx* c = (RefCtx*)h;
  if (out) std::memcpy(out, c->data->P.data(), sizeof(double) * c->data->P.size());
  return (long long)c->data->P.size();
}

// Data::write_residuals (Data.cpp:242-291) via the reference writer.
int ref_write_residuals(void* h) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] { c->data->write_residuals(c->op->S, c->op->U, c->op->V.transpose()); });
}

// ArnoldiOpData::perform_op (Arnoldi.cpp:18-46) on an out-of-core run: y = sum_b G_b G_b^T x.
int ref_perform_op(void* h, int update, int standardize, const double* x, double* y) {
  RefCtx* c = (RefCtx*)h;
  return guarded([&] {
    ArnoldiOpData op(c->data);
    op.setFlags(update, standardize);
    op.perform_op(x, y);
  });
}

// PCAone::RsvdOne<MatrixXd> (RSVD.hpp:327-362) on a dense column-major matrix: the reference's own
// template, instantiated here. U: rows x k, S: k, V: cols x k. Returns non-zero if it throws.
int ref_rsvd_one(const double* A, long long rows, long long cols, int k, int os, int rand, int p, int windows,
                 int finder, double* U, double* S, double* V) {
  return guarded([&] {
    Eigen::Map<const Eigen::MatrixXd> M(A, rows, cols);
    Eigen::MatrixXd mat = M;
    PCAone::RsvdOne<Eigen::MatrixXd> rsvd(mat, (uint32_t)k, (uint32_t)os, (uint32_t)rand);
    rsvd.setRangeFinder(finder);
    rsvd.compute((uint32_t)p, (uint32_t)windows);
    Eigen::Map<Eigen::MatrixXd>(U, rows, k) = rsvd.matrixU();
    Eigen::Map<Eigen::VectorXd>(S, k) = rsvd.singularValues();
    Eigen::Map<Eigen::MatrixXd>(V, cols, k) = rsvd.matrixV();
  });
}

}  // extern "C"
