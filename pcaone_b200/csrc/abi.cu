// pcaone_b200 — the C-ABI of include/pcaone_b200.h.
#include <random>
#include "ctx.hpp"

using namespace pcaone;

namespace {
thread_local std::string g_create_err;

pcaone_ld_source ld_source_of(const double* G) {
  pcaone_ld_source s{};
  s.kind = G ? PCAONE_LD_DENSE_F64 : PCAONE_LD_PACKED;
  s.data = G;
  return s;
}

int supported_nt(int l) {
  static const int opts[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  const int need = (l + 7) / 8;
  for (int o : opts)
    if (o >= need) return o;
  return -1;
}
}  // namespace

// =============================================================================== C-ABI
extern "C" {

int pcaone_abi_version(void) { return 1; }

const char* pcaone_last_error(const pcaone_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pcaone_create(const pcaone_config* cfg, pcaone_ctx** out) {
  if (!cfg || !out) return 1;
  pcaone_ctx* c = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw std::runtime_error(std::string("pcaone_b200 needs a CUDA device (no CPU fallback): ") +
                               cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) throw std::runtime_error("invalid CUDA device ordinal");
    PCA_CUDA(cudaSetDevice(cfg->device));
    c = new pcaone_ctx();
    c->cfg = *cfg;
    if (c->cfg.world < 1) c->cfg.world = 1;
    if (c->cfg.nsnps_total == 0) c->cfg.nsnps_total = c->cfg.nsnps;
    if (c->cfg.bands == 0) c->cfg.bands = 64;
    c->N = cfg->nsamples;
    c->M = cfg->nsnps;
    c->shard_samples = cfg->shard_samples != 0 && c->cfg.world > 1;
    c->N_total = (c->shard_samples && cfg->nsamples_total) ? cfg->nsamples_total : cfg->nsamples;
    c->samp0 = c->shard_samples ? cfg->sample_offset : 0;
    if (c->shard_samples) {
      if (c->samp0 % 4 != 0) throw std::runtime_error("sample_offset must be a multiple of 4 (whole bed bytes)");
      if (c->samp0 + c->N > c->N_total) throw std::runtime_error("sample shard exceeds nsamples_total");
      if (cfg->precision == PCAONE_PREC_FP64 || cfg->emu)
        throw std::runtime_error("sample-sharded jobs run on the int8 route without --emu");
    }
    c->k = (int)cfg->k;
    c->l = (int)(cfg->k + cfg->oversamples);
    if (c->N == 0 || c->M == 0 || c->k == 0) throw std::runtime_error("nsamples, nsnps and k must be positive");
    if (c->l > kMaxL) throw std::runtime_error("k + oversamples must be <= 112");
    if ((uint64_t)c->l > c->N || (uint64_t)c->l > c->M) throw std::runtime_error("k + oversamples exceeds the matrix size");
    if (cfg->precision != PCAONE_PREC_FP64 && cfg->precision != PCAONE_PREC_INT8X2 &&
        cfg->precision != PCAONE_PREC_INT8X3 && cfg->precision != PCAONE_PREC_INT8X4)
      throw std::runtime_error("precision must be PCAONE_PREC_FP64 or PCAONE_PREC_INT8X2/3/4");
    if (cfg->svd != PCAONE_SVD_SSVD && cfg->svd != PCAONE_SVD_WINSVD) throw std::runtime_error("svd must be 1 or 2");
    c->NT = supported_nt(c->l);
    c->lp = c->NT * 8;
    if (cfg->precision != PCAONE_PREC_FP64) {
      c->slices = cfg->precision;
      c->NP = (int)round_up((size_t)c->slices * c->l, 16);
      if (c->NP > kTcMaxNP) {
        // the UMMA N dimension holds slices * l columns: beyond 256 the request runs on the FP64 DMMA
        // kernels instead (the reference has no bound on k); pcaone_precision() tells the host
        if (cfg->shard_samples) throw std::runtime_error("sample-sharded jobs need slices * (k + oversamples) <= 256");
        c->slices = 0;
        c->NP = 0;
        c->cfg.precision = PCAONE_PREC_FP64;
      } else {
        c->RT = c->NP <= 128 ? 2 : 1;
      }
    }
    c->bpr = (uint32_t)((c->N + 3) >> 2);
    c->pitch = (uint32_t)round_up(c->bpr, 16);
    c->lut.sqrt_ploidy = sqrt((double)cfg->ploidy);
    c->lut.standardize = 0;
    c->lut.mask = 0;
    cudaDeviceProp prop;
    PCA_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    c->sms = prop.multiProcessorCount;
    PCA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PCA_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    const size_t NL = c->N * c->lp, ML = c->M * c->lp, LL = (size_t)c->lp * c->lp;
    dmalloc(&c->d_Omg0, NL);
    dmalloc(&c->d_Omg, NL);
    dmalloc(&c->d_Omg2, NL);
    dmalloc(&c->d_H, NL);
    dmalloc(&c->d_Bt, NL);
    dmalloc(&c->d_Ucur, NL);
    dmalloc(&c->d_Upre, NL);
    dmalloc(&c->d_U, NL);
    if (cfg->svd == PCAONE_SVD_WINSVD) {
      dmalloc(&c->d_H1, NL);
      dmalloc(&c->d_H2, NL);
      PCA_CUDA(cudaMemset(c->d_H1, 0, NL * sizeof(double)));
      PCA_CUDA(cudaMemset(c->d_H2, 0, NL * sizeof(double)));
    }
    dmalloc(&c->d_G, ML);
    dmalloc(&c->d_V, ML);
    if (cfg->emu) dmalloc(&c->d_Vpre, ML);
    PCA_CUDA(cudaMemset(c->d_V, 0, ML * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_U, 0, NL * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_H, 0, NL * sizeof(double)));
    dmalloc(&c->d_S, c->lp);
    dmalloc(&c->d_F, c->M);
    dmalloc(&c->d_nmiss, c->M);
    PCA_CUDA(cudaMemset(c->d_F, 0, c->M * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_nmiss, 0xff, c->M * sizeof(uint32_t)));  // unknown -> treated as "has missing"
    const uint32_t tiles = ceil_div(c->N, kFp64TileRows);
    c->max_splits = std::max<uint32_t>(1, std::min<uint32_t>(64, (2u * c->sms + tiles - 1) / tiles));
    dmalloc(&c->d_Hpart, (size_t)c->max_splits * NL);
    for (double** p : {&c->d_W, &c->d_R, &c->d_Rinv, &c->d_T1, &c->d_T2, &c->d_T, &c->d_Vr, &c->d_Z}) dmalloc(p, LL);
    dmalloc(&c->d_sigma, c->lp);
    dmalloc(&c->d_sign, c->lp);
    dmalloc(&c->d_hsign, c->lp);
    dmalloc(&c->d_flipbuf, 4 * (size_t)c->lp);
    dmalloc(&c->d_scal, 64);
    dmalloc(&c->d_status, 8);
    PCA_CUDA(cudaMemset(c->d_status, 0, 8 * sizeof(int)));
    dmalloc(&c->d_jscratch, (size_t)2 * c->l * c->l + 2 * c->l + 8);
    if (const char* e = getenv("PCAONE_FUSED_ORTH")) c->fused_orth = atoi(e);
    if (const char* e = getenv("PCAONE_ORTH_ONE_SHOT")) c->one_shot_q = atoi(e);
    if (const char* e = getenv("PCAONE_OMEGA_SKIP2")) c->omega_skip2 = atoi(e);
    if (const char* e = getenv("PCAONE_EMU_TC")) c->emu_tc = atoi(e);
    if (const char* e = getenv("PCAONE_EMU_SPLIT")) c->emu_split = atoi(e);
    PCA_CUDA(cudaHostAlloc((void**)&c->h_status, 4 * sizeof(int), cudaHostAllocDefault));
    PCA_CUDA(cudaHostAlloc((void**)&c->h_scal, 64 * sizeof(double), cudaHostAllocDefault));
    c->part_doubles = (size_t)(2 * c->sms + 8) * 128 * c->lp;
    dmalloc(&c->d_part, c->part_doubles);
    dmalloc(&c->d_pidx, (size_t)(2 * c->sms + 8) * 128);
    PCA_CUDA(cudaDeviceSynchronize());
    *out = c;
    return 0;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    delete c;
    return 1;
  }
}

void pcaone_destroy(pcaone_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  comm_destroy(c);
  for (void* p : {(void*)c->d_Omg0, (void*)c->d_packed, (void*)c->d_F, (void*)c->d_nmiss, (void*)c->d_Omg, (void*)c->d_Omg2,
                  (void*)c->d_H, (void*)c->d_H1, (void*)c->d_H2, (void*)c->d_Bt, (void*)c->d_Ucur, (void*)c->d_Upre,
                  (void*)c->d_U, (void*)c->d_G, (void*)c->d_V, (void*)c->d_Vpre, (void*)c->d_S, (void*)c->d_Hpart,
                  (void*)c->d_W, (void*)c->d_R, (void*)c->d_Rinv, (void*)c->d_T1, (void*)c->d_T2, (void*)c->d_T,
                  (void*)c->d_Vr, (void*)c->d_Z, (void*)c->d_sigma, (void*)c->d_sign, (void*)c->d_hsign, (void*)c->d_scal,
                  (void*)c->d_status, (void*)c->d_part, (void*)c->d_pidx, (void*)c->d_stage, (void*)c->d_raw[0],
                  (void*)c->d_raw[1], (void*)c->d_blk[0], (void*)c->d_blk[1], (void*)c->d_PG, (void*)c->d_PH, (void*)c->d_PGb[0],
                  (void*)c->d_PGb[1], (void*)c->d_PHb[0], (void*)c->d_PHb[1], (void*)c->d_BimgO, (void*)c->d_BimgW,
                  (void*)c->d_dense, (void*)c->d_P, (void*)c->d_dos, (void*)c->d_Racc, (void*)c->d_Racc2, (void*)c->d_BimgD, (void*)c->d_tcs, (void*)c->d_Fpart, (void*)c->d_jscratch,
                  (void*)c->d_flipbuf, (void*)c->d_cnt, (void*)c->d_cache_pg, (void*)c->d_cache_ph, (void*)c->d_emu_us, (void*)c->d_emu_part})
    if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (c->h_pin[i]) cudaFreeHost(c->h_pin[i]);
    if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  }
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  for (auto& e : c->evs) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (c->bed_file) fclose(c->bed_file);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

int pcaone_precision(const pcaone_ctx* c) { return c ? c->cfg.precision : -1; }
void* pcaone_stream(pcaone_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pcaone_alloc_pinned(void** out, size_t bytes) {
  if (!out) return 1;
  *out = nullptr;
  return cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable) == cudaSuccess ? 0 : 1;
}
void pcaone_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
int pcaone_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int pcaone_sync(pcaone_ctx* c) { CTX_GUARD(c, PCA_CUDA(cudaStreamSynchronize(c->stream))); }
int pcaone_set_allreduce(pcaone_ctx* c, pcaone_allreduce_fn fn, void* user) {
  CTX_GUARD(c, {
    c->allreduce = fn;
    c->allreduce_user = user;
  });
}

int pcaone_set_flags(pcaone_ctx* c, int update, int standardize) {
  CTX_GUARD(c, {
    c->update = update;
    c->standardize = standardize;
  });
}
int pcaone_set_omega(pcaone_ctx* c, const double* Omg) {
  CTX_GUARD(c, {
    upload_colmajor(c, Omg, c->N, c->l, c->d_Omg0);
    const size_t nb = c->N * c->lp * sizeof(double);
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg, c->d_Omg0, nb, cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg2, c->d_Omg0, nb, cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->have_omg0 = true;
  });
}
int pcaone_get_omega(pcaone_ctx* c, double* Omg) { CTX_GUARD(c, download_colmajor(c, c->d_Omg, c->N, c->l, Omg)); }
int pcaone_set_usv(pcaone_ctx* c, const double* U, const double* S, const double* V) {
  CTX_GUARD(c, {
    upload_colmajor(c, U, c->N, c->k, c->d_U);
    upload_colmajor(c, V, c->M, c->k, c->d_V);
    PCA_CUDA(cudaMemcpyAsync(c->d_S, S, c->k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->have_usv = true;
  });
}
int pcaone_get_usv(pcaone_ctx* c, double* U, double* S, double* V) {
  CTX_GUARD(c, {
    if (U) download_colmajor(c, c->d_U, c->N, c->k, U);
    if (V) download_colmajor(c, c->d_V, c->M, c->k, V);
    if (S) {
      PCA_CUDA(cudaMemcpyAsync(S, c->d_S, c->k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
    }
  });
}
int pcaone_get_GH(pcaone_ctx* c, double* G, double* H) {
  CTX_GUARD(c, {
    if (G) download_colmajor(c, c->d_G, c->M, c->l, G);
    if (H) download_colmajor(c, c->d_H, c->N, c->l, H);
  });
}
int pcaone_set_H(pcaone_ctx* c, const double* H) { CTX_GUARD(c, upload_colmajor(c, H, c->N, c->l, c->d_H)); }

int pcaone_compute_gandh(pcaone_ctx* c, int pi) { CTX_GUARD(c, compute_gandh(c, pi)); }
int pcaone_small_stage(pcaone_ctx* c) { CTX_GUARD(c, small_stage(c)); }
int pcaone_compute_usv(pcaone_ctx* c, int maxp, double tol, double* diff_out, int* epochs_out) {
  CTX_GUARD(c, {
    compute_usv(c, maxp, tol);
    if (diff_out) *diff_out = c->last_diff;
    if (epochs_out) *epochs_out = c->last_epochs;
  });
}
int pcaone_run_em(pcaone_ctx* c, int* iters_out) { CTX_GUARD(c, run_em(c, iters_out)); }
int pcaone_orth_omega(pcaone_ctx* c, int flip) { CTX_GUARD(c, update_omega(c, c->d_H, flip != 0)); }

int pcaone_mev(pcaone_ctx* c, const double* X, const double* Y, uint64_t rows, uint32_t cols, double* out) {
  CTX_GUARD(c, {
    if ((int)cols > c->lp) throw std::runtime_error("mev: too many columns");
    double *dx = nullptr, *dy = nullptr;
    dmalloc(&dx, rows * c->lp);
    dmalloc(&dy, rows * c->lp);
    ensure_stage(c, rows * cols);
    for (int i = 0; i < 2; ++i) {
      PCA_CUDA(cudaMemcpyAsync(c->d_stage, i ? Y : X, rows * cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      colmajor_to_rowmajor(c, c->d_stage, rows, (int)cols, i ? dy : dx);
    }
    const int ksave = c->k;
    c->k = (int)cols;
    double r = 0.0;
    try {
      r = device_mev(c, dx, dy, rows, false);
    } catch (...) {
      c->k = ksave;
      throw;
    }
    c->k = ksave;
    cudaFree(dx);
    cudaFree(dy);
    *out = r;
  });
}

int pcaone_upload_dense(pcaone_ctx* c, const double* A, uint64_t rows, uint64_t cols) {
  CTX_GUARD(c, {
    const bool trans = rows < cols;  // RSVD.hpp:113-121: a wide matrix is used transposed
    const uint64_t nrow = trans ? cols : rows, ncol = trans ? rows : cols;
    if (nrow != c->M || ncol != c->N)
      throw std::runtime_error("upload_dense: context must be created with nsnps = max(rows, cols), nsamples = min(rows, cols)");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_dense: the dense source runs in FP64");
    if (c->cfg.world > 1) throw std::runtime_error("upload_dense: single-GPU only");
    c->ldd = (uint32_t)round_up(c->N, 8);
    if (!c->d_dense) dmalloc(&c->d_dense, c->M * (size_t)c->ldd);
    if (trans) {
      // A^T in row-major is A in column-major: rows of length N, re-pitched to ldd
      PCA_CUDA(cudaMemsetAsync(c->d_dense, 0, c->M * (size_t)c->ldd * sizeof(double), c->stream));
      PCA_CUDA(cudaMemcpy2DAsync(c->d_dense, (size_t)c->ldd * sizeof(double), A, c->N * sizeof(double),
                                 c->N * sizeof(double), c->M, cudaMemcpyHostToDevice, c->stream));
    } else {
      double* stage = nullptr;
      dmalloc(&stage, c->M * c->N);
      PCA_CUDA(cudaMemcpyAsync(stage, A, c->M * c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      dense_transpose_in(c, stage);
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(stage);
    }
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.h2d_bytes += c->M * c->N * sizeof(double);
    c->source = PCAONE_SRC_DENSE;
  });
}

int pcaone_upload_dense_data(pcaone_ctx* c, const double* G) {
  CTX_GUARD(c, {
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_dense_data: the dense source runs in FP64");
    if (c->cfg.world > 1) throw std::runtime_error("upload_dense_data: single-GPU only");
    c->ldd = (uint32_t)round_up(c->N, 8);
    if (!c->d_dense) dmalloc(&c->d_dense, c->M * (size_t)c->ldd);
    // column j of the column-major N x M matrix is row j of the feature-major operand, re-pitched to ldd
    PCA_CUDA(cudaMemsetAsync(c->d_dense, 0, c->M * (size_t)c->ldd * sizeof(double), c->stream));
    PCA_CUDA(cudaMemcpy2DAsync(c->d_dense, (size_t)c->ldd * sizeof(double), G, c->N * sizeof(double), c->N * sizeof(double),
                               c->M, cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.h2d_bytes += c->M * c->N * sizeof(double);
    c->source = PCAONE_SRC_DENSE;
  });
}

int pcaone_ld_prune(pcaone_ctx* c, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
                    const double* af, double r2_tol, uint8_t* keep_out) {
  CTX_GUARD(c, {
    if (!keep_out) throw std::runtime_error("ld_prune: keep_out is NULL");
    ld_r2(c, ld_source_of(G), nsnps, ws, we, nwin, nullptr, af, r2_tol, keep_out);
  });
}

int pcaone_xt_times(pcaone_ctx* c, const double* A, uint32_t ncols, double* out, double* sqnorm) {
  CTX_GUARD(c, xt_times(c, A, ncols, out, sqnorm));
}
int pcaone_x_times(pcaone_ctx* c, const double* B, uint32_t ncols, double* out) { CTX_GUARD(c, x_times(c, B, ncols, out)); }

int pcaone_mask_times(pcaone_ctx* c, const double* B, uint32_t ncols, double* out) {
  CTX_GUARD(c, {
    if (c->source != PCAONE_SRC_RESIDENT && c->source != PCAONE_SRC_HOST && c->source != PCAONE_SRC_FILE)
      throw std::runtime_error("mask_times: needs a packed genotype source");
    const int upd = c->update;
    c->lut.mask = 1;
    c->update = 0;  // the indicator of the calls themselves, never an EMU fill
    try {
      x_times(c, B, ncols, out);
    } catch (...) {
      c->lut.mask = 0;
      c->update = upd;
      throw;
    }
    c->lut.mask = 0;
    c->update = upd;
  });
}

int pcaone_perform_op(pcaone_ctx* c, const double* x_in, double* y_out) { CTX_GUARD(c, perform_op(c, x_in, y_out)); }

int pcaone_sample_covariance(pcaone_ctx* c, double* K_out) { CTX_GUARD(c, sample_covariance(c, K_out)); }

int pcaone_gl_grm(pcaone_ctx* c, double* C_out, double* Dc_out) { CTX_GUARD(c, gl_grm(c, C_out, Dc_out)); }

int pcaone_sym_svd(pcaone_ctx* c, const double* A, uint64_t n, double* U_out, double* S_out, int* sweeps_out) {
  CTX_GUARD(c, {
    const int sw = sym_svd(c, A, n, U_out, S_out);
    if (sweeps_out) *sweeps_out = sw;
  });
}

int pcaone_upload_dosage(pcaone_ctx* c, const float* dosage, uint64_t nsnps, int device_ptr) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("upload_dosage: nsnps does not match the context");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_dosage: the dosage source runs in FP64");
    if (c->cfg.emu) throw std::runtime_error("--emu on a dosage source is not implemented");
    c->ldf = (uint32_t)round_up(c->N, 8);
    if (!c->d_dos) dmalloc(&c->d_dos, c->M * (size_t)c->ldf);
    PCA_CUDA(cudaMemsetAsync(c->d_dos, 0, c->M * (size_t)c->ldf * sizeof(float), c->stream));
    PCA_CUDA(cudaMemcpy2DAsync(c->d_dos, (size_t)c->ldf * sizeof(float), dosage, c->N * sizeof(float),
                               c->N * sizeof(float), c->M, device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    if (!device_ptr) c->tm.h2d_bytes += c->M * c->N * sizeof(float);
    c->source = PCAONE_SRC_DOSAGE;
    c->af_done = false;
  });
}

int pcaone_upload_gl(pcaone_ctx* c, const double* P, uint64_t nsnps, int device_ptr) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("upload_gl: nsnps does not match the context");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_gl: genotype likelihoods run in FP64");
    if (c->cfg.emu) throw std::runtime_error("upload_gl: --emu does not apply to genotype likelihoods (PCAngsd EM is pcaone_run_em with emu = 0)");
    if (c->cfg.world > 1) throw std::runtime_error("upload_gl: single-GPU only");
    c->ldd = (uint32_t)round_up(c->N, 8);
    if (!c->d_P) dmalloc(&c->d_P, c->M * 2 * c->N);
    if (!c->d_dense) dmalloc(&c->d_dense, c->M * (size_t)c->ldd);
    PCA_CUDA(cudaMemsetAsync(c->d_dense, 0, c->M * (size_t)c->ldd * sizeof(double), c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->d_P, P, c->M * 2 * c->N * sizeof(double),
                             device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    if (!device_ptr) c->tm.h2d_bytes += c->M * 2 * c->N * sizeof(double);
    c->source = PCAONE_SRC_GL;
    c->af_done = false;
  });
}

int pcaone_gl_em_maf(pcaone_ctx* c, uint32_t maxiter, double tolmaf, int* iters_out) {
  CTX_GUARD(c, {
    if (c->source != PCAONE_SRC_GL) throw std::runtime_error("gl_em_maf: call pcaone_upload_gl first");
    const int it = gl_em_maf(c, maxiter, tolmaf);
    if (iters_out) *iters_out = it;
    PCA_CUDA(cudaMemsetAsync(c->d_nmiss, 0, c->M * sizeof(uint32_t), c->stream));
    c->af_done = true;
  });
}

int pcaone_dense_rsvd(pcaone_ctx* c, uint32_t p, uint32_t windows, int finder) {
  CTX_GUARD(c, dense_onepass(c, p, windows, finder));
}

int pcaone_ld_r2(pcaone_ctx* c, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
                 double* r2_out) {
  CTX_GUARD(c, ld_r2(c, ld_source_of(G), nsnps, ws, we, nwin, r2_out, nullptr, 0.0, nullptr));
}

int pcaone_ld_r2_ex(pcaone_ctx* c, const pcaone_ld_source* src, uint64_t nsnps, const int32_t* ws, const int32_t* we,
                    uint64_t nwin, double* r2_out, const double* af, double r2_tol, uint8_t* keep_out) {
  CTX_GUARD(c, {
    if (!src) throw std::runtime_error("ld_r2_ex: source is NULL");
    ld_r2(c, *src, nsnps, ws, we, nwin, r2_out, af, r2_tol, keep_out);
  });
}

int pcaone_residuals_block(pcaone_ctx* c, uint64_t start, uint64_t stop, int ld_stats, float* out) {
  CTX_GUARD(c, residuals_block(c, start, stop, ld_stats, out));
}

int pcaone_get_timers(pcaone_ctx* c, pcaone_timers* out, int reset) {
  CTX_GUARD(c, {
    resolve_timers(c);
    c->tm.tc_ranges = c->tc_ranges;
    c->tm.fp64_ranges = c->fp64_ranges;
    c->tm.tc_miss_ranges = c->tc_miss_ranges;
    c->tm.tc_emu_ranges = c->tc_emu_ranges;
    if (out) *out = c->tm;
    if (reset) {
      c->tm = pcaone_timers{};
      c->tc_ranges = c->fp64_ranges = c->tc_miss_ranges = c->tc_emu_ranges = 0;
    }
  });
}
int pcaone_enable_timing(pcaone_ctx* c, int on) { CTX_GUARD(c, c->timing = on != 0); }

// ---- host helpers that must match the reference's libstdc++ streams bit for bit -------------
// RsvdOpData::initOmg (Halko.cpp:15-23) with StandardNormalRandom / UniformRandom
// (RSVD.hpp:20-59): std::default_random_engine seeded with `seed`, values drawn in
// column-major order (Eigen NullaryExpr evaluation order for a column-major MatrixXd).
int pcaone_init_omega(uint64_t rows, uint32_t cols, int seed, int gaussian, double* out) {
  auto rng = std::default_random_engine{};
  rng.seed(seed);
  const uint64_t n = rows * cols;
  if (gaussian) {
    std::normal_distribution<double> dist{0, 1};
    for (uint64_t i = 0; i < n; ++i) out[i] = dist(rng);
  } else {
    std::uniform_real_distribution<double> dist{-1, 1};
    for (uint64_t i = 0; i < n; ++i) out[i] = dist(rng);
  }
  return 0;
}
// permute_matrix (RSVD.hpp:61-71): std::shuffle of 0..n-1 with an UNSEEDED default engine
int pcaone_shuffle_indices(uint64_t n, uint32_t* out) {
  std::vector<int> idx(n);
  for (uint64_t i = 0; i < n; ++i) idx[i] = (int)i;
  auto rng = std::default_random_engine{};
  std::shuffle(idx.data(), idx.data() + n, rng);
  for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t)idx[i];
  return 0;
}

}  // extern "C"
