// pcaone_b200 — launchers of the FP64 DMMA products (packed, dense and dosage operands).
#include "ctx.hpp"
#include "gemm_fp64.cuh"
#include "dense_gemm.cuh"

namespace pcaone {

// ---------------------------------------------------------------- kernel dispatch on NT
template <int NT>
void gemm_g_nt(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, const double* F, double* G, const double* Vrows) {
  const size_t smem = 2 * GemmSmem<NT>::kStageG;
  const int grid = ceil_div(nrows, kTileRows);
  if (c->update && c->cfg.emu) {
    ensure_smem(c, k_gemm_g<NT, true>, smem);
    k_gemm_g<NT, true><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, c->d_Omg, F,
                                                                c->lut, G, c->d_U, c->lp, c->d_S, Vrows, c->lp, c->k);
  } else {
    ensure_smem(c, k_gemm_g<NT, false>, smem);
    k_gemm_g<NT, false><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, c->d_Omg, F,
                                                                 c->lut, G, nullptr, 0, nullptr, nullptr, 0, 0);
  }
  PCA_CHECK_LAUNCH();
}

template <int NT>
void gemm_h_nt(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, const double* F, const double* G, uint32_t splits,
               uint32_t rows_per_split, const double* Vrows) {
  const size_t smem = 2 * GemmSmem<NT>::kStageH;
  dim3 grid(ceil_div(c->N, kTileRows), splits);
  if (c->update && c->cfg.emu) {
    ensure_smem(c, k_gemm_h<NT, true>, smem);
    k_gemm_h<NT, true><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, G, F, c->lut,
                                                                c->d_Hpart, rows_per_split, c->d_U, c->lp, c->d_S,
                                                                Vrows, c->lp, c->k);
  } else {
    ensure_smem(c, k_gemm_h<NT, false>, smem);
    k_gemm_h<NT, false><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, G, F, c->lut,
                                                                 c->d_Hpart, rows_per_split, nullptr, 0, nullptr,
                                                                 nullptr, 0, 0);
  }
  PCA_CHECK_LAUNCH();
}

// G rows [0,nrows) of the range = X^T Omega ; Hacc (+)= X G      (FP64 DMMA kernels)
void range_gemms_fp64(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, uint64_t snp0, double* Hacc, bool accumulate) {
  if (nrows == 0) return;
  const double* F = c->d_F + snp0;
  double* G = c->d_G + snp0 * c->lp;
  const double* Vrows = c->d_V + snp0 * c->lp;
  if (c->half & 1) {
    Timed t(c, 0);
    NT_DISPATCH(gemm_g_nt, c, P, nrows, F, G, Vrows);
    c->tm.gemm_g_launches++;
    c->tm.kernel_launches++;
  }
  if (!(c->half & 2)) return;
  const uint32_t tiles = ceil_div(c->N, kTileRows);
  uint32_t splits = std::max<uint32_t>(1, (2u * c->sms + tiles - 1) / tiles);
  splits = std::min<uint32_t>(splits, c->max_splits);
  splits = std::min<uint32_t>(splits, (uint32_t)ceil_div(nrows, kKC));
  uint32_t rps = (uint32_t)round_up((size_t)ceil_div(nrows, splits), kKC);
  splits = ceil_div(nrows, rps);
  {
    Timed t(c, 1);
    NT_DISPATCH(gemm_h_nt, c, P, nrows, F, G, splits, rps, Vrows);
    const uint64_t count = c->N * c->lp;
    k_reduce_partials<<<grid_for(count, 256, c->sms), 256, 0, c->stream>>>(c->d_Hpart, splits, count, Hacc,
                                                                           accumulate ? 1 : 0);
    PCA_CHECK_LAUNCH();
    c->tm.gemm_h_launches++;
    c->tm.kernel_launches += 2;
  }
}

// ---------------------------------------------------------------- generic dense matrix (RsvdOpOnePass)
template <int NT>
void dense_g_nt(pcaone_ctx* c, const double* D, uint32_t nrows, double* G) {
  const size_t smem = 2 * DenseSmem<NT>::kStageG;
  ensure_smem(c, k_dense_g<NT>, smem);
  k_dense_g<NT><<<ceil_div(nrows, kDenseRows), kDenseThreads, smem, c->stream>>>(D, c->ldd, nrows, (uint32_t)c->N,
                                                                                   c->d_Omg, G);
  PCA_CHECK_LAUNCH();
}
template <int NT>
void dense_h_nt(pcaone_ctx* c, const double* D, uint32_t nrows, const double* G, uint32_t splits, uint32_t rps) {
  const size_t smem = 2 * DenseSmem<NT>::kStageH;
  ensure_smem(c, k_dense_h<NT>, smem);
  dim3 grid(ceil_div(c->N, kDenseRows), splits);
  k_dense_h<NT><<<grid, kDenseThreads, smem, c->stream>>>(D, c->ldd, nrows, (uint32_t)c->N, G, c->d_Hpart, rps);
  PCA_CHECK_LAUNCH();
}

// rows [r0, r0 + nrows) of the tall matrix: G rows = D_b Omega ; Hacc (+)= D_b^T G_b   (RSVD.hpp:139-144)
void range_gemms_dense(pcaone_ctx* c, uint64_t r0, uint32_t nrows, double* Hacc, bool accumulate) {
  const double* D = c->d_dense + r0 * c->ldd;
  double* G = c->d_G + r0 * c->lp;
  if (c->half & 1) {
    Timed t(c, 0);
    NT_DISPATCH(dense_g_nt, c, D, nrows, G);
    c->tm.gemm_g_launches++;
    c->tm.kernel_launches++;
  }
  if (!(c->half & 2)) return;
  const uint32_t tiles = ceil_div(c->N, kDenseRows);
  uint32_t splits = std::max<uint32_t>(1, (2u * c->sms + tiles - 1) / tiles);
  splits = std::min<uint32_t>(splits, c->max_splits);
  splits = std::min<uint32_t>(splits, (uint32_t)ceil_div(nrows, kDenseKC));
  const uint32_t rps = (uint32_t)round_up((size_t)ceil_div(nrows, splits), kDenseKC);
  splits = ceil_div(nrows, rps);
  {
    Timed t(c, 1);
    NT_DISPATCH(dense_h_nt, c, D, nrows, G, splits, rps);
    const uint64_t count = c->N * c->lp;
    k_reduce_partials<<<grid_for(count, 256, c->sms), 256, 0, c->stream>>>(c->d_Hpart, splits, count, Hacc,
                                                                           accumulate ? 1 : 0);
    PCA_CHECK_LAUNCH();
    c->tm.gemm_h_launches++;
    c->tm.kernel_launches += 2;
  }
}

// ---------------------------------------------------------------- BGEN-style dosages (FileBgen.cpp:15-168)
template <int NT>
void dos_g_nt(pcaone_ctx* c, const float* D, uint32_t nrows, const double* F, double* G) {
  const size_t smem = 2 * DenseSmem<NT>::kDosStageG;
  ensure_smem(c, k_dos_g<NT>, smem);
  k_dos_g<NT><<<ceil_div(nrows, kDenseRows), kDenseThreads, smem, c->stream>>>(D, c->ldf, nrows, (uint32_t)c->N, F,
                                                                                 c->lut, c->d_Omg, G);
  PCA_CHECK_LAUNCH();
}
template <int NT>
void dos_h_nt(pcaone_ctx* c, const float* D, uint32_t nrows, const double* F, const double* G, uint32_t splits,
              uint32_t rps) {
  const size_t smem = 2 * DenseSmem<NT>::kDosStageH;
  ensure_smem(c, k_dos_h<NT>, smem);
  dim3 grid(ceil_div(c->N, kDenseRows), splits);
  k_dos_h<NT><<<grid, kDenseThreads, smem, c->stream>>>(D, c->ldf, nrows, (uint32_t)c->N, F, c->lut, G, c->d_Hpart,
                                                        rps);
  PCA_CHECK_LAUNCH();
}

// variants [r0, r0 + nrows): G rows = X_b^T Omega ; Hacc (+)= X_b G_b with X decoded from float dosages
void range_gemms_dosage(pcaone_ctx* c, uint64_t r0, uint32_t nrows, double* Hacc, bool accumulate) {
  if (c->update && c->cfg.emu) throw std::runtime_error("--emu on a dosage source is not implemented");
  const float* D = c->d_dos + r0 * c->ldf;
  const double* F = c->d_F + r0;
  double* G = c->d_G + r0 * c->lp;
  if (c->half & 1) {
    Timed t(c, 0);
    NT_DISPATCH(dos_g_nt, c, D, nrows, F, G);
    c->tm.gemm_g_launches++;
    c->tm.kernel_launches++;
  }
  if (!(c->half & 2)) return;
  const uint32_t tiles = ceil_div(c->N, kDenseRows);
  uint32_t splits = std::max<uint32_t>(1, (2u * c->sms + tiles - 1) / tiles);
  splits = std::min<uint32_t>(splits, c->max_splits);
  splits = std::min<uint32_t>(splits, (uint32_t)ceil_div(nrows, kDenseKC));
  const uint32_t rps = (uint32_t)round_up((size_t)ceil_div(nrows, splits), kDenseKC);
  splits = ceil_div(nrows, rps);
  {
    Timed t(c, 1);
    NT_DISPATCH(dos_h_nt, c, D, nrows, F, G, splits, rps);
    const uint64_t count = c->N * c->lp;
    k_reduce_partials<<<grid_for(count, 256, c->sms), 256, 0, c->stream>>>(c->d_Hpart, splits, count, Hacc,
                                                                           accumulate ? 1 : 0);
    PCA_CHECK_LAUNCH();
    c->tm.gemm_h_launches++;
    c->tm.kernel_launches += 2;
  }
}

// Beagle / PCAngsd: (re)build the expected genotypes E from the likelihoods — with pt = F
// (FileBeagle.cpp:57-66) or, on update passes, the individual allele frequencies of the current
// U, S, V (Data::fit_with_pi, Data.cpp:296-316, called at pi == 0 by Halko.cpp:108-118).
void gl_refresh(pcaone_ctx* c, uint64_t r0, uint64_t nrows, bool update, double* E, uint32_t ldd) {
  if (!c->af_done) throw std::runtime_error("GL source: call pcaone_gl_em_maf (or pcaone_set_F) first");
  if (update && !c->have_usv) throw std::runtime_error("GL update pass without U,S,V");
  k_gl_expected<<<grid_for(nrows * c->N, 256, c->sms), 256, 0, c->stream>>>(
      c->d_P + r0 * 2ull * c->N, (uint32_t)c->N, nrows, c->d_F + r0, update ? c->d_U : nullptr, c->lp, c->d_S,
      c->d_V + r0 * c->lp, c->lp, c->k, E, ldd);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// pcangsd_standardize_E into the dense operand + the diagonal sums Dc (N doubles on the device, zeroed here)
void gl_grm_standardize(pcaone_ctx* c, double* d_Dc) {
  if (c->source != PCAONE_SRC_GL) throw std::runtime_error("gl_grm: the context has no genotype-likelihood source");
  if (!c->af_done || !c->have_usv) throw std::runtime_error("gl_grm: run the PCAngsd EM (pcaone_run_em) first");
  PCA_CUDA(cudaMemsetAsync(d_Dc, 0, c->N * sizeof(double), c->stream));
  const uint32_t gx = (uint32_t)ceil_div(c->N, 256);
  const uint32_t slices = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(c->M, (uint64_t)(4 * c->sms + gx - 1) / gx));
  const uint64_t rps = (c->M + slices - 1) / slices;
  k_gl_grm<<<dim3(gx, (unsigned)ceil_div(c->M, rps)), 256, 0, c->stream>>>(c->d_P, (uint32_t)c->N, c->M, rps, c->d_F, c->d_U, c->lp,
                                                                          c->d_S, c->d_V, c->lp, c->k, c->d_dense, c->ldd, d_Dc);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// ---------------------------------------------------------------- entry points of the dosage / GL / dense sources
void dosage_allele_freq(pcaone_ctx* c) {
  k_dosage_af<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_dos, c->ldf, (uint32_t)c->N, c->M, c->d_F,
                                                                      c->d_nmiss);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

void dosage_decode(pcaone_ctx* c, uint64_t start, uint64_t B, const LutParams& p, double* out) {
  ensure_stage(c, c->N * B);
  k_dosage_decode<<<grid_for(c->N * B, 256, c->sms), 256, 0, c->stream>>>(c->d_dos + start * c->ldf, c->ldf, (uint32_t)c->N,
                                                                         B, c->d_F + start, p, c->d_stage);
  PCA_CHECK_LAUNCH();
  PCA_CUDA(cudaMemcpyAsync(out, c->d_stage, c->N * B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
}

void dosage_sqnorm(pcaone_ctx* c, double* out) {
  k_dosage_sqnorm<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_dos, c->ldf, (uint32_t)c->N, c->M, c->d_F,
                                                                          c->lut, out);
  PCA_CHECK_LAUNCH();
}

// emMAF_with_GL (Utils.cpp:745-775): F = 0.25, EM steps until the RMS change over all variants < tolmaf
int gl_em_maf(pcaone_ctx* c, uint32_t maxiter, double tolmaf) {
  std::vector<double> f0(c->M, 0.25);
  PCA_CUDA(cudaMemcpyAsync(c->d_F, f0.data(), c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  ensure_stage(c, 2 * c->M + 8);
  double* Fnew = c->d_stage;
  double* sq = c->d_stage + c->M;
  int it = 0;
  for (; it < (int)maxiter; ++it) {
    k_gl_maf_step<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_P, (uint32_t)c->N, c->M, c->d_F, Fnew, sq);
    PCA_CHECK_LAUNCH();
    k_sum_fixed<<<1, 1024, 0, c->stream>>>(sq, c->M, c->d_scal);
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaMemcpyAsync(c->d_F, Fnew, c->M * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.kernel_launches += 2;
    if (sqrt(c->h_scal[0] / (double)c->M) < tolmaf) {
      ++it;
      break;
    }
  }
  return it;
}

void dense_transpose_in(pcaone_ctx* c, const double* stage) {
  dim3 grid((unsigned)ceil_div(c->M, 32), (unsigned)ceil_div(c->ldd, 32));
  k_dense_transpose_in<<<grid, 256, 0, c->stream>>>(stage, c->M, c->N, c->d_dense, c->ldd);
  PCA_CHECK_LAUNCH();
}

static_assert(kTileRows == kFp64TileRows, "ctx.hpp mirrors the tile geometry of gemm_fp64.cuh");

}  // namespace pcaone
