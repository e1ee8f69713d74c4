// pcaone_b200 — the sample covariance GEMM and the symmetric SVD behind `--svd 3` and the PCAngsd GRM step.
#include <algorithm>
#include <numeric>

#include "ctx.hpp"
#include "sym_jacobi.cuh"

namespace pcaone {

// K = X X^T (N x N, column-major on the host) for the source and flags of the context: panels of l identity
// columns through the operator H = X (X^T Omega) (the IRAM operator with a block of unit vectors), on the FP64
// kernels whatever the context's GEMM precision — this is the exact path (Main.cpp:186-188 `G * G^T`,
// Halko.cpp:323 `data->G * data->G.transpose()`; the caller divides by nsnps).
void sample_covariance(pcaone_ctx* c, double* K_out) {
  if (c->source < 0) throw std::runtime_error("no genotype source set");
  if (c->shard_samples) throw std::runtime_error("sample_covariance: not available on a sample-sharded context");
  if (c->update && c->cfg.emu && !c->have_usv) throw std::runtime_error("sample_covariance(update) without U,S,V");
  c->lut.standardize = (c->standardize && c->cfg.scale == -9) ? 1 : 0;
  const int saved_slices = c->slices;
  c->slices = 0;
  try {
    for (uint64_t p0 = 0; p0 < c->N; p0 += (uint64_t)c->l) {
      const uint32_t ncol = (uint32_t)std::min<uint64_t>((uint64_t)c->l, c->N - p0);
      zero_async(c, c->d_Omg, c->N * c->lp);
      symj::k_identity_panel<<<(ncol + 127) / 128, 128, 0, c->stream>>>(c->d_Omg, c->lp, p0, ncol);
      PCA_CHECK_LAUNCH();
      c->omega_img_valid = c->omega_colmax_valid = false;
      walk_ranges(c);
      allreduce_H(c, c->d_H);
      download_colmajor(c, c->d_H, c->N, (int)ncol, K_out + p0 * c->N);
    }
  } catch (...) {
    c->slices = saved_slices;
    throw;
  }
  c->slices = saved_slices;
  c->omega_img_valid = c->omega_colmax_valid = false;
}

// The PCAngsd GRM step (Halko.cpp:320-326): E <- pcangsd_standardize_E(U, S, V), C = E E^T / nsnps with the diagonal
// replaced by Dc / nsnps. C_out: N x N column-major on the host; Dc_out (N) may be NULL. The dense operand of the
// context holds the standardised E afterwards (the next computeUSV rebuilds it at pi = 0).
void gl_grm(pcaone_ctx* c, double* C_out, double* Dc_out) {
  double* d_Dc = nullptr;
  try {
    dmalloc(&d_Dc, c->N);
    gl_grm_standardize(c, d_Dc);
    std::vector<double> dc(c->N);
    PCA_CUDA(cudaMemcpyAsync(dc.data(), d_Dc, c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    sample_covariance(c, C_out);  // synchronises the stream
    const double inv = 1.0 / (double)c->M;
    for (uint64_t e = 0; e < c->N * c->N; ++e) C_out[e] *= inv;
    for (uint64_t i = 0; i < c->N; ++i) C_out[i * c->N + i] = dc[i] * inv;
    if (Dc_out) std::copy(dc.begin(), dc.end(), Dc_out);
  } catch (...) {
    cudaFree(d_Dc);
    throw;
  }
  cudaFree(d_Dc);
}

// A (n x n symmetric, column-major, host) = U diag(S) V^T: S descending, U n x n column-major. One-sided Jacobi
// on the device; returns the number of sweeps.
int sym_svd(pcaone_ctx* c, const double* A, uint64_t n, double* U_out, double* S_out) {
  if (n == 0) return 0;
  if (n > (1ull << 20)) throw std::runtime_error("sym_svd: matrix too large");
  double *d_A = nullptr, *d_sig = nullptr;
  unsigned int* d_rot = nullptr;
  int sweeps = 0;
  try {
    dmalloc(&d_A, n * n);
    dmalloc(&d_sig, n);
    dmalloc(&d_rot, 1);
    PCA_CUDA(cudaMemcpyAsync(d_A, A, n * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const uint32_t np = (uint32_t)((n + 1) & ~1ull);
    const double tol = 8.0 * 2.220446049250313e-16;
    for (; sweeps < 60; ++sweeps) {
      PCA_CUDA(cudaMemsetAsync(d_rot, 0, sizeof(unsigned int), c->stream));
      for (uint32_t r = 0; r + 1 < np; ++r) {
        symj::k_jacobi_round<<<np / 2, symj::kThreads, 0, c->stream>>>(d_A, (uint32_t)n, np, r, tol, d_rot);
        c->tm.kernel_launches++;
      }
      PCA_CHECK_LAUNCH();
      unsigned int rot = 0;
      PCA_CUDA(cudaMemcpyAsync(&rot, d_rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      if (rot == 0) break;
    }
    symj::k_jacobi_finish<<<(unsigned)n, symj::kThreads, 0, c->stream>>>(d_A, (uint32_t)n, d_sig);
    PCA_CHECK_LAUNCH();
    std::vector<double> sig(n), Uraw(n * n);
    PCA_CUDA(cudaMemcpyAsync(sig.data(), d_sig, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaMemcpyAsync(Uraw.data(), d_A, n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](uint64_t a, uint64_t b) { return sig[a] > sig[b]; });
    for (uint64_t j = 0; j < n; ++j) {
      S_out[j] = sig[ord[j]];
      std::copy(Uraw.begin() + ord[j] * n, Uraw.begin() + (ord[j] + 1) * n, U_out + j * n);
    }
    c->tm.d2h_bytes += (n * n + n) * sizeof(double);
    c->tm.h2d_bytes += n * n * sizeof(double);
  } catch (...) {
    cudaFree(d_A);
    cudaFree(d_sig);
    cudaFree(d_rot);
    throw;
  }
  cudaFree(d_A);
  cudaFree(d_sig);
  cudaFree(d_rot);
  return sweeps + 1;
}

}  // namespace pcaone
