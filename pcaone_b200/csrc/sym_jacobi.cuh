// pcaone_b200 — SVD of a symmetric n x n matrix on the device (one-sided Jacobi, Hestenes).
//
// Callers: the exact PCA of `--svd 3` (Main.cpp:180-217: SelfAdjointEigenSolver of K = G G^T / nsnps) and the
// PCAngsd GRM step (Halko.cpp:320-334: JacobiSVD of the N x N covariance). For a symmetric matrix the SVD A = U S V^T
// has S = |eigenvalues| and U = eigenvectors (up to sign); K is positive semi-definite, so S are the eigenvalues.
//
// One-sided Jacobi works on the columns of A alone: a rotation of the column pair (p, q) makes a_p . a_q = 0; when
// every pair is orthogonal the column norms are the singular values and the normalised columns are U. A sweep is
// n - 1 rounds of a round-robin tournament, each round n / 2 disjoint pairs: one block per pair, one launch per
// round. The matrix (8 n^2 bytes: 50 MB at n = 2,504) stays in L2 between rounds.
#pragma once
#include "common.cuh"

namespace pcaone {
namespace symj {

constexpr int kThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) t += red[w];
  return t;
}

// round r of the tournament on np = n rounded up to even players: block i plays (p, q)
__global__ void __launch_bounds__(kThreads) k_jacobi_round(double* __restrict__ A, uint32_t n, uint32_t np, uint32_t r, double tol,
                                                           unsigned int* __restrict__ rotations) {
  __shared__ double red[kThreads / 32];
  const uint32_t i = blockIdx.x, m1 = np - 1;
  uint32_t p = i == 0 ? m1 : (r + i) % m1;
  uint32_t q = i == 0 ? r % m1 : (r + m1 - i) % m1;
  if (p >= n || q >= n) return;  // the padding player
  if (p > q) {
    const uint32_t t = p;
    p = q;
    q = t;
  }
  double* ap = A + (uint64_t)p * n;
  double* aq = A + (uint64_t)q * n;
  double al = 0.0, be = 0.0, ga = 0.0;
  for (uint32_t e = threadIdx.x; e < n; e += kThreads) {
    const double x = ap[e], y = aq[e];
    al += x * x;
    be += y * y;
    ga += x * y;
  }
  al = block_sum(al, red);
  be = block_sum(be, red);
  ga = block_sum(ga, red);
  if (!(fabs(ga) > tol * sqrt(al * be))) return;  // orthogonal already (or a zero column)
  // rotation that zeroes a_p . a_q; the larger norm ends up in the lower column index
  const double zeta = (be - al) / (2.0 * ga);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  const bool swap = (al - t * ga) < (be + t * ga);  // new norms^2: al - t ga, be + t ga
  for (uint32_t e = threadIdx.x; e < n; e += kThreads) {
    const double x = ap[e], y = aq[e];
    const double xn = cs * x - sn * y, yn = sn * x + cs * y;
    ap[e] = swap ? yn : xn;
    aq[e] = swap ? xn : yn;
  }
  if (threadIdx.x == 0) atomicAdd(rotations, 1u);
}

// sigma[j] = |a_j|, a_j <- a_j / sigma[j] (zero column when sigma is below tiny)
__global__ void __launch_bounds__(kThreads) k_jacobi_finish(double* __restrict__ A, uint32_t n, double* __restrict__ sigma) {
  __shared__ double red[kThreads / 32];
  double* a = A + (uint64_t)blockIdx.x * n;
  double s = 0.0;
  for (uint32_t e = threadIdx.x; e < n; e += kThreads) s += a[e] * a[e];
  s = sqrt(block_sum(s, red));
  const double inv = s > 0.0 ? 1.0 / s : 0.0;
  for (uint32_t e = threadIdx.x; e < n; e += kThreads) a[e] *= inv;
  if (threadIdx.x == 0) sigma[blockIdx.x] = s;
}

// Omega[p0 + j][j] = 1 for j < ncol (row-major [N][lp], zeroed by the caller)
__global__ void k_identity_panel(double* __restrict__ Omg, int lp, uint64_t p0, uint32_t ncol) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncol) Omg[(p0 + j) * lp + j] = 1.0;
}

}  // namespace symj
}  // namespace pcaone
