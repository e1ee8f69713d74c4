// pcaone_b200 — single-CTA FP64 kernels on the l x l core (l <= 112):
// Cholesky + triangular inverse (CholeskyQR2), one-sided Jacobi SVD (the "small on-device SVD"
// that replaces Eigen::JacobiSVD of the l x N matrix B, Halko.cpp:69), l x l products, MEV.
// Matrices are row-major with leading dimension LS.
#pragma once
#include "common.cuh"

namespace pcaone {

constexpr int kSmallThreads = 1024;

// status[0] = 0 ok / 1 breakdown (pivot <= tol * max diag: numerically rank deficient)
// W (l x l, symmetric) -> R upper (W = R^T R), Rinv upper. W is not modified.
__global__ void __launch_bounds__(kSmallThreads)
k_chol_inv(const double* __restrict__ W, int l, int ld, double* __restrict__ R, double* __restrict__ Rinv,
           int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* A = reinterpret_cast<double*>(smem_raw);  // l x l, ld = l
  __shared__ double s_piv, s_maxd;
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < l * l; i += nt) A[i] = W[(i / l) * ld + (i % l)];
  if (tid == 0) {
    s_fail = 0;
    double m = 0.0;
    for (int i = 0; i < l; ++i) m = fmax(m, W[i * ld + i]);
    s_maxd = m;
  }
  __syncthreads();
  const double tol = 64.0 * l * 2.220446049250313e-16 * s_maxd;
  for (int j = 0; j < l; ++j) {
    if (tid == 0) {
      const double d = A[j * l + j];
      if (!(d > tol)) s_fail = 1;
      s_piv = sqrt(d > tol ? d : 1.0);
      A[j * l + j] = s_piv;
    }
    __syncthreads();
    if (s_fail) break;
    const double inv = 1.0 / s_piv;
    for (int c = j + 1 + tid; c < l; c += nt) A[j * l + c] *= inv;
    __syncthreads();
    const int m = l - j - 1;  // trailing update on the upper triangle r <= c
    for (int idx = tid; idx < m * m; idx += nt) {
      const int r = j + 1 + idx / m, c = j + 1 + idx % m;
      if (c >= r) A[r * l + c] -= A[j * l + r] * A[j * l + c];
    }
    __syncthreads();
  }
  if (s_fail) {
    if (tid == 0) status[0] = 1;
    return;
  }
  if (tid == 0) status[0] = 0;
  for (int i = tid; i < l * l; i += nt) {
    const int r = i / l, c = i % l;
    R[r * ld + c] = c >= r ? A[i] : 0.0;
  }
  // inverse of upper-triangular R, one column per thread (back substitution)
  for (int c = tid; c < l; c += nt) {
    for (int r = l - 1; r > c; --r) Rinv[r * ld + c] = 0.0;
    Rinv[c * ld + c] = 1.0 / A[c * l + c];
    for (int r = c - 1; r >= 0; --r) {
      double s = 0.0;
      for (int p = r + 1; p <= c; ++p) s += A[r * l + p] * Rinv[p * ld + c];
      Rinv[r * ld + c] = -s / A[r * l + r];
    }
  }
}

// C (m x n) = op(A) (m x p) * op(B) (p x n), tiny sizes, row-major ld.
__global__ void k_small_matmul(const double* __restrict__ A, int transA, const double* __restrict__ B, int transB,
                               int m, int p, int n, int ld, double* __restrict__ C) {
  for (int idx = threadIdx.x; idx < m * n; idx += blockDim.x) {
    const int r = idx / n, c = idx - r * n;
    double s = 0.0;
    for (int q = 0; q < p; ++q) {
      const double a = transA ? A[q * ld + r] : A[r * ld + q];
      const double b = transB ? B[c * ld + q] : B[q * ld + c];
      s += a * b;
    }
    C[r * ld + c] = s;
  }
}

// One-sided (Hestenes) Jacobi SVD of the l x l matrix A (row-major, ld): A = U diag(sigma) V^T.
// Outputs sigma (descending) and V (row-major l x l, columns = right singular vectors, same
// order). Round-robin pair ordering, one warp per column pair; all state in shared memory.
// If `symmetric_psd` the caller passes W = B^T B and wants its eigen-decomposition: then
// sigma_out = sqrt(singular values of W) and V = eigenvectors (fallback path for a rank-deficient
// Gram where Cholesky broke down).
__global__ void __launch_bounds__(kSmallThreads)
k_jacobi_svd(const double* __restrict__ Ain, int l, int ld, int symmetric_psd, double* __restrict__ sigma,
             double* __restrict__ Vout, int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* A = reinterpret_cast<double*>(smem_raw);  // column-major l x l: column j at A + j*l
  double* V = A + (size_t)l * l;                    // column-major l x l
  __shared__ int s_rot, s_big;
  __shared__ double s_norm[kMaxL];
  __shared__ int s_ord[kMaxL];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int i = tid; i < l * l; i += nt) {
    const int r = i / l, c = i % l;
    A[c * l + r] = Ain[r * ld + c];
    V[c * l + r] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int n = (l + 1) & ~1;  // even number of players
  const double eps = 1e-15;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    if (tid == 0) s_rot = s_big = 0;
    __syncthreads();
    for (int round = 0; round < n - 1; ++round) {
      // one HALF-warp per column pair: 64 pairs in flight per pass, so a round of l <= 128 columns is
      // one pass (a warp per pair took two passes at l = 80 and left half of its lanes idle at
      // l <= 40); the three dot products fold over 16 lanes
      for (int pi = (tid >> 4); pi < n / 2; pi += (nt >> 4)) {
        const int hl = tid & 15;
        const unsigned hmask = 0xffffu << (tid & 16);
        int p, q;
        if (pi == 0) {
          p = n - 1;
          q = round;
        } else {
          p = (round + pi) % (n - 1);
          q = (round - pi + (n - 1)) % (n - 1);
        }
        if (p > q) {
          const int tmp = p;
          p = q;
          q = tmp;
        }
        if (q >= l) continue;  // bye (odd l)
        double* ap = A + p * l;
        double* aq = A + q * l;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int r = hl; r < l; r += 16) {
          const double x = ap[r], y = aq[r];
          alpha += x * x;
          beta += y * y;
          gamma += x * y;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          alpha += __shfl_xor_sync(hmask, alpha, o);
          beta += __shfl_xor_sync(hmask, beta, o);
          gamma += __shfl_xor_sync(hmask, gamma, o);
        }
        // the rotation parameters are a serial chain every lane waits for: reciprocal / rsqrt
        // forms instead of three divisions and three square roots (same test, squared)
        if (gamma * gamma > (eps * eps) * (alpha * beta) && gamma != 0.0) {
          if (hl == 0) {
            s_rot = 1;
            if (gamma * gamma > 1e-16 * (alpha * beta)) s_big = 1;
          }
          const double zeta = (beta - alpha) * __drcp_rn(2.0 * gamma);
          const double z2 = 1.0 + zeta * zeta;
          const double tt = copysign(1.0, zeta) * __drcp_rn(fabs(zeta) + z2 * rsqrt(z2));
          const double c = rsqrt(1.0 + tt * tt), s = c * tt;
          double* vp = V + p * l;
          double* vq = V + q * l;
          for (int r = hl; r < l; r += 16) {
            const double x = ap[r], y = aq[r];
            ap[r] = c * x - s * y;
            aq[r] = s * x + c * y;
            const double vx = vp[r], vy = vq[r];
            vp[r] = c * vx - s * vy;
            vq[r] = s * vx + c * vy;
          }
        }
      }
      __syncthreads();
    }
    const int rot = s_rot, big = s_big;
    __syncthreads();
    // Cyclic Jacobi converges quadratically: a sweep that met no pair with |cos| > 1e-8 (and rotated
    // the ones above eps) leaves every pair below ~1e-16 — the sweep that would only confirm
    // "no rotations" is not run.
    if (!rot || !big) break;
  }
  // column norms -> singular values
  for (int j = warp; j < l; j += nw) {
    double s = 0.0;
    for (int r = lane; r < l; r += 32) s += A[j * l + r] * A[j * l + r];
    s = warp_sum(s);
    if (lane == 0) s_norm[j] = sqrt(s);
  }
  __syncthreads();
  if (tid == 0) {  // insertion sort, descending, stable
    for (int j = 0; j < l; ++j) s_ord[j] = j;
    for (int a = 1; a < l; ++a) {
      const int key = s_ord[a];
      int b = a - 1;
      while (b >= 0 && s_norm[s_ord[b]] < s_norm[key]) {
        s_ord[b + 1] = s_ord[b];
        --b;
      }
      s_ord[b + 1] = key;
    }
    if (sweeps_out) sweeps_out[0] = sweep;
  }
  __syncthreads();
  for (int j = tid; j < l; j += nt) {
    const double s = s_norm[s_ord[j]];
    sigma[j] = symmetric_psd ? sqrt(s) : s;
  }
  for (int i = tid; i < l * l; i += nt) {
    const int r = i / l, c = i % l;
    Vout[r * ld + c] = V[s_ord[c] * l + r];
  }
}

// Z (l x k, ld) = Vr[:, :k] * diag(1/sigma[:k]) ; sigma <= tiny -> column zeroed.
__global__ void k_scale_v_by_inv_sigma(const double* __restrict__ Vr, const double* __restrict__ sigma, int l,
                                       int k, int ld, double* __restrict__ Z) {
  for (int idx = threadIdx.x; idx < l * ld; idx += blockDim.x) {
    const int r = idx / ld, c = idx - r * ld;
    double v = 0.0;
    if (c < k) {
      const double s = sigma[c];
      v = (s > 1e-300 && s > 1e-14 * sigma[0]) ? Vr[r * ld + c] / s : 0.0;
    }
    Z[r * ld + c] = v;
  }
}

// SVQB-style factor for a numerically rank-deficient Gram W = V diag(lam) V^T:
// T = V diag(lam^-1/2), columns with lam <= tol*lam_max zeroed; Tinv^T = V diag(lam^1/2).
// sigma holds sqrt(lam) (from k_jacobi_svd with symmetric_psd=1).
__global__ void k_svqb_factor(const double* __restrict__ Vr, const double* __restrict__ sig, int l, int ld,
                              double* __restrict__ T) {
  const double smax = sig[0];
  for (int idx = threadIdx.x; idx < l * l; idx += blockDim.x) {
    const int r = idx / l, c = idx - r * l;
    const double s = sig[c];
    T[r * ld + c] = (s > 1e-7 * smax && s > 0.0) ? Vr[r * ld + c] / s : 0.0;
  }
}

// Column signs that make a CholeskyQR basis Q (R diagonal > 0) equal the thin Q an unblocked
// Householder QR would return (Eigen::HouseholderQR / LAPACK dgeqrf convention:
// beta = -sign(c0)*norm, Householder.h makeHouseholder). The decisions only depend on the top
// l x l block of Q: with orthonormal columns the i-th reflector turns the trailing block into
// its Schur complement w.r.t. (Q_ii - beta_i), so a sign-modified LU of that block replays them
// (Ballard et al., "Reconstructing Householder vectors from TSQR"). winSVD adds H1 and H2 that
// were built with different Omegas, so the column signs of Omega are part of the algorithm
// (that is what flipOmg is for) and must follow the reference's QR, not ours.
__global__ void __launch_bounds__(kSmallThreads)
k_householder_signs(const double* __restrict__ Qtop, int l, int ld, double* __restrict__ sign) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* W = reinterpret_cast<double*>(smem_raw);  // l x l row-major
  __shared__ double s_piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < l * l; i += nt) W[i] = Qtop[(i / l) * ld + (i % l)];
  __syncthreads();
  for (int i = 0; i < l; ++i) {
    if (tid == 0) {
      const double c0 = W[i * l + i];
      const double beta = (c0 >= 0.0) ? -1.0 : 1.0;
      sign[i] = beta;
      s_piv = c0 - beta;  // |c0 - beta| >= 1
    }
    __syncthreads();
    const double inv = 1.0 / s_piv;
    const int m = l - i - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int r = i + 1 + idx / m, c = i + 1 + idx % m;
      W[r * l + c] -= W[r * l + i] * W[i * l + c] * inv;
    }
    __syncthreads();
  }
}

// mev(X, Y) = mean_i || X^T Y[:, i] ||  (Utils.cpp:194-200); C = X^T Y (k x k, row-major ld)
__global__ void k_mev_from_xty(const double* __restrict__ C, int k, int ld, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double res = 0.0;
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int r = 0; r < k; ++r) s += C[r * ld + i] * C[r * ld + i];
      res += sqrt(s);
    }
    out[0] = res / k;
  }
}

}  // namespace pcaone
