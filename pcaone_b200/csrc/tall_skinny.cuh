// pcaone_b200 — tall-skinny FP64 kernels for the orthonormalisation / small-SVD stage
// (reference: Householder QR + JacobiSVD in Halko.cpp:55-70, 120-124, 208-213).
// All tall matrices are row-major [rows][ld]; the l x l results are row-major with ld = LS.
// These are HBM-bound sweeps over rows*l*8 bytes; every reduction is two-stage with a fixed
// summation order, so results are run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace pcaone {

constexpr int kTsThreads = 256;
constexpr int kTsKR = 16;  // rows staged per step (keeps static smem < 48 KB at l = 128)

// Cpart[cta] (l1 x l2, ld = ldc) = A[chunk]^T * B[chunk]; thread (ty,tx) of a 16x16 grid owns
// C[ty+16i][tx+16j], i<RM, j<RN.
template <int RM, int RN>
__global__ void __launch_bounds__(kTsThreads)
k_ts_gemm_tn(const double* __restrict__ A, int lda, int l1, const double* __restrict__ B, int ldb, int l2,
             uint64_t rows, uint64_t rows_per_cta, double* __restrict__ Cpart, int ldc) {
  __shared__ double As[kTsKR][16 * RM + 1];
  __shared__ double Bs[kTsKR][16 * RN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const uint64_t r0 = blockIdx.x * rows_per_cta;
  const uint64_t r1 = min(rows, r0 + rows_per_cta);
  double acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.0;

  for (uint64_t r = r0; r < r1; r += kTsKR) {
    for (int idx = tid; idx < kTsKR * 16 * RM; idx += kTsThreads) {
      const int rr = idx / (16 * RM), c = idx - rr * (16 * RM);
      As[rr][c] = (r + rr < r1 && c < l1) ? A[(r + rr) * lda + c] : 0.0;
    }
    for (int idx = tid; idx < kTsKR * 16 * RN; idx += kTsThreads) {
      const int rr = idx / (16 * RN), c = idx - rr * (16 * RN);
      Bs[rr][c] = (r + rr < r1 && c < l2) ? B[(r + rr) * ldb + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < kTsKR; ++rr) {
      double a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[rr][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Bs[rr][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  double* C = Cpart + (uint64_t)blockIdx.x * ldc * (16 * RM);
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      if (r < l1 && c < l2) C[r * ldc + c] = acc[i][j];
    }
}

// C (l1 x l2) = sum over parts, fixed order
__global__ void k_reduce_small(const double* __restrict__ Cpart, int nparts, int part_stride, int l1, int l2,
                               int ldc, double* __restrict__ C) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < l1 * l2; idx += gridDim.x * blockDim.x) {
    const int r = idx / l2, c = idx - r * l2;
    // same order as before (p = 0, 1, ...), but 16 loads in flight: the plain loop was a chain of
    // nparts (~300) dependent L2 round trips, 81 us per call (ncu launch list)
    double v = 0.0;
    for (int p0 = 0; p0 < nparts; p0 += 16) {
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = (p0 + u < nparts) ? Cpart[(uint64_t)(p0 + u) * part_stride + r * ldc + c] : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u) v += t[u];
    }
    C[r * ldc + c] = v;
  }
}

// Out[rows][ldo] (first l2 cols) = A[rows][lda] (first l1 cols) * T (l1 x l2, ld = ldt); columns
// l2..ldo-1 of Out are zeroed. In-place (Out == A) is allowed: a CTA stages its 64 rows first.
template <int RN>
__global__ void __launch_bounds__(kTsThreads)
k_ts_rightmult(const double* __restrict__ A, int lda, int l1, const double* __restrict__ T, int ldt, int l2,
               uint64_t rows, double* __restrict__ Out, int ldo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ts = reinterpret_cast<double*>(smem_raw);  // [l1][16*RN]
  double* As = Ts + (size_t)l1 * 16 * RN;            // [64][l1+1]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lda_s = l1 + 1;
  for (int idx = tid; idx < l1 * 16 * RN; idx += kTsThreads) {
    const int r = idx / (16 * RN), c = idx - r * (16 * RN);
    Ts[idx] = c < l2 ? T[r * ldt + c] : 0.0;
  }
  for (uint64_t r0 = (uint64_t)blockIdx.x * 64; r0 < rows; r0 += (uint64_t)gridDim.x * 64) {
    __syncthreads();
    for (int idx = tid; idx < 64 * l1; idx += kTsThreads) {
      const int rr = idx / l1, c = idx - rr * l1;
      As[rr * lda_s + c] = (r0 + rr < rows) ? A[(r0 + rr) * lda + c] : 0.0;
    }
    __syncthreads();
    double acc[4][RN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] = 0.0;
    for (int a = 0; a < l1; ++a) {
      double av[4], tv[RN];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[(ty + 16 * i) * lda_s + a];
#pragma unroll
      for (int j = 0; j < RN; ++j) tv[j] = Ts[a * 16 * RN + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] += av[i] * tv[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint64_t r = r0 + ty + 16 * i;
      if (r < rows) {
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          const int c = tx + 16 * j;
          if (c < ldo) Out[r * ldo + c] = c < l2 ? acc[i][j] : 0.0;
        }
      }
    }
  }
}

// flipOmg (RSVD.hpp:80-89) stage 1: per-CTA partial column sums of |O2-O| and |O2+O|
// `pre` (optional) is a per-column sign applied to O first (the Householder sign convention).
__global__ void k_flip_partial(const double* __restrict__ O2, const double* __restrict__ O, int ld, int l,
                               uint64_t rows, uint64_t rows_per_cta, const double* __restrict__ pre,
                               double* __restrict__ part) {
  __shared__ double sm[2][8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const uint64_t r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (int c0 = 0; c0 < l; c0 += 32) {
    const int c = c0 + tx;
    double d = 0.0, s = 0.0;
    if (c < l) {
      const double ps = pre ? pre[c] : 1.0;
      for (uint64_t r = r0 + ty; r < r1; r += 8) {
        const double a = O2[r * ld + c], b = ps * O[r * ld + c];
        d += fabs(a - b);
        s += fabs(a + b);
      }
    }
    if (c < 128) {
      sm[0][ty][c] = d;
      sm[1][ty][c] = s;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < l; c += blockDim.x) {
    double d = 0.0, s = 0.0;
    for (int y = 0; y < 8; ++y) {
      d += sm[0][y][c];
      s += sm[1][y][c];
    }
    part[(uint64_t)blockIdx.x * 2 * l + c] = d;
    part[(uint64_t)blockIdx.x * 2 * l + l + c] = s;
  }
}
__global__ void k_flip_sign(const double* __restrict__ part, int nparts, int l, const double* __restrict__ pre,
                            double* __restrict__ sign) {
  for (int c = threadIdx.x; c < l; c += blockDim.x) {
    double d = 0.0, s = 0.0;
    for (int p = 0; p < nparts; ++p) {
      d += part[(uint64_t)p * 2 * l + c];
      s += part[(uint64_t)p * 2 * l + l + c];
    }
    sign[c] = ((d > 2 * s) ? -1.0 : 1.0) * (pre ? pre[c] : 1.0);
  }
}
// O[:,c] *= sign[c]; O2 = O
__global__ void k_flip_apply(double* __restrict__ O, double* __restrict__ O2, int ld, int l, uint64_t rows,
                             const double* __restrict__ sign) {
  const uint64_t total = rows * ld;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    double v = O[i];
    if (c < l) v *= sign[c];
    O[i] = v;
    O2[i] = v;
  }
}

// host col-major (rows x cols, ld = rows) <-> device row-major [rows][ld]
__global__ void k_colmajor_to_rowmajor(const double* __restrict__ src, uint64_t rows, int cols,
                                       double* __restrict__ dst, int ld) {
  __shared__ double tile[32][33];
  const uint64_t r0 = (uint64_t)blockIdx.x * 32;
  for (int c0 = 0; c0 < ld; c0 += 32) {
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {  // y: column, x: row (coalesced read)
      const uint64_t r = r0 + threadIdx.x;
      const int c = c0 + y;
      tile[y][threadIdx.x] = (r < rows && c < cols) ? src[(uint64_t)c * rows + r] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {  // y: row, x: column (coalesced write)
      const uint64_t r = r0 + y;
      const int c = c0 + threadIdx.x;
      if (r < rows && c < ld) dst[r * ld + c] = tile[threadIdx.x][y];
    }
    __syncthreads();
  }
}
__global__ void k_rowmajor_to_colmajor(const double* __restrict__ src, int ld, uint64_t rows, int cols,
                                       double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const uint64_t r0 = (uint64_t)blockIdx.x * 32;
  for (int c0 = 0; c0 < cols; c0 += 32) {
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
      const uint64_t r = r0 + y;
      const int c = c0 + threadIdx.x;
      tile[y][threadIdx.x] = (r < rows && c < cols) ? src[r * ld + c] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
      const uint64_t r = r0 + threadIdx.x;
      const int c = c0 + y;
      if (r < rows && c < cols) dst[(uint64_t)c * rows + r] = tile[threadIdx.x][y];
    }
    __syncthreads();
  }
}

__global__ void k_add2(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                       uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

// flip_UV(U, V, ubase=false) (Utils.cpp:136-143) pieces: per column of V the entry of largest
// |value| (first occurrence on ties, like Eigen's maxCoeff) decides the sign.
__global__ void k_colabsmax_partial(const double* __restrict__ V, int ld, int k, uint64_t rows,
                                    uint64_t rows_per_cta, double* __restrict__ pval, double* __restrict__ psgn,
                                    unsigned long long* __restrict__ pidx) {
  __shared__ double sv[256];
  __shared__ unsigned long long si[256];
  const uint64_t r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (int c = 0; c < k; ++c) {
    double best = -1.0;
    unsigned long long bi = ~0ull;
    for (uint64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      const double a = fabs(V[r * ld + c]);
      if (a > best) {
        best = a;
        bi = r;
      }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 1; t < (int)blockDim.x; ++t)
        if (sv[t] > best || (sv[t] == best && si[t] < bi)) {
          best = sv[t];
          bi = si[t];
        }
      pval[(uint64_t)blockIdx.x * k + c] = best;
      pidx[(uint64_t)blockIdx.x * k + c] = bi;
      psgn[(uint64_t)blockIdx.x * k + c] = (bi != ~0ull && V[bi * ld + c] < 0) ? -1.0 : 1.0;
    }
    __syncthreads();
  }
}
__global__ void k_colabsmax_final(const double* __restrict__ pval, const double* __restrict__ psgn,
                                  const unsigned long long* __restrict__ pidx, int nparts, int k,
                                  double* __restrict__ out_val, double* __restrict__ out_sgn) {
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    double best = -1.0, sg = 1.0;
    unsigned long long bi = ~0ull;
    for (int p = 0; p < nparts; ++p) {
      const double v = pval[(uint64_t)p * k + c];
      const unsigned long long i = pidx[(uint64_t)p * k + c];
      if (v > best || (v == best && i < bi)) {
        best = v;
        bi = i;
        sg = psgn[(uint64_t)p * k + c];
      }
    }
    out_val[c] = best;
    out_sgn[c] = sg;
  }
}
// SNP-sharded flip_UV: every rank writes (|v|max, sign) of its shard into ITS slot of a zeroed
// [world][2k] buffer; a sum-allreduce then acts as an all-gather; the winner per column is the
// largest |v|, lowest rank on ties (contiguous shards: lowest rank = lowest SNP index, the
// first-occurrence rule of maxCoeff, Utils.cpp:138-139).
__global__ void k_flip_slot_write(const double* __restrict__ val, const double* __restrict__ sgn, int k, int rank,
                                  int world, double* __restrict__ slots) {
  for (int i = threadIdx.x; i < world * 2 * k; i += blockDim.x) {
    const int r = i / (2 * k), c = i - r * 2 * k;
    slots[i] = (r != rank) ? 0.0 : (c < k ? val[c] : sgn[c - k]);
  }
}
__global__ void k_flip_slot_pick(const double* __restrict__ slots, int k, int world, double* __restrict__ out_sgn) {
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    double best = -1.0, sg = 1.0;
    for (int r = 0; r < world; ++r) {
      const double v = slots[r * 2 * k + c];
      if (v > best) {
        best = v;
        sg = slots[r * 2 * k + k + c];
      }
    }
    out_sgn[c] = sg;
  }
}
__global__ void k_scale_cols(double* __restrict__ A, int ld, int k, uint64_t rows, const double* __restrict__ sign) {
  const uint64_t total = rows * k;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / k;
    const int c = (int)(i - r * k);
    A[r * ld + c] *= sign[c];
  }
}

}  // namespace pcaone
