// pcaone_b200 — EMU update passes on the int8 route: FP64 correction over the missing calls.
//
// Reference: FilePlink.cpp:246-259 (read_block_update, EMU branch) and Data.cpp:334-348 (fit_with_pi):
// on an update pass a missing call (code 01) of sample i at SNP j holds
//     v_ij = s_j * clamp(sum_x U[i][x] S[x] V[j][x], -F_j, 1 - F_j)
// instead of the mean-imputed 0. The block is then X = X0 + E with X0 the mean-imputed block (what the
// tensor-core products of tc_gemm.cuh compute exactly: counts + mask) and E the sparse matrix of the v_ij.
// The two kernels below add the E terms of both products in FP64,
//     G[j][:]  += sum_{i missing at j}  v_ij * Omega[i][:]        (k_emu_fix_g, after k_tc_finish_g)
//     H[i][:]  += sum_{j missing at i}  v_ij * G~[j][:]           (k_emu_fix_h, after k_tc_finish_h)
// straight from the tiled 2-bit operands the GEMMs read (PG: rows = SNPs, PH: rows = samples): one thread owns
// one row of a 128-row tile and walks its 16-byte k-blocks; the missing calls of a k-block are the set bits of
// lo & ~hi. The operands a k-block needs (64 rows of U*S or V, 64 rows of Omega or G~) are staged in shared
// memory once per block and k-block; the thread's own row of V (or U*S) and its output columns live in
// registers, so every output element has exactly one owner: no atomics, sums in a fixed order.
#pragma once
#include "common.cuh"

namespace pcaone {
namespace emu {

constexpr int kLT = 24;        // output columns per thread and column tile (grid.y walks the tiles)
constexpr int kLTP = kLT + 1;  // odd leading dimension of the staged rows (reads hit thread-dependent rows)
constexpr int kThreads = 128;  // = rows of one operand tile
constexpr int kKB = 64;        // entries of one k-block
constexpr int kMaxK = 56;      // largest k the register-resident row supports (kMaxL / 2)

__host__ __device__ constexpr int krp(int KR) { return KR | 1; }
__host__ __device__ constexpr size_t smem_g(int KR) { return (size_t)kKB * (krp(KR) + kLTP) * sizeof(double); }
__host__ __device__ constexpr size_t smem_h(int KR) { return (size_t)kKB * (krp(KR) + kLTP + 3) * sizeof(double); }

// set bits (at even positions 2b) = entries b of the word whose code is 01
__device__ __forceinline__ uint32_t missing_bits(uint32_t w) { return w & 0x55555555u & ~(w >> 1); }

// G rows of the range += E^T Omega. grid.x = row tiles the range touches, grid.y = column tiles.
// V, F, G: row 0 = first SNP of the range. loc0 = index of that SNP inside the tiling PG points at.
template <int KR>
__global__ void __launch_bounds__(kThreads)
k_emu_fix_g(const uint8_t* __restrict__ PG, uint64_t stride_rt, uint32_t nkb, uint32_t N, uint64_t loc0, uint32_t nrows,
            const double* __restrict__ U, int ldu, const double* __restrict__ S, int k, const double* __restrict__ V,
            int ldv, const double* __restrict__ Omg, int lp, int l, const double* __restrict__ F, LutParams lut,
            double* __restrict__ G, unsigned long long* __restrict__ colmax) {
  constexpr int KP = krp(KR);
  extern __shared__ double sm[];
  double* Us = sm;             // [64][KP]   U[i][x] * S[x], zero beyond k and beyond N
  double* Os = sm + kKB * KP;  // [64][kLTP] Omega[i][c0 + cc]
  const int tid = threadIdx.x;
  const uint32_t rt = (uint32_t)(loc0 / kThreads) + blockIdx.x;
  const int c0 = blockIdx.y * kLT;
  const long long jr = (long long)rt * kThreads + tid - (long long)loc0;  // this thread's SNP inside the range
  const bool live = jr >= 0 && jr < (long long)nrows;
  double v[KR], acc[kLT];
  double fj = 0.0, sj = 1.0;
#pragma unroll
  for (int x = 0; x < KR; ++x) v[x] = 0.0;
#pragma unroll
  for (int cc = 0; cc < kLT; ++cc) acc[cc] = 0.0;
  if (live) {
    fj = F[jr];
    sj = snp_scale(fj, lut);
#pragma unroll
    for (int x = 0; x < KR; ++x)
      if (x < k) v[x] = V[(uint64_t)jr * ldv + x];
  }
  const double lo = -fj, hi = 1.0 - fj;
  const uint8_t* prow = PG + (uint64_t)rt * stride_rt + tid * 16;
  uint4 q = make_uint4(0, 0, 0, 0);
  if (live) q = *reinterpret_cast<const uint4*>(prow);
  for (uint32_t kb = 0; kb < nkb; ++kb) {
    __syncthreads();  // the previous k-block's rows are consumed
    const uint32_t s0 = kb * kKB;
    for (int idx = tid; idx < kKB * KP; idx += kThreads) {
      const int i = idx / KP, x = idx - i * KP;
      const uint32_t smp = s0 + i;
      Us[idx] = (x < k && smp < N) ? U[(uint64_t)smp * ldu + x] * S[x] : 0.0;
    }
    for (int idx = tid; idx < kKB * kLT; idx += kThreads) {
      const int i = idx / kLT, cc = idx - i * kLT;
      const uint32_t smp = s0 + i;
      Os[i * kLTP + cc] = (smp < N && c0 + cc < l) ? Omg[(uint64_t)smp * lp + c0 + cc] : 0.0;
    }
    uint4 qn = make_uint4(0, 0, 0, 0);
    if (live && kb + 1 < nkb) qn = *reinterpret_cast<const uint4*>(prow + (uint64_t)(kb + 1) * (kThreads * 16));
    __syncthreads();
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int x4 = 0; x4 < 4; ++x4) {
      uint32_t m = missing_bits(w[x4]);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int i = 16 * x4 + (b >> 1);
        const double* ui = Us + i * KP;
        double f = 0.0;
#pragma unroll
        for (int x = 0; x < KR; ++x) f += ui[x] * v[x];
        const double a = __dmul_rn(fmin(fmax(f, lo), hi), sj);
        const double* oi = Os + i * kLTP;
#pragma unroll
        for (int cc = 0; cc < kLT; ++cc) acc[cc] += a * oi[cc];
      }
    }
    q = qn;
  }
  // write-out + column maxima of W = s o G for the slicing that follows (bounds from above are enough)
  __syncthreads();
  double* red = sm;  // [4 warps][kLT]
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int cc = 0; cc < kLT; ++cc) {
    double mx = 0.0;
    if (live && c0 + cc < l) {
      double* g = G + (uint64_t)jr * lp + c0 + cc;
      const double gn = *g + acc[cc];
      *g = gn;
      mx = fabs(gn * sj);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp * kLT + cc] = mx;
  }
  __syncthreads();
  if (tid < kLT && c0 + tid < l) {
    const double mx = fmax(fmax(red[tid], red[kLT + tid]), fmax(red[2 * kLT + tid], red[3 * kLT + tid]));
    if (mx > 0.0) atomicMax(&colmax[c0 + tid], (unsigned long long)__double_as_longlong(mx));
  }
}

// Hacc (and Hsum) += E G~ for the SNPs of the range. grid.x = sample row tiles, grid.y = column tiles.
// V, F, G: row 0 = first SNP of the range; loc0 = index of that SNP inside the tiling PH points at.
template <int KR>
__global__ void __launch_bounds__(kThreads)
k_emu_fix_h(const uint8_t* __restrict__ PH, uint32_t nrt, uint32_t N, uint64_t loc0, uint32_t nrows,
            const double* __restrict__ U, int ldu, const double* __restrict__ S, int k, const double* __restrict__ V,
            int ldv, const double* __restrict__ G, int lp, int l, const double* __restrict__ F, LutParams lut,
            double* __restrict__ Hacc, double* __restrict__ Hsum) {
  constexpr int KP = krp(KR);
  extern __shared__ double sm[];
  double* Vs = sm;                    // [64][KP]   V[j][x], zero beyond k and outside the range
  double* Gs = Vs + kKB * KP;         // [64][kLTP] G~[j][c0 + cc]
  double* s_lo = Gs + kKB * kLTP;     // [64] -F_j
  double* s_hi = s_lo + kKB;          // [64] 1 - F_j
  double* s_sc = s_hi + kKB;          // [64] s_j, 0 outside the range
  const int tid = threadIdx.x;
  const uint32_t rt = blockIdx.x;
  const int c0 = blockIdx.y * kLT;
  const uint64_t smp = (uint64_t)rt * kThreads + tid;
  const bool live = smp < N;
  double us[KR], acc[kLT];
#pragma unroll
  for (int x = 0; x < KR; ++x) us[x] = (live && x < k) ? U[smp * ldu + x] * S[x] : 0.0;
#pragma unroll
  for (int cc = 0; cc < kLT; ++cc) acc[cc] = 0.0;
  const uint32_t kb0 = (uint32_t)(loc0 / kKB), kb1 = (uint32_t)((loc0 + nrows - 1) / kKB);
  const uint8_t* pcol = PH + (uint64_t)rt * (kThreads * 16) + tid * 16;
  const uint64_t stride_kb = (uint64_t)nrt * (kThreads * 16);
  uint4 q = *reinterpret_cast<const uint4*>(pcol + kb0 * stride_kb);
  for (uint32_t kb = kb0; kb <= kb1; ++kb) {
    __syncthreads();
    const long long j0 = (long long)kb * kKB - (long long)loc0;  // SNP (inside the range) of entry 0 of this k-block
    for (int idx = tid; idx < kKB * KP; idx += kThreads) {
      const int t = idx / KP, x = idx - t * KP;
      const long long j = j0 + t;
      Vs[idx] = (x < k && j >= 0 && j < (long long)nrows) ? V[(uint64_t)j * ldv + x] : 0.0;
    }
    for (int idx = tid; idx < kKB * kLT; idx += kThreads) {
      const int t = idx / kLT, cc = idx - t * kLT;
      const long long j = j0 + t;
      Gs[t * kLTP + cc] = (j >= 0 && j < (long long)nrows && c0 + cc < l) ? G[(uint64_t)j * lp + c0 + cc] : 0.0;
    }
    if (tid < kKB) {
      const long long j = j0 + tid;
      double f = 0.0, s = 0.0;
      if (j >= 0 && j < (long long)nrows) {
        f = F[j];
        s = snp_scale(f, lut);
      }
      s_lo[tid] = -f;
      s_hi[tid] = 1.0 - f;
      s_sc[tid] = s;  // 0 for the entries of a k-block the range does not cover
    }
    uint4 qn = make_uint4(0, 0, 0, 0);
    if (kb < kb1) qn = *reinterpret_cast<const uint4*>(pcol + (uint64_t)(kb + 1) * stride_kb);
    __syncthreads();
    if (live) {
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int x4 = 0; x4 < 4; ++x4) {
        uint32_t m = missing_bits(w[x4]);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int t = 16 * x4 + (b >> 1);
          const double* vt = Vs + t * KP;
          double f = 0.0;
#pragma unroll
          for (int x = 0; x < KR; ++x) f += us[x] * vt[x];
          const double a = __dmul_rn(fmin(fmax(f, s_lo[t]), s_hi[t]), s_sc[t]);
          const double* gt = Gs + t * kLTP;
#pragma unroll
          for (int cc = 0; cc < kLT; ++cc) acc[cc] += a * gt[cc];
        }
      }
    }
    q = qn;
  }
  if (live) {
#pragma unroll
    for (int cc = 0; cc < kLT; ++cc)
      if (c0 + cc < l) {
        const uint64_t idx = smp * lp + c0 + cc;
        Hacc[idx] += acc[cc];
        if (Hsum) Hsum[idx] += acc[cc];
      }
  }
}

}  // namespace emu
}  // namespace pcaone
