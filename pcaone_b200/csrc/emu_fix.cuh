// pcaone_b200 — EMU update passes on the int8 route: FP64 correction over the missing calls.
//
// Reference: FilePlink.cpp:246-259 (read_block_update, EMU branch) and Data.cpp:334-348 (fit_with_pi):
// on an update pass a missing call (code 01) of sample i at SNP j holds
//     v_ij = s_j * clamp(sum_x (U[i][x] S[x]) V[j][x], -F_j, 1 - F_j)
// instead of the mean-imputed 0. The block is then X = X0 + E with X0 the mean-imputed block (what the
// tensor-core products of tc_gemm.cuh compute exactly: counts + mask) and E the sparse matrix of the v_ij.
// The two kernels below add the E terms of both products in FP64,
//     G[j][:]  += sum_{i missing at j}  v_ij * Omega[i][:]        (k_emu_fix_g, after k_tc_finish_g)
//     H[i][:]  += sum_{j missing at i}  v_ij * G~[j][:]           (k_emu_fix_h, after k_tc_finish_h)
// straight from the tiled 2-bit operands the GEMMs read (PG: rows = SNPs, PH: rows = samples): one thread owns
// one row of a 128-row tile and walks its 16-byte k-blocks; the missing calls of a k-block are the set bits of
// lo & ~hi. The operand rows a group of SB k-blocks needs (64 SB rows of U*S or V, and of Omega or G~) are
// staged in shared memory once per block; the thread's own row of V (or U*S) and its output columns live in
// registers, so every output element has exactly one owner: no atomics, sums in a fixed order.
//
// Cost model (measured, configs[3]: 10 % missing, 2.5e9 missing calls, k = 10, l = 20): the kernels are bound by
// shared-memory bandwidth — every missing call reads k + l doubles at a row that differs from lane to lane, so the
// loads conflict (16 random rows on 16 eight-byte banks: ~3 wavefronts each). Hence 16-byte loads (row pitches are
// odd multiples of 16 bytes), register widths KR / LT matched to k / l instead of one padded size, and SB k-blocks
// per barrier so that the lanes of a warp (different rows, different numbers of missing calls) even out:
// 80 -> 56 ms per product at configs[3]. Tried and dropped: eight lanes per row (the operand row of a missing call as
// one contiguous 256-byte shared-memory row, conflict-free, dot product by a 3-step butterfly) — correct, but one
// missing call at a time per quarter-warp is a ~200-cycle dependent chain (load, FMAs, three 64-bit shuffles) with
// 8 warps per SM to hide it: 300 ms per product.
#pragma once
#include "common.cuh"

namespace pcaone {
namespace emu {

constexpr int kThreads = 128;  // = rows of one operand tile
constexpr int kKB = 64;        // entries of one k-block
constexpr int kMaxK = 56;      // largest k the register-resident row supports (kMaxL / 2)
constexpr int kMaxLT = 24;     // output columns per thread and column tile (grid.y walks the tiles)

// leading dimension (doubles) of a staged row of W doubles (W even): an odd number of 16-byte units
__host__ __device__ constexpr int pitch_of(int W) { return ((W / 2) | 1) * 2; }
__host__ __device__ constexpr size_t smem_bytes(int KR, int LT, int SB, bool hside) {
  return (size_t)SB * kKB * (pitch_of(KR) + pitch_of(LT) + (hside ? 3 : 0)) * sizeof(double);
}
// k-blocks per barrier: as many as keep two blocks per SM resident. Measured at configs[3] (115 ms of correction per
// update pass): two k-blocks per barrier with four resident blocks instead of three, 127 ms; ONE loop over a row's
// missing calls per group instead of one per 16-entry word (ncu: 16.7 of 32 lanes active), with the word picked from
// registers by selects, 127 ms as well — the LSU data pipe (63 % of its peak wavefronts in ncu) is the limit, not the
// lane utilisation.
__host__ __device__ constexpr int sb_of(int KR, int LT) {
  return smem_bytes(KR, LT, 4, true) <= 100 * 1024 ? 4 : smem_bytes(KR, LT, 2, true) <= 100 * 1024 ? 2 : 1;
}

// set bits (at even positions 2b) = entries b of the word whose code is 01
__device__ __forceinline__ uint32_t missing_bits(uint32_t w) { return w & 0x55555555u & ~(w >> 1); }

// US[i][x] = U[i][x] * S[x] (the reference's association: (U S) V), row pitch ld, zero for x >= k
__global__ void k_emu_scale_u(const double* __restrict__ U, const double* __restrict__ S, uint64_t N, int k, int ld,
                              double* __restrict__ US) {
  const uint64_t total = N * (uint64_t)ld;
  for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % ld);
    US[e] = x < k ? U[e] * S[x] : 0.0;
  }
}

// rows [r0, r0 + nr) x columns [c0, c0 + W) of a row-major matrix (pitch ld doubles, c0 and ld even) -> dst[nr][P];
// rows outside [lo, hi) and columns >= ncols are zero. 16-byte copies.
template <int W, int P>
__device__ __forceinline__ void stage_rows(double* dst, const double* __restrict__ src, int ld, long long r0, int nr,
                                           long long lo, long long hi, int c0, int ncols, int tid) {
  constexpr int W2 = W / 2;
  for (int idx = tid; idx < nr * W2; idx += kThreads) {
    const int r = idx / W2, c2 = idx - r * W2;
    const long long row = r0 + r;
    const int c = c0 + 2 * c2;
    double2 v = make_double2(0.0, 0.0);
    if (row >= lo && row < hi && c < ncols) {
      v = *reinterpret_cast<const double2*>(src + (uint64_t)row * ld + c);
      if (c + 1 >= ncols) v.y = 0.0;
    }
    *reinterpret_cast<double2*>(dst + r * P + 2 * c2) = v;
  }
}

// one missing call: f = row . us (KR terms), a = clamp(f) * s, acc += a * out_row (LT terms); 16-byte shared loads
template <int KR, int LT>
__device__ __forceinline__ void fill_term(const double* __restrict__ ur, const double (&v)[KR], double lo, double hi, double s,
                                          const double* __restrict__ orow, double (&acc)[LT]) {
  double f = 0.0;
#pragma unroll
  for (int x = 0; x < KR; x += 2) {
    const double2 u = *reinterpret_cast<const double2*>(ur + x);
    f += u.x * v[x];
    f += u.y * v[x + 1];
  }
  const double a = __dmul_rn(fmin(fmax(f, lo), hi), s);
#pragma unroll
  for (int cc = 0; cc < LT; cc += 2) {
    const double2 o = *reinterpret_cast<const double2*>(orow + cc);
    acc[cc] += a * o.x;
    acc[cc + 1] += a * o.y;
  }
}

// G rows of the range += E^T Omega. grid.x = row tiles the range touches, grid.y = column tiles of LT.
// V, F, G: row 0 = first SNP of the range. loc0 = index of that SNP inside the tiling PG points at.
// US: k_emu_scale_u's output (pitch ldu); all pitches even.
template <int KR, int LT, int SB>
__global__ void __launch_bounds__(kThreads)
k_emu_fix_g(const uint8_t* __restrict__ PG, uint64_t stride_rt, uint32_t nkb, uint32_t N, uint64_t loc0, uint32_t nrows,
            const double* __restrict__ US, int ldu, int k, const double* __restrict__ V, int ldv,
            const double* __restrict__ Omg, int lp, int l, const double* __restrict__ F, LutParams lut,
            double* __restrict__ G, unsigned long long* __restrict__ colmax, uint32_t kb_per_split,
            double* __restrict__ part) {
  constexpr int KP = pitch_of(KR), LTP = pitch_of(LT);
  extern __shared__ __align__(16) double sm[];
  double* Us = sm;                  // [SB * 64][KP]
  double* Os = sm + SB * kKB * KP;  // [SB * 64][LTP]
  // blockIdx.z: slice of the sample axis (ranges of a few row tiles would otherwise run on a few SMs); the slices'
  // sums go to part[z][row][col] and k_emu_reduce_g adds them in order
  const uint32_t kb_begin = blockIdx.z * kb_per_split;
  const uint32_t kb_end = min(nkb, kb_begin + kb_per_split);
  const int tid = threadIdx.x;
  const uint32_t rt = (uint32_t)(loc0 / kThreads) + blockIdx.x;
  const int c0 = blockIdx.y * LT;
  const long long jr = (long long)rt * kThreads + tid - (long long)loc0;  // this thread's SNP inside the range
  const bool live = jr >= 0 && jr < (long long)nrows;
  double v[KR], acc[LT];
  double fj = 0.0, sj = 1.0;
#pragma unroll
  for (int x = 0; x < KR; ++x) v[x] = 0.0;
#pragma unroll
  for (int cc = 0; cc < LT; ++cc) acc[cc] = 0.0;
  if (live) {
    fj = F[jr];
    sj = snp_scale(fj, lut);
#pragma unroll
    for (int x = 0; x < KR; ++x)
      if (x < k) v[x] = V[(uint64_t)jr * ldv + x];
  }
  const double lo = -fj, hi = 1.0 - fj;
  const uint8_t* prow = PG + (uint64_t)rt * stride_rt + tid * 16;
  uint4 q[SB], qn[SB];
  auto load_codes = [&](uint32_t kb0, uint4 (&dst)[SB]) {
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      dst[b] = make_uint4(0, 0, 0, 0);
      if (live && kb0 + b < kb_end) dst[b] = *reinterpret_cast<const uint4*>(prow + (uint64_t)(kb0 + b) * (kThreads * 16));
    }
  };
  load_codes(kb_begin, q);
  for (uint32_t kb0 = kb_begin; kb0 < kb_end; kb0 += SB) {
    __syncthreads();  // the previous group's rows are consumed
    const long long s0 = (long long)kb0 * kKB;
    stage_rows<KR, KP>(Us, US, ldu, s0, SB * kKB, 0, (long long)N, 0, k, tid);
    stage_rows<LT, LTP>(Os, Omg, lp, s0, SB * kKB, 0, (long long)N, c0, l, tid);
    load_codes(kb0 + SB, qn);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      const uint32_t w[4] = {q[b].x, q[b].y, q[b].z, q[b].w};
#pragma unroll
      for (int x4 = 0; x4 < 4; ++x4) {
        uint32_t m = missing_bits(w[x4]);
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int i = b * kKB + 16 * x4 + (bit >> 1);
          fill_term<KR, LT>(Us + i * KP, v, lo, hi, sj, Os + i * LTP, acc);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < SB; ++b) q[b] = qn[b];
  }
  if (part) {  // sliced sample axis: partial sums only
    if (live) {
      double* o = part + ((uint64_t)blockIdx.z * nrows + (uint64_t)jr) * lp + c0;
#pragma unroll
      for (int cc = 0; cc < LT; ++cc)
        if (c0 + cc < l) o[cc] = acc[cc];
    }
    return;
  }
  // write-out + column maxima of W = s o G for the slicing that follows (bounds from above are enough)
  __syncthreads();
  double* red = sm;  // [4 warps][LT]
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int cc = 0; cc < LT; ++cc) {
    double mx = 0.0;
    if (live && c0 + cc < l) {
      double* g = G + (uint64_t)jr * lp + c0 + cc;
      const double gn = *g + acc[cc];
      *g = gn;
      mx = fabs(gn * sj);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp * LT + cc] = mx;
  }
  __syncthreads();
  if (tid < LT && c0 + tid < l) {
    const double mx = fmax(fmax(red[tid], red[LT + tid]), fmax(red[2 * LT + tid], red[3 * LT + tid]));
    if (mx > 0.0) atomicMax(&colmax[c0 + tid], (unsigned long long)__double_as_longlong(mx));
  }
}

// G[row][c] += sum_z part[z][row][c] (z in order) and the column maxima of W = s o G; one thread per column pair of a row
__global__ void __launch_bounds__(256) k_emu_reduce_g(const double* __restrict__ part, uint32_t nsplit, uint32_t nrows, int lp,
                                                      int l, const double* __restrict__ F, LutParams lut,
                                                      double* __restrict__ G, unsigned long long* __restrict__ colmax) {
  __shared__ double s_m[256];
  const int rpp = 256 / lp;                       // rows per pass of the block (lp <= 128)
  const int g0 = threadIdx.x / lp, c = threadIdx.x - g0 * lp;
  double m = 0.0;
  if (g0 < rpp && c < l) {
    for (uint64_t row = (uint64_t)blockIdx.x * rpp + g0; row < nrows; row += (uint64_t)gridDim.x * rpp) {
      double a = 0.0;
      for (uint32_t z = 0; z < nsplit; ++z) a += part[((uint64_t)z * nrows + row) * lp + c];
      const double gn = G[row * lp + c] + a;
      G[row * lp + c] = gn;
      m = fmax(m, fabs(gn * snp_scale(F[row], lut)));
    }
  }
  s_m[threadIdx.x] = m;
  __syncthreads();
  if (g0 == 0 && c < l) {
    double v = 0.0;
    for (int q = 0; q < rpp; ++q) v = fmax(v, s_m[q * lp + c]);
    if (v > 0.0) atomicMax(&colmax[c], (unsigned long long)__double_as_longlong(v));
  }
}

// Hacc (and Hsum) += E G~ for the SNPs of the range. grid.x = sample row tiles, grid.y = column tiles of LT.
// V, F, G: row 0 = first SNP of the range; loc0 = index of that SNP inside the tiling PH points at.
template <int KR, int LT, int SB>
__global__ void __launch_bounds__(kThreads)
k_emu_fix_h(const uint8_t* __restrict__ PH, uint32_t nrt, uint32_t N, uint64_t loc0, uint32_t nrows,
            const double* __restrict__ US, int ldu, int k, const double* __restrict__ V, int ldv,
            const double* __restrict__ G, int lp, int l, const double* __restrict__ F, LutParams lut,
            double* __restrict__ Hacc, double* __restrict__ Hsum) {
  constexpr int KP = pitch_of(KR), LTP = pitch_of(LT);
  extern __shared__ __align__(16) double sm[];
  double* Vs = sm;                     // [SB * 64][KP]   V[j][x], zero outside the range
  double* Gs = Vs + SB * kKB * KP;     // [SB * 64][LTP]  G~[j][c0 + cc]
  double* s_lo = Gs + SB * kKB * LTP;  // [SB * 64] -F_j
  double* s_hi = s_lo + SB * kKB;      // [SB * 64] 1 - F_j
  double* s_sc = s_hi + SB * kKB;      // [SB * 64] s_j, 0 outside the range
  const int tid = threadIdx.x;
  const uint32_t rt = blockIdx.x;
  const int c0 = blockIdx.y * LT;
  const uint64_t smp = (uint64_t)rt * kThreads + tid;
  const bool live = smp < N;
  double us[KR], acc[LT];
#pragma unroll
  for (int x = 0; x < KR; ++x) us[x] = (live && x < k) ? US[smp * ldu + x] : 0.0;
#pragma unroll
  for (int cc = 0; cc < LT; ++cc) acc[cc] = 0.0;
  const uint32_t kbA = (uint32_t)(loc0 / kKB), kbB = (uint32_t)((loc0 + nrows - 1) / kKB);  // first / last k-block
  const uint8_t* pcol = PH + (uint64_t)rt * (kThreads * 16) + tid * 16;
  const uint64_t stride_kb = (uint64_t)nrt * (kThreads * 16);
  uint4 q[SB], qn[SB];
  auto load_codes = [&](uint32_t kb0, uint4 (&dst)[SB]) {
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      dst[b] = make_uint4(0, 0, 0, 0);
      if (live && kb0 + b <= kbB) dst[b] = *reinterpret_cast<const uint4*>(pcol + (uint64_t)(kb0 + b) * stride_kb);
    }
  };
  load_codes(kbA, q);
  for (uint32_t kb0 = kbA; kb0 <= kbB; kb0 += SB) {
    __syncthreads();
    const long long j0 = (long long)kb0 * kKB - (long long)loc0;  // SNP (inside the range) of entry 0 of this group
    stage_rows<KR, KP>(Vs, V, ldv, j0, SB * kKB, 0, (long long)nrows, 0, k, tid);
    stage_rows<LT, LTP>(Gs, G, lp, j0, SB * kKB, 0, (long long)nrows, c0, l, tid);
    for (int t = tid; t < SB * kKB; t += kThreads) {
      const long long j = j0 + t;
      double f = 0.0, s = 0.0;
      if (j >= 0 && j < (long long)nrows) {
        f = F[j];
        s = snp_scale(f, lut);
      }
      s_lo[t] = -f;
      s_hi[t] = 1.0 - f;
      s_sc[t] = s;  // 0 for the entries of a k-block the range does not cover: their bits belong to other blocks
    }
    load_codes(kb0 + SB, qn);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      const uint32_t w[4] = {q[b].x, q[b].y, q[b].z, q[b].w};
#pragma unroll
      for (int x4 = 0; x4 < 4; ++x4) {
        uint32_t m = missing_bits(w[x4]);
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int t = b * kKB + 16 * x4 + (bit >> 1);
          fill_term<KR, LT>(Vs + t * KP, us, s_lo[t], s_hi[t], s_sc[t], Gs + t * LTP, acc);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < SB; ++b) q[b] = qn[b];
  }
  if (live) {
#pragma unroll
    for (int cc = 0; cc < LT; ++cc)
      if (c0 + cc < l) {
        const uint64_t idx = smp * lp + c0 + cc;
        Hacc[idx] += acc[cc];
        if (Hsum) Hsum[idx] += acc[cc];
      }
  }
}

}  // namespace emu
}  // namespace pcaone
