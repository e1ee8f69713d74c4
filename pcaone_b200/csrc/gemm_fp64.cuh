// pcaone_b200 — fused 2-bit decode -> FP64 tensor-core (DMMA m8n8k4) GEMMs.
//
//   k_gemm_g :  G_b = X_b^T * Omega     (reference: Halko.cpp:125,150,195-196,243)
//   k_gemm_h :  Hpart = X_b * G_b       (reference: Halko.cpp:126,151,200-202,246-248)
//
// X_b (N x B doubles in the reference, FilePlink.cpp:139-162) is never materialised: the packed
// 2-bit rows are expanded through the per-SNP 4-entry table (common.cuh make_lut) straight into
// the A fragments of the MMAs. Tall operands live in HBM as row-major [rows][ld] doubles
// (ld = 8*NT), so one operand row is one contiguous, 64-byte aligned run.
//
// Both kernels: 256 threads = 8 warps, each warp owns 16 rows of the output (2 m8 tiles) and
// all 8*NT columns; the B operand chunk (64 x 8*NT doubles) is staged in shared memory with
// cp.async, double-buffered, and shared by the 8 warps.
#pragma once
#include "common.cuh"

namespace pcaone {

constexpr int kGemmThreads = 256;
constexpr int kTileRows = 128;  // output rows per CTA (8 warps x 16)
constexpr int kKC = 64;         // contraction chunk per pipeline stage

template <int NT>
struct GemmSmem {
  static constexpr int LP = NT * 8;
  static constexpr int LDB = smem_ld(LP);
  static constexpr size_t kBBytes = (size_t)kKC * LDB * sizeof(double);
  // gemm_g: B only. gemm_h: B + packed tile (64 x 32 B) + LUT (64 x 4 doubles)
  static constexpr size_t kStageG = kBBytes;
  static constexpr size_t kStageH = kBBytes + kKC * 32 + kKC * 4 * sizeof(double);
};

__device__ __forceinline__ double lut_select(const double (&v)[4], uint32_t code) {
  const double lo = (code & 1u) ? v[1] : v[0];
  const double hi = (code & 1u) ? v[3] : v[2];
  return (code & 2u) ? hi : lo;
}

// ------------------------------------------------------------------------------------------
// G[row][c] = sum_i X[i][row] * Omega[i][c]          rows = SNPs of the range, i = samples
//   A fragment: row = SNP (lane>>2), k = sample (lane&3)  -> one packed byte = one k4 step
//   B fragment: Omega_s[k0 + (lane&3)][8n + (lane>>2)]
// EMU: entries with code 01 take clamp(U_i . (S*V_row)) * s instead of 0 (FilePlink.cpp:252-259)
// ------------------------------------------------------------------------------------------
template <int NT, bool EMU>
__global__ void __launch_bounds__(kGemmThreads, (NT <= 10 ? 2 : 1))
k_gemm_g(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t nrows, uint32_t N,
         const double* __restrict__ Omg,  // [N][8NT]
         const double* __restrict__ F, LutParams lp, double* __restrict__ G,  // [nrows][8NT]
         const double* __restrict__ U, int ldu, const double* __restrict__ S, const double* __restrict__ V,
         int ldv, int kk) {
  using SM = GemmSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Bs[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + SM::kStageG)};

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t row_base = blockIdx.x * kTileRows + warp * 16;
  uint32_t rows[2];
  const uint8_t* prow[2];
  double lut[2][4], fj[2], sj[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    rows[u] = row_base + 8 * u + g;
    const uint32_t rc = rows[u] < nrows ? rows[u] : nrows - 1;  // clamp: loads stay in range
    prow[u] = P + (uint64_t)rc * pitch;
    fj[u] = F[rc];
    const SnpLut tt = make_lut(fj[u], lp);
    sj[u] = snp_scale(fj[u], lp);
#pragma unroll
    for (int c = 0; c < 4; ++c) lut[u][c] = tt.v[c];
  }

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  const int nchunks = (int)((N + kKC - 1) / kKC);

  auto load_B = [&](int chunk, int stage) {
    // 64 rows x LP doubles, 16-byte pieces; rows >= N are zero-filled (src-size 0)
    constexpr int PIECES = LP / 2;
    for (int idx = tid; idx < kKC * PIECES; idx += kGemmThreads) {
      const int r = idx / PIECES, pc = idx - r * PIECES;
      const uint32_t i = (uint32_t)chunk * kKC + r;
      const bool ok = i < N;
      const double* src = Omg + (uint64_t)(ok ? i : 0) * LP + pc * 2;
      cp_async16(Bs[stage] + r * LDB + pc * 2, src, ok ? 16 : 0);
    }
  };

  uint32_t wcur[2], wnext[2] = {0u, 0u};
  load_B(0, 0);
  cp_async_commit();
#pragma unroll
  for (int u = 0; u < 2; ++u) wcur[u] = __ldg(reinterpret_cast<const uint32_t*>(prow[u]) + t);

  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) {
      load_B(c + 1, st ^ 1);
#pragma unroll
      for (int u = 0; u < 2; ++u)
        wnext[u] = __ldg(reinterpret_cast<const uint32_t*>(prow[u]) + (c + 1) * 4 + t);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* B = Bs[st];
#pragma unroll 4
    for (int s = 0; s < 16; ++s) {
      const int w = s >> 2, q = s & 3;
      // the 4 lanes of a row hold the 4 words of this 64-sample chunk; fetch word w
      const uint32_t v0 = __shfl_sync(0xffffffffu, wcur[0], (lane & ~3) | w);
      const uint32_t v1 = __shfl_sync(0xffffffffu, wcur[1], (lane & ~3) | w);
      const uint32_t c0 = (v0 >> (8 * q + 2 * t)) & 3u;
      const uint32_t c1 = (v1 >> (8 * q + 2 * t)) & 3u;
      double a0 = lut_select(lut[0], c0);
      double a1 = lut_select(lut[1], c1);
      if (EMU) {
        const uint32_t i = (uint32_t)c * kKC + 4 * s + t;
        if (i < N) {
          if (c0 == 1u) {
            double f = 0.0;
            for (int x = 0; x < kk; ++x)
              f += (U[(uint64_t)i * ldu + x] * S[x]) * V[(uint64_t)(rows[0] < nrows ? rows[0] : nrows - 1) * ldv + x];
            a0 = __dmul_rn(fmin(fmax(f, -fj[0]), 1.0 - fj[0]), sj[0]);
          }
          if (c1 == 1u) {
            double f = 0.0;
            for (int x = 0; x < kk; ++x)
              f += (U[(uint64_t)i * ldu + x] * S[x]) * V[(uint64_t)(rows[1] < nrows ? rows[1] : nrows - 1) * ldv + x];
            a1 = __dmul_rn(fmin(fmax(f, -fj[1]), 1.0 - fj[1]), sj[1]);
          }
        }
      }
      const double* brow = B + (4 * s + t) * LDB + g;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = brow[8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
    wcur[0] = wnext[0];
    wcur[1] = wnext[1];
  }

#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (rows[u] < nrows) {
      double* out = G + (uint64_t)rows[u] * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(out + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Hpart[split][i][c] = sum_{j in split} X[i][j] * G[j][c]     i = samples, j = SNPs of the range
//   A fragment: row = sample (lane>>2), k = SNP (lane&3)
//   B fragment: G_s[j0 + (lane&3)][8n + (lane>>2)]
// grid.x = sample tiles of 128, grid.y = SNP splits (each a multiple of 64 SNPs)
// ------------------------------------------------------------------------------------------
template <int NT, bool EMU>
__global__ void __launch_bounds__(kGemmThreads, (NT <= 10 ? 2 : 1))
k_gemm_h(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t nrows, uint32_t N,
         const double* __restrict__ G,  // [nrows][8NT]
         const double* __restrict__ F, LutParams lp, double* __restrict__ Hpart,  // [splits][N][8NT]
         uint32_t rows_per_split, const double* __restrict__ U, int ldu, const double* __restrict__ S,
         const double* __restrict__ V, int ldv, int kk) {
  using SM = GemmSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Bs[2];
  uint32_t* Ps[2];
  double* Ls[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    unsigned char* base = smem_raw + s * SM::kStageH;
    Bs[s] = reinterpret_cast<double*>(base);
    Ls[s] = reinterpret_cast<double*>(base + SM::kBBytes);
    Ps[s] = reinterpret_cast<uint32_t*>(base + SM::kBBytes + kKC * 4 * sizeof(double));
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t samp_tile = blockIdx.x * kTileRows;
  const uint32_t j_begin = blockIdx.y * rows_per_split;
  const uint32_t j_end = min(nrows, j_begin + rows_per_split);
  const uint32_t tile_byte = samp_tile >> 2;  // byte offset of this sample tile in a packed row
  const uint32_t samp[2] = {samp_tile + warp * 16 + g, samp_tile + warp * 16 + 8 + g};

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  const int nchunks = j_end > j_begin ? (int)((j_end - j_begin + kKC - 1) / kKC) : 0;

  auto load_stage = [&](int chunk, int stage) {
    const uint32_t j0 = j_begin + (uint32_t)chunk * kKC;
    constexpr int PIECES = LP / 2;
    for (int idx = tid; idx < kKC * PIECES; idx += kGemmThreads) {
      const int r = idx / PIECES, pc = idx - r * PIECES;
      const uint32_t j = j0 + r;
      const bool ok = j < j_end;
      cp_async16(Bs[stage] + r * LDB + pc * 2, G + (uint64_t)(ok ? j : j_begin) * LP + pc * 2, ok ? 16 : 0);
    }
    if (tid < kKC * 2) {  // packed tile: 64 SNP rows x 32 bytes (128 samples)
      const int r = tid >> 1, h = tid & 1;
      const uint32_t j = j0 + r;
      const bool ok = (j < j_end) && (tile_byte + 16 * h < pitch);
      cp_async16(reinterpret_cast<unsigned char*>(Ps[stage]) + r * 32 + h * 16,
                 P + (uint64_t)(ok ? j : j_begin) * pitch + (ok ? tile_byte + 16 * h : 0), ok ? 16 : 0);
    } else if (tid < kKC * 3) {  // per-SNP decode table
      const int r = tid - kKC * 2;
      const uint32_t j = j0 + r;
      SnpLut tt;
      if (j < j_end) {
        tt = make_lut(F[j], lp);
      } else {
        tt.v[0] = tt.v[1] = tt.v[2] = tt.v[3] = 0.0;
      }
      double* d = Ls[stage] + r * 4;
      d[0] = tt.v[0];
      d[1] = tt.v[1];
      d[2] = tt.v[2];
      d[3] = tt.v[3];
    }
  };

  if (nchunks > 0) load_stage(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) load_stage(c + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* B = Bs[st];
    const double* L = Ls[st];
    const uint32_t* PP = Ps[st];
#pragma unroll 4
    for (int s = 0; s < 16; ++s) {
      const int jl = 4 * s + t;
      const uint32_t w = PP[jl * 8 + warp];  // this warp's 16 samples of SNP jl
      const uint32_t c0 = (w >> (2 * g)) & 3u;
      const uint32_t c1 = (w >> (16 + 2 * g)) & 3u;
      double a0 = L[jl * 4 + c0];
      double a1 = L[jl * 4 + c1];
      if (EMU) {
        const uint32_t j = j_begin + (uint32_t)c * kKC + jl;
        if (j < j_end) {
          const double f = F[j];
          const double sc = snp_scale(f, lp);
          if (c0 == 1u && samp[0] < N) {
            double x = 0.0;
            for (int q = 0; q < kk; ++q) x += (U[(uint64_t)samp[0] * ldu + q] * S[q]) * V[(uint64_t)j * ldv + q];
            a0 = __dmul_rn(fmin(fmax(x, -f), 1.0 - f), sc);
          }
          if (c1 == 1u && samp[1] < N) {
            double x = 0.0;
            for (int q = 0; q < kk; ++q) x += (U[(uint64_t)samp[1] * ldu + q] * S[q]) * V[(uint64_t)j * ldv + q];
            a1 = __dmul_rn(fmin(fmax(x, -f), 1.0 - f), sc);
          }
        }
      }
      const double* brow = B + jl * LDB + g;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = brow[8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
  }

  double* out = Hpart + (uint64_t)blockIdx.y * N * LP;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (samp[u] < N) {
      double* o = out + (uint64_t)samp[u] * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(o + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}

// Hacc[i] = (accumulate ? Hacc[i] : 0) + sum_s part[s][i]   (fixed order -> deterministic)
__global__ void k_reduce_partials(const double* __restrict__ part, uint32_t splits, uint64_t count,
                                  double* __restrict__ Hacc, int accumulate) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    double v = accumulate ? Hacc[i] : 0.0;
    for (uint32_t s = 0; s < splits; ++s) v += part[(uint64_t)s * count + i];
    Hacc[i] = v;
  }
}

}  // namespace pcaone
