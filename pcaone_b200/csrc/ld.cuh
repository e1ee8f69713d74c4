// pcaone_b200 — LD r^2 on the FP64 tensor cores: banded tile Gram of column-centred genotypes.
//
//   reference: calc_sds (LD.cpp:48-51), ld_r2_big (LD.cpp:450-473), windows from
//   divide_pos_by_window (LD.cpp:154-168).
//
// The reference walks every (lead SNP i, partner k in its bp-window) pair with a BLAS-1 dot
// product. Here the SNP-major matrix Gs [M][Np] (Np = N rounded up to 16, zero padded; this is
// Eigen's column-major N x M with every column 128-byte aligned) is cut into 128 x 128 tiles of
// the Gram G^T G; only the tiles the windows touch are computed (host tile list), each by one
// CTA with DMMA m8n8k4, and the epilogue writes r^2 = (dot * (1/sd_i)(1/sd_k)/(N-1))^2 straight
// to the reference's output order (window after window, partners ascending).
#pragma once
#include "common.cuh"

namespace pcaone {
namespace ld {

constexpr int kTile = 128;     // SNPs per tile side
constexpr int kKC = 16;        // samples per pipeline stage
constexpr int kLD = 20;        // smem row stride in doubles (== 4 mod 16: conflict-free fragments)
constexpr int kStages = 4;
constexpr int kThreads = 256;  // 8 warps: 4 (rows) x 2 (cols), warp tile 32 x 64
constexpr size_t kStageDoubles = 2 * kTile * kLD;
constexpr size_t kSmemBytes = kStages * kStageDoubles * sizeof(double);

// 1/sd per SNP: sd = sqrt(sum g^2 * df)   (calc_sds, then `1.0 / calc_sds(G)` LD.cpp:454)
__global__ void __launch_bounds__(256) k_inv_sd(const double* __restrict__ Gs, uint64_t rows, uint32_t Np, double df,
                                                 double* __restrict__ inv_sd) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t j = warp; j < rows; j += nwarps) {
    const double2* row = reinterpret_cast<const double2*>(Gs + j * Np);
    double s = 0.0;
    for (uint32_t i = lane; i < Np / 2; i += 32) {
      const double2 v = row[i];
      s += v.x * v.x + v.y * v.y;
    }
    s = warp_sum(s);
    if (lane == 0) inv_sd[j] = 1.0 / sqrt(s * df);
  }
}

// host column-major block (rows SNPs x N samples, SNP contiguous) -> padded [rows][Np]
__global__ void k_pad_rows(const double* __restrict__ src, uint64_t rows, uint32_t N, uint32_t Np,
                           double* __restrict__ dst) {
  const uint64_t total = rows * Np;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / Np;
    const uint32_t i = (uint32_t)(idx - r * Np);
    dst[idx] = i < N ? src[r * N + i] : 0.0;
  }
}

// packed 2-bit rows -> padded centred doubles [rows][Np] (read_block_initial with
// standardize = false, FilePlink.cpp:139-162: code 01 -> 0, else BED2GENO - F)
__global__ void k_decode_rows(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t N, uint32_t Np, uint64_t rows,
                              const double* __restrict__ F, LutParams lp, double* __restrict__ dst) {
  const uint32_t nq = Np >> 2;
  const uint64_t total = rows * nq;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = idx / nq;
    const uint32_t q = (uint32_t)(idx - j * nq);
    const SnpLut t = make_lut(F[j], lp);
    const uint32_t byte = (q < pitch) ? P[j * pitch + q] : 0u;
    double v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = (4 * q + r < N) ? t.v[(byte >> (2 * r)) & 3u] : 0.0;
    double2* o = reinterpret_cast<double2*>(dst + j * Np + 4 * q);
    o[0] = make_double2(v[0], v[1]);
    o[1] = make_double2(v[2], v[3]);
  }
}

// ---- residual operands (Data::write_residuals, Data.cpp:242-291; FileBin::read_all, FileBinary.cpp:21-30;
// run_ld_stuff's (I - U U^T) G, LD.cpp:491-496) ---------------------------------------------------------
// dst[j][i] = centred unscaled genotype (code 01 -> 0) - sum_k (U[i][k] S[k]) V[j][k]: the `G -= U * S * V^T`
// of write_residuals with ld_stats = 0 on a chunk of SNP rows. CTA = 32 SNPs x 128 samples; the U tile and
// the S-scaled V rows sit in shared memory. nk == 0: plain centred decode.
constexpr int kResSnps = 32, kResSamples = 128;
__global__ void __launch_bounds__(256) k_resid_rows(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t N, uint32_t Np,
                                                     uint64_t rows, const double* __restrict__ F, LutParams lp,
                                                     const double* __restrict__ U, int ldu, const double* __restrict__ S,
                                                     const double* __restrict__ V, int ldv, int nk,
                                                     double* __restrict__ dst) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Us = reinterpret_cast<double*>(smem_raw);      // [128][nk + 1]
  double* Vs = Us + (size_t)kResSamples * (nk + 1);      // [32][nk]  (V[j][k] * S[k])
  const int tid = threadIdx.x;
  const uint32_t ntile_s = (Np + kResSamples - 1) / kResSamples;
  const uint64_t j0 = (uint64_t)(blockIdx.x / ntile_s) * kResSnps;
  const uint32_t i0 = (blockIdx.x % ntile_s) * kResSamples;
  for (int idx = tid; idx < kResSamples * nk; idx += 256) {
    const int r = idx / nk, k = idx - r * nk;
    Us[r * (nk + 1) + k] = (i0 + r < N) ? U[(uint64_t)(i0 + r) * ldu + k] : 0.0;
  }
  for (int idx = tid; idx < kResSnps * nk; idx += 256) {
    const int r = idx / nk, k = idx - r * nk;
    Vs[r * nk + k] = (j0 + r < rows) ? V[(j0 + r) * ldv + k] * S[k] : 0.0;
  }
  __syncthreads();
  const int si = tid & (kResSamples - 1), half = tid >> 7;   // sample of this thread, which 16 SNPs
  const uint32_t i = i0 + si;
  if (i >= Np) return;
  const double* u = Us + si * (nk + 1);
#pragma unroll 1
  for (int r = half * 16; r < half * 16 + 16; ++r) {
    const uint64_t j = j0 + r;
    if (j >= rows) break;
    double x = 0.0;
    if (i < N) {
      const SnpLut t = make_lut(F[j], lp);
      const uint32_t byte = P[j * pitch + (i >> 2)];
      x = t.v[(byte >> (2 * (i & 3))) & 3u];
      const double* v = Vs + r * nk;
      double acc = 0.0;
      for (int k = 0; k < nk; ++k) acc = fma(u[k], v[k], acc);
      x -= acc;
    }
    dst[j * Np + i] = x;
  }
}

// Per SNP row: subtract the mean over the N samples (`G.rowwise() -= G.colwise().mean()`), optionally
// round through float32 the way the .residuals file does (write_residuals casts, FileBin::read_all reads
// back and centres AGAIN), and optionally hand the floats out ([rows][N], the file's row layout).
// One warp per row.
__global__ void __launch_bounds__(256) k_center_rows(double* __restrict__ Gs, uint64_t rows, uint32_t N, uint32_t Np,
                                                      int through_f32, float* __restrict__ f32_out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t j = warp; j < rows; j += nwarps) {
    double* row = Gs + j * Np;
    double s = 0.0;
    for (uint32_t i = lane; i < N; i += 32) s += row[i];
    s = warp_sum(s);
    const double mean = s / (double)N;
    if (!through_f32) {
      for (uint32_t i = lane; i < N; i += 32) row[i] -= mean;
      continue;
    }
    double s2 = 0.0;
    for (uint32_t i = lane; i < N; i += 32) {
      const float f = (float)(row[i] - mean);
      if (f32_out) f32_out[j * N + i] = f;
      row[i] = (double)f;
      s2 += (double)f;
    }
    s2 = warp_sum(s2);
    const double mean2 = s2 / (double)N;
    for (uint32_t i = lane; i < N; i += 32) row[i] -= mean2;
  }
}

// float32 rows of a .residuals file ([rows][N]) -> padded doubles [rows][Np] (centred afterwards by k_center_rows)
__global__ void k_pad_rows_f32(const float* __restrict__ src, uint64_t rows, uint32_t N, uint32_t Np,
                               double* __restrict__ dst) {
  const uint64_t total = rows * Np;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / Np;
    const uint32_t i = (uint32_t)(idx - r * Np);
    dst[idx] = i < N ? (double)src[r * N + i] : 0.0;
  }
}

// (I - U U^T) g for every SNP row g (LD.cpp:494): one warp per row, two sweeps over the samples
// (t = U^T g in registers / shared memory, then g -= U t). U: [N][ldu] row-major, nk <= 64.
__global__ void __launch_bounds__(256) k_project_out(double* __restrict__ Gs, uint64_t rows, uint32_t N, uint32_t Np,
                                                      const double* __restrict__ U, int ldu, int nk) {
  __shared__ double s_t[8][64];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t j = warp; j < rows; j += nwarps) {
    double* row = Gs + j * Np;
    for (int k0 = 0; k0 < nk; k0 += 8) {   // eight projections at a time
      double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (uint32_t i = lane; i < N; i += 32) {
        const double g = row[i];
        const double* u = U + (uint64_t)i * ldu + k0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (k0 + q < nk) t[q] = fma(u[q], g, t[q]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        t[q] = warp_sum(t[q]);
        if (lane == 0 && k0 + q < nk) s_t[wib][k0 + q] = t[q];
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < N; i += 32) {
      const double* u = U + (uint64_t)i * ldu;
      double acc = 0.0;
      for (int k = 0; k < nk; ++k) acc = fma(u[k], s_t[wib][k], acc);
      row[i] -= acc;
    }
    __syncwarp();
  }
}

struct LdArgs {
  const double* Gs;         // [rows][Np], chunk-local row 0 = SNP snp0
  uint32_t Np;              // padded sample count (multiple of 16)
  uint64_t rows;            // rows present in Gs
  const double* inv_sd;     // [rows]
  double df;                // 1 / (N - 1)
  const int2* tiles;        // (lead tile, partner tile), chunk-local tile indices
  const int32_t* win_of;    // [rows] window index of a lead SNP (chunk-local row), -1 = none
  const int32_t* we;        // [nwin] SNPs per window incl. the lead
  const uint64_t* offs;     // [nwin] output offset of the window's first pair
  uint64_t out0;            // offset of the chunk's first output value
  double* out;              // chunk output
};

__global__ void __launch_bounds__(kThreads, 1) k_ld_tiles(const LdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;  // warp tile: rows 32*wr.., cols 64*wc..
  const int2 tile = a.tiles[blockIdx.x];
  const uint64_t i0 = (uint64_t)tile.x * kTile, k0 = (uint64_t)tile.y * kTile;

  double acc[4][8][2];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

  // stage = A tile (128 SNP rows x 16 samples) + B tile; 8 x 16-byte pieces per row
  auto load_stage = [&](int chunk, int stage) {
    double* As = sm + (size_t)stage * kStageDoubles;
    double* Bs = As + kTile * kLD;
    for (int idx = tid; idx < 2 * kTile * 8; idx += kThreads) {
      const int which = idx >> 10, rem = idx & 1023;
      const int r = rem >> 3, pc = rem & 7;
      const uint64_t row = (which ? k0 : i0) + r;
      const bool ok = row < a.rows;
      const double* src = a.Gs + (ok ? row : 0) * a.Np + (uint64_t)chunk * kKC + pc * 2;
      cp_async16((which ? Bs : As) + r * kLD + pc * 2, src, ok ? 16 : 0);
    }
  };

  const int nchunks = (int)(a.Np / kKC);
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    if (c + kStages - 1 < nchunks) load_stage(c + kStages - 1, (c + kStages - 1) % kStages);
    cp_async_commit();
    const double* As = sm + (size_t)(c % kStages) * kStageDoubles + (size_t)(32 * wr + g) * kLD + t;
    const double* Bs = sm + (size_t)(c % kStages) * kStageDoubles + kTile * kLD + (size_t)(64 * wc + g) * kLD + t;
#pragma unroll
    for (int ks = 0; ks < kKC / 4; ++ks) {
      double af[4], bf[8];
#pragma unroll
      for (int m = 0; m < 4; ++m) af[m] = As[m * 8 * kLD + ks * 4];
#pragma unroll
      for (int n = 0; n < 8; ++n) bf[n] = Bs[n * 8 * kLD + ks * 4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) dmma884(acc[m][n][0], acc[m][n][1], af[m], bf[n]);
    }
  }

  // epilogue: element (i, k) of the Gram -> r^2 at offs[w(i)] + (k - i - 1) when k is in i's window
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint64_t i = i0 + 32 * wr + 8 * m + g;
    if (i >= a.rows) continue;
    const int w = a.win_of[i];
    if (w < 0) continue;
    const int64_t n_in = a.we[w];
    const double si = a.inv_sd[i];
    double* orow = a.out + (a.offs[w] - a.out0);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t k = k0 + 64 * wc + 8 * n + 2 * t + h;
        const int64_t d = (int64_t)k - (int64_t)i;
        if (d >= 1 && d < n_in && k < a.rows) {
          const double r = acc[m][n][h] * (si * a.inv_sd[k] * a.df);
          orow[d - 1] = r * r;
        }
      }
    }
  }
}


// Greedy LD pruning over the r^2 values of one chunk (reference ld_prune_big, LD.cpp:252-266):
// windows in ascending lead order; a window is skipped when its lead was pruned before; inside a
// window every partner k still kept with r^2 > tol prunes o = k (no allele frequencies) or the
// one of {lead, k} with the smaller MAF (`MAF(af[k]) > MAF(af[i]) ? i : k`). The windows are a
// serial chain through keep[], the partners of one window are independent (the reference runs them
// as an OpenMP parallel for): ONE CTA walks the windows, its threads take the partners.
__global__ void __launch_bounds__(1024) k_ld_prune(const double* __restrict__ r2, uint64_t out0,
                                                    const uint64_t* __restrict__ offs, const int32_t* __restrict__ ws,
                                                    const int32_t* __restrict__ we, uint64_t w_lo, uint64_t w_hi,
                                                    const double* __restrict__ af, double r2_tol,
                                                    unsigned char* keep) {
  __shared__ int s_lead_kept;
  for (uint64_t w = w_lo; w < w_hi; ++w) {
    const int i = ws[w], n = we[w];
    if (threadIdx.x == 0) s_lead_kept = ((volatile unsigned char*)keep)[i];
    __syncthreads();
    const int kept = s_lead_kept;
    if (kept) {
      const double* rw = r2 + (offs[w] - out0);
      const double mi = af ? (af[i] > 0.5 ? 1.0 - af[i] : af[i]) : 0.0;
      for (int j = 1 + (int)threadIdx.x; j < n; j += (int)blockDim.x) {
        const int k = i + j;
        const double v = rw[j - 1];
        if (((volatile unsigned char*)keep)[k] && v > r2_tol) {
          int o = k;
          if (af) {
            const double mk = af[k] > 0.5 ? 1.0 - af[k] : af[k];
            o = mk > mi ? i : k;
          }
          keep[o] = 0;
        }
      }
    }
    __syncthreads();  // keep[] of this window is final before the next lead is read
  }
}

}  // namespace ld
}  // namespace pcaone
