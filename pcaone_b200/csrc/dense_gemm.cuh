// pcaone_b200 — the same two power-iteration products for a GENERIC dense FP64 matrix
// (reference: RsvdOpOnePass::computeGandH, src/RSVD.hpp:137-166 and the window variant :168-252 —
// the PCAoneR entry point).
//
//   k_dense_g :  G_b   = D_b * Omega          D_b = rows [r0, r0 + nrows) of the tall matrix
//   k_dense_h :  Hpart = D_b^T * G_b          (split over the rows of the range)
//
// The same kernels with DOS = true read FLOAT32 DOSAGES (BGEN `minor_allele_dosage`, one row of N
// floats per variant, NaN = missing; reference FileBgen.cpp:15-72 in-core, :86-168 blocks) and fuse
// the reference's decode into the fragment load: value = NaN ? 0 : (d / 2 - F_j) * s_j with
// s_j = sqrt(ploidy) / sqrt(F_j (1 - F_j)) when standardising — mean imputation, centring and
// scaling, never materialising the N x M double matrix (4 bytes per genotype in HBM).
//
// D is the tall orientation of the input (rows >= cols; a wide input is used transposed, exactly
// like `trans` in RSVD.hpp:113-121), row-major [rows][ldd] doubles in HBM, ldd = cols rounded up
// to 8 with zero padding. Only the operand loader differs from gemm_fp64.cuh: the A fragments of
// the FP64 tensor-core MMAs (DMMA m8n8k4) come from a cp.async-staged tile of doubles instead of
// being decoded from 2-bit codes. Everything downstream (Omega update, QR(G) x2, SVD of the core)
// is the same device code as for genotypes.
#pragma once
#include "common.cuh"

namespace pcaone {

constexpr int kDenseThreads = 256;
constexpr int kDenseRows = 128;  // output rows per CTA (8 warps x 16)
constexpr int kDenseKC = 32;     // contraction chunk per pipeline stage
constexpr int kDenseLDA = kDenseKC + 4;    // row-major A tile [128][36]: 36 = 4 (mod 16) -> conflict-free A fragments
constexpr int kDenseLDT = kDenseRows + 4;  // k-major A tile [32][132] for the transposed product

constexpr int kDosLDA = kDenseKC + 4;     // float row-major A tile: 36 = 4 (mod 32) -> conflict-free
constexpr int kDosLDT = kDenseRows + 8;   // float k-major A tile: 136 = 8 (mod 32) -> conflict-free

__device__ __forceinline__ double dosage_value(float d, double F, double s) {
  return isnan(d) ? 0.0 : __dmul_rn(__dsub_rn((double)d * 0.5, F), s);
}

// allele frequency of dosage rows: F_j = sum(d / 2 over non-NaN) / #non-NaN (0 if none),
// FileBgen.cpp:27-41. One warp per variant; the double sum is a fixed-order lane/shuffle tree (the
// reference's own OpenMP reduction is not reproducible run to run, so this is tolerance-level).
__global__ void __launch_bounds__(256) k_dosage_af(const float* __restrict__ D, uint32_t ldd, uint32_t N, uint64_t rows,
                                                    double* __restrict__ F, uint32_t* __restrict__ nmiss) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t j = warp; j < rows; j += nwarps) {
    const float* row = D + j * ldd;
    double gs = 0.0;
    uint32_t gc = 0;
    for (uint32_t i = lane; i < N; i += 32) {
      const float d = row[i];
      if (!isnan(d)) {
        gs += (double)d * 0.5;
        gc += 1;
      }
    }
    gs = warp_sum(gs);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gc += __shfl_xor_sync(0xffffffffu, gc, o);
    if (lane == 0) {
      F[j] = gc ? gs / (double)gc : 0.0;
      nmiss[j] = N - gc;
    }
  }
}

template <int NT>
struct DenseSmem {
  static constexpr int LP = NT * 8;
  static constexpr int LDB = smem_ld(LP);
  static constexpr size_t kStageG = ((size_t)kDenseRows * kDenseLDA + (size_t)kDenseKC * LDB) * sizeof(double);
  static constexpr size_t kStageH = ((size_t)kDenseKC * kDenseLDT + (size_t)kDenseKC * LDB) * sizeof(double);
  // dosage variants: the A tile holds floats; the H kernel also stages F_r and s_r of the chunk
  static constexpr size_t kDosA_G = (size_t)kDenseRows * kDosLDA * sizeof(float);
  static constexpr size_t kDosA_H = (size_t)kDenseKC * kDosLDT * sizeof(float);
  static constexpr size_t kDosStageG = kDosA_G + (size_t)kDenseKC * LDB * sizeof(double);
  static constexpr size_t kDosStageH = kDosA_H + (size_t)kDenseKC * LDB * sizeof(double) + 2 * kDenseKC * sizeof(double);
};

// G[r][c] = sum_i D[r0 + r][i] * Omega[i][c],  r < nrows, i < ncols (ldd >= ncols, pad columns zero)
template <int NT>
__global__ void __launch_bounds__(kDenseThreads)
k_dense_g(const double* __restrict__ D, uint32_t ldd, uint32_t nrows, uint32_t ncols,
          const double* __restrict__ Omg,  // [ncols][8NT]
          double* __restrict__ G) {        // [nrows][8NT]
  using SM = DenseSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + SM::kStageG)};
  double* Bs[2] = {As[0] + kDenseRows * kDenseLDA, As[1] + kDenseRows * kDenseLDA};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t row0 = blockIdx.x * kDenseRows;
  const int nchunks = (int)((ncols + kDenseKC - 1) / kDenseKC);

  auto load = [&](int chunk, int stage) {
    const uint32_t k0 = (uint32_t)chunk * kDenseKC;
    constexpr int APIECES = kDenseKC / 2;  // 16-byte pieces per A row
    for (int idx = tid; idx < kDenseRows * APIECES; idx += kDenseThreads) {
      const int r = idx / APIECES, pc = idx - r * APIECES;
      const uint32_t row = row0 + r, k = k0 + 2 * pc;
      const bool ok = row < nrows && k < ldd;
      cp_async16(As[stage] + r * kDenseLDA + 2 * pc, D + (uint64_t)(ok ? row : 0) * ldd + (ok ? k : 0), ok ? 16 : 0);
    }
    constexpr int BPIECES = LP / 2;
    for (int idx = tid; idx < kDenseKC * BPIECES; idx += kDenseThreads) {
      const int r = idx / BPIECES, pc = idx - r * BPIECES;
      const uint32_t i = k0 + r;
      const bool ok = i < ncols;
      cp_async16(Bs[stage] + r * LDB + 2 * pc, Omg + (uint64_t)(ok ? i : 0) * LP + 2 * pc, ok ? 16 : 0);
    }
  };

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  load(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) load(c + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* A = As[st] + (warp * 16 + g) * kDenseLDA + t;
    const double* B = Bs[st] + t * LDB + g;
#pragma unroll
    for (int s = 0; s < kDenseKC / 4; ++s) {
      const double a0 = A[4 * s], a1 = A[8 * kDenseLDA + 4 * s];
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = B[4 * s * LDB + 8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t row = row0 + warp * 16 + 8 * u + g;
    if (row < nrows) {
      double* out = G + (uint64_t)row * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(out + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}

// Hpart[split][i][c] = sum_{r in split} D[r][i] * G[r][c],  i < ncols; grid = (col tiles of 128, splits)
template <int NT>
__global__ void __launch_bounds__(kDenseThreads)
k_dense_h(const double* __restrict__ D, uint32_t ldd, uint32_t nrows, uint32_t ncols,
          const double* __restrict__ G,    // [nrows][8NT]
          double* __restrict__ Hpart,      // [splits][ncols][8NT]
          uint32_t rows_per_split) {
  using SM = DenseSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + SM::kStageH)};
  double* Bs[2] = {As[0] + kDenseKC * kDenseLDT, As[1] + kDenseKC * kDenseLDT};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t i0 = blockIdx.x * kDenseRows;
  const uint32_t rbeg = blockIdx.y * rows_per_split;
  const uint32_t rend = min(nrows, rbeg + rows_per_split);
  const int nchunks = rend > rbeg ? (int)((rend - rbeg + kDenseKC - 1) / kDenseKC) : 0;

  auto load = [&](int chunk, int stage) {
    const uint32_t r0 = rbeg + (uint32_t)chunk * kDenseKC;
    constexpr int APIECES = kDenseRows / 2;
    for (int idx = tid; idx < kDenseKC * APIECES; idx += kDenseThreads) {
      const int r = idx / APIECES, pc = idx - r * APIECES;
      const uint32_t row = r0 + r, i = i0 + 2 * pc;
      const bool ok = row < rend && i < ldd;
      cp_async16(As[stage] + r * kDenseLDT + 2 * pc, D + (uint64_t)(ok ? row : 0) * ldd + (ok ? i : 0), ok ? 16 : 0);
    }
    constexpr int BPIECES = LP / 2;
    for (int idx = tid; idx < kDenseKC * BPIECES; idx += kDenseThreads) {
      const int r = idx / BPIECES, pc = idx - r * BPIECES;
      const uint32_t row = r0 + r;
      const bool ok = row < rend;
      cp_async16(Bs[stage] + r * LDB + 2 * pc, G + (uint64_t)(ok ? row : 0) * LP + 2 * pc, ok ? 16 : 0);
    }
  };

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  if (nchunks > 0) load(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) load(c + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* A = As[st] + t * kDenseLDT + warp * 16 + g;
    const double* B = Bs[st] + t * LDB + g;
#pragma unroll
    for (int s = 0; s < kDenseKC / 4; ++s) {
      const double a0 = A[4 * s * kDenseLDT], a1 = A[4 * s * kDenseLDT + 8];
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = B[4 * s * LDB + 8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  double* Hp = Hpart + (uint64_t)blockIdx.y * ncols * LP;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t i = i0 + warp * 16 + 8 * u + g;
    if (i < ncols) {
      double* out = Hp + (uint64_t)i * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(out + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}


// ------------------------------------------------------------------------------------------
// dosage variants: A tile = float dosages, decoded in the fragment load
// ------------------------------------------------------------------------------------------

// G[r][c] = sum_i x(r, i) * Omega[i][c], x(r, i) = dosage_value(D[r][i], F[r], s[r]); rows = variants
template <int NT>
__global__ void __launch_bounds__(kDenseThreads)
k_dos_g(const float* __restrict__ D, uint32_t ldd, uint32_t nrows, uint32_t ncols, const double* __restrict__ F,
        LutParams lp, const double* __restrict__ Omg, double* __restrict__ G) {
  using SM = DenseSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw + SM::kDosStageG)};
  double* Bs[2] = {reinterpret_cast<double*>(smem_raw + SM::kDosA_G),
                   reinterpret_cast<double*>(smem_raw + SM::kDosStageG + SM::kDosA_G)};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t row0 = blockIdx.x * kDenseRows;
  const int nchunks = (int)((ncols + kDenseKC - 1) / kDenseKC);
  double fj[2], sj[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t row = row0 + warp * 16 + 8 * u + g;
    fj[u] = row < nrows ? F[row] : 0.0;
    sj[u] = row < nrows ? snp_scale(fj[u], lp) : 0.0;
  }

  auto load = [&](int chunk, int stage) {
    const uint32_t k0 = (uint32_t)chunk * kDenseKC;
    constexpr int APIECES = kDenseKC / 4;  // 16-byte pieces (4 floats) per A row
    for (int idx = tid; idx < kDenseRows * APIECES; idx += kDenseThreads) {
      const int r = idx / APIECES, pc = idx - r * APIECES;
      const uint32_t row = row0 + r, k = k0 + 4 * pc;
      const bool ok = row < nrows && k < ldd;
      cp_async16(As[stage] + r * kDosLDA + 4 * pc, D + (uint64_t)(ok ? row : 0) * ldd + (ok ? k : 0), ok ? 16 : 0);
    }
    constexpr int BPIECES = LP / 2;
    for (int idx = tid; idx < kDenseKC * BPIECES; idx += kDenseThreads) {
      const int r = idx / BPIECES, pc = idx - r * BPIECES;
      const uint32_t i = k0 + r;
      const bool ok = i < ncols;
      cp_async16(Bs[stage] + r * LDB + 2 * pc, Omg + (uint64_t)(ok ? i : 0) * LP + 2 * pc, ok ? 16 : 0);
    }
  };

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  load(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) load(c + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* A = As[st] + (warp * 16 + g) * kDosLDA + t;
    const double* B = Bs[st] + t * LDB + g;
#pragma unroll
    for (int s = 0; s < kDenseKC / 4; ++s) {
      const double a0 = dosage_value(A[4 * s], fj[0], sj[0]);
      const double a1 = dosage_value(A[8 * kDosLDA + 4 * s], fj[1], sj[1]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = B[4 * s * LDB + 8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t row = row0 + warp * 16 + 8 * u + g;
    if (row < nrows) {
      double* out = G + (uint64_t)row * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(out + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}

// Hpart[split][i][c] = sum_{r in split} x(r, i) * G[r][c]; i = samples, r = variants of the range
template <int NT>
__global__ void __launch_bounds__(kDenseThreads)
k_dos_h(const float* __restrict__ D, uint32_t ldd, uint32_t nrows, uint32_t ncols, const double* __restrict__ F,
        LutParams lp, const double* __restrict__ G, double* __restrict__ Hpart, uint32_t rows_per_split) {
  using SM = DenseSmem<NT>;
  constexpr int LP = SM::LP, LDB = SM::LDB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As[2];
  double *Bs[2], *Fs[2];
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    unsigned char* base = smem_raw + st * SM::kDosStageH;
    As[st] = reinterpret_cast<float*>(base);
    Bs[st] = reinterpret_cast<double*>(base + SM::kDosA_H);
    Fs[st] = Bs[st] + kDenseKC * LDB;  // [0, KC): F_r, [KC, 2KC): s_r
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t i0 = blockIdx.x * kDenseRows;
  const uint32_t rbeg = blockIdx.y * rows_per_split;
  const uint32_t rend = min(nrows, rbeg + rows_per_split);
  const int nchunks = rend > rbeg ? (int)((rend - rbeg + kDenseKC - 1) / kDenseKC) : 0;

  auto load = [&](int chunk, int stage) {
    const uint32_t r0 = rbeg + (uint32_t)chunk * kDenseKC;
    constexpr int APIECES = kDenseRows / 4;
    for (int idx = tid; idx < kDenseKC * APIECES; idx += kDenseThreads) {
      const int r = idx / APIECES, pc = idx - r * APIECES;
      const uint32_t row = r0 + r, i = i0 + 4 * pc;
      const bool ok = row < rend && i < ldd;
      cp_async16(As[stage] + r * kDosLDT + 4 * pc, D + (uint64_t)(ok ? row : 0) * ldd + (ok ? i : 0), ok ? 16 : 0);
    }
    constexpr int BPIECES = LP / 2;
    for (int idx = tid; idx < kDenseKC * BPIECES; idx += kDenseThreads) {
      const int r = idx / BPIECES, pc = idx - r * BPIECES;
      const uint32_t row = r0 + r;
      const bool ok = row < rend;
      cp_async16(Bs[stage] + r * LDB + 2 * pc, G + (uint64_t)(ok ? row : 0) * LP + 2 * pc, ok ? 16 : 0);
    }
    if (tid < kDenseKC) {  // plain stores: visible after the __syncthreads that follows the wait
      const uint32_t row = r0 + tid;
      const double f = row < rend ? F[row] : 0.0;
      Fs[stage][tid] = f;
      Fs[stage][kDenseKC + tid] = row < rend ? snp_scale(f, lp) : 0.0;
    }
  };

  double acc[2][NT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[u][n][0] = acc[u][n][1] = 0.0;

  if (nchunks > 0) load(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (c + 1 < nchunks) load(c + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* A = As[st] + t * kDosLDT + warp * 16 + g;
    const double* B = Bs[st] + t * LDB + g;
    const double* Fc = Fs[st];
#pragma unroll
    for (int s = 0; s < kDenseKC / 4; ++s) {
      const double f = Fc[4 * s + t], sc = Fc[kDenseKC + 4 * s + t];
      const double a0 = dosage_value(A[4 * s * kDosLDT], f, sc);
      const double a1 = dosage_value(A[4 * s * kDosLDT + 8], f, sc);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double b = B[4 * s * LDB + 8 * n];
        dmma884(acc[0][n][0], acc[0][n][1], a0, b);
        dmma884(acc[1][n][0], acc[1][n][1], a1, b);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  double* Hp = Hpart + (uint64_t)blockIdx.y * ncols * LP;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t i = i0 + warp * 16 + 8 * u + g;
    if (i < ncols) {
      double* out = Hp + (uint64_t)i * LP + 2 * t;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<double2*>(out + 8 * n) = make_double2(acc[u][n][0], acc[u][n][1]);
    }
  }
}

// sum_i x_ij^2 of every dosage row (Selection.cpp:21,31), one warp per variant
__global__ void __launch_bounds__(256) k_dosage_sqnorm(const float* __restrict__ D, uint32_t ldd, uint32_t N, uint64_t rows,
                                                        const double* __restrict__ F, LutParams lp,
                                                        double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t j = warp; j < rows; j += nwarps) {
    const double f = F[j], sc = snp_scale(f, lp);
    double s = 0.0;
    for (uint32_t i = lane; i < N; i += 32) {
      const double x = dosage_value(D[j * ldd + i], f, sc);
      s += x * x;
    }
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
  }
}

// dense decode of dosage rows [start, start + nrows) -> col-major N x nrows doubles (Eigen layout
// of data->G), for read_block parity checks (FileBgen.cpp:96-110)
__global__ void k_dosage_decode(const float* __restrict__ D, uint32_t ldd, uint32_t N, uint64_t nrows,
                                const double* __restrict__ F, LutParams lp, double* __restrict__ out) {
  const uint64_t total = nrows * N;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / N;
    const uint32_t i = (uint32_t)(idx - r * N);
    const double f = F[r];
    out[idx] = dosage_value(D[r * ldd + i], f, snp_scale(f, lp));
  }
}


// ------------------------------------------------------------------------------------------
// Beagle genotype likelihoods (PCAngsd): FileBeagle.cpp:14-68, Utils.cpp:745-775, Data.cpp:296-316
// P: [rows][2N] doubles, (P0, P1) of sample i at 2i, 2i+1 — the columns of the reference's 2N x M
// matrix P. The expected genotypes E = G are MATERIALISED as the dense source ([rows][ldd] doubles),
// like the reference does in core: E changes once per EM iteration, not per pass, so evaluating the
// k-term individual allele frequency inside every GEMM operand load would repeat that work
// (2 products x epochs) times.
// ------------------------------------------------------------------------------------------

// one EM step of emMAF_with_GL (Utils.cpp:753-765) for every variant: one warp per variant.
// Fnew[j] = sum_i (p1 + 2 p2) / (p0 + p1 + p2) / (2N); sq[j] = (Fnew - F)^2
__global__ void __launch_bounds__(256) k_gl_maf_step(const double* __restrict__ P, uint32_t N, uint64_t rows,
                                                      const double* __restrict__ F, double* __restrict__ Fnew,
                                                      double* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const double scale = 1.0 / (2.0 * (double)N);
  for (uint64_t j = warp; j < rows; j += nwarps) {
    const double2* row = reinterpret_cast<const double2*>(P + j * 2ull * N);
    const double f = F[j], omf = 1.0 - f;
    double pt = 0.0;
    for (uint32_t i = lane; i < N; i += 32) {
      const double2 g = row[i];
      const double p0 = __dmul_rn(__dmul_rn(g.x, omf), omf);
      const double p1 = __dmul_rn(__dmul_rn(__dmul_rn(g.y, 2.0), f), omf);
      const double p2 = __dmul_rn(__dmul_rn(__dsub_rn(__dsub_rn(1.0, g.x), g.y), f), f);
      pt += __ddiv_rn(__dadd_rn(p1, __dmul_rn(2.0, p2)), __dadd_rn(__dadd_rn(p0, p1), p2));
    }
    pt = warp_sum(pt);
    if (lane == 0) {
      const double fn = pt * scale;
      Fnew[j] = fn;
      sq[j] = (fn - f) * (fn - f);
    }
  }
}

// out[0] = sum of v[0..n) in a fixed order (one CTA)
__global__ void __launch_bounds__(1024) k_sum_fixed(const double* __restrict__ v, uint64_t n, double* __restrict__ out) {
  __shared__ double s[1024];
  double a = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += 1024) a += v[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

// E[j][i] = (p1 + 2 p2) / (p0 + p1 + p2) - 2 F_j with pt = F_j (initial E, FileBeagle.cpp:57-66) or the
// individual allele frequency pt = clamp((U_i . (S o V_j) + 2 F_j) / 2, 1e-4, 1 - 1e-4) (fit_with_pi,
// Data.cpp:296-316). U: [N][ldu], V: [rows][ldv] row-major device layouts; U == nullptr -> initial.
__global__ void k_gl_expected(const double* __restrict__ P, uint32_t N, uint64_t rows, const double* __restrict__ F,
                              const double* __restrict__ U, int ldu, const double* __restrict__ S,
                              const double* __restrict__ V, int ldv, int k, double* __restrict__ E, uint32_t ldd) {
  const uint64_t total = rows * N;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = idx / N;
    const uint32_t i = (uint32_t)(idx - j * N);
    const double f = F[j];
    double pt = f;
    if (U) {
      pt = 0.0;
      for (int r = 0; r < k; ++r) pt = __dadd_rn(pt, __dmul_rn(__dmul_rn(U[(uint64_t)i * ldu + r], S[r]), V[j * ldv + r]));
      pt = __ddiv_rn(__dadd_rn(pt, __dmul_rn(2.0, f)), 2.0);
      pt = fmin(fmax(pt, 1e-4), 1.0 - 1e-4);
    }
    const double2 g = reinterpret_cast<const double2*>(P + j * 2ull * N)[i];
    const double omp = __dsub_rn(1.0, pt);
    const double p0 = __dmul_rn(__dmul_rn(g.x, omp), omp);
    const double p1 = __dmul_rn(__dmul_rn(__dmul_rn(g.y, 2.0), pt), omp);
    const double p2 = __dmul_rn(__dmul_rn(__dsub_rn(__dsub_rn(1.0, g.x), g.y), pt), pt);
    E[j * ldd + i] = __dsub_rn(__ddiv_rn(__dadd_rn(p1, __dmul_rn(2.0, p2)), __dadd_rn(__dadd_rn(p0, p1), p2)),
                               __dmul_rn(2.0, f));
  }
}

// pcangsd_standardize_E (Data.cpp:364-407): with the individual allele frequencies of the final U, S, V,
//   E[j][i] = ((p1 + 2 p2) / pSum - 2 F_j) / sqrt(2 F_j (1 - F_j))      (the division only when the norm > kVarTol)
//   Dc[i]  += ((0-2F)^2 p0 + (1-2F)^2 p1 + (2-2F)^2 p2) / pSum / (2 F_j (1 - F_j))
// One thread per sample, blockIdx.y walks SNP slices: the diagonal sums stay in a register and reach memory as ONE
// atomic per thread and slice.
__global__ void __launch_bounds__(256) k_gl_grm(const double* __restrict__ P, uint32_t N, uint64_t rows,
                                                uint64_t rows_per_slice, const double* __restrict__ F,
                                                const double* __restrict__ U, int ldu, const double* __restrict__ S,
                                                const double* __restrict__ V, int ldv, int k, double* __restrict__ E,
                                                uint32_t ldd, double* __restrict__ Dc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint64_t j0 = (uint64_t)blockIdx.y * rows_per_slice, j1 = min(rows, j0 + rows_per_slice);
  double dc = 0.0;
  for (uint64_t j = j0; j < j1; ++j) {
    const double f = F[j];
    double pt = 0.0;
    for (int r = 0; r < k; ++r) pt = __dadd_rn(pt, __dmul_rn(__dmul_rn(U[(uint64_t)i * ldu + r], S[r]), V[j * ldv + r]));
    pt = __ddiv_rn(__dadd_rn(pt, __dmul_rn(2.0, f)), 2.0);
    pt = fmin(fmax(pt, 1e-4), 1.0 - 1e-4);
    const double2 g = reinterpret_cast<const double2*>(P + j * 2ull * N)[i];
    const double omp = __dsub_rn(1.0, pt);
    const double p0 = __dmul_rn(__dmul_rn(g.x, omp), omp);
    const double p1 = __dmul_rn(__dmul_rn(__dmul_rn(g.y, 2.0), pt), omp);
    const double p2 = __dmul_rn(__dmul_rn(__dsub_rn(__dsub_rn(1.0, g.x), g.y), pt), pt);
    const double ps = __dadd_rn(__dadd_rn(p0, p1), p2);
    const double tf = __dmul_rn(2.0, f);
    const double norm = __dsqrt_rn(__dmul_rn(tf, __dsub_rn(1.0, f)));
    double e = __dsub_rn(__ddiv_rn(__dadd_rn(p1, __dmul_rn(2.0, p2)), ps), tf);
    if (norm > kVarTol) e = __ddiv_rn(e, norm);
    E[j * ldd + i] = e;
    const double d0 = __dsub_rn(0.0, tf), d1 = __dsub_rn(1.0, tf), d2 = __dsub_rn(2.0, tf);
    double t = __dmul_rn(__dmul_rn(d0, d0), __ddiv_rn(p0, ps));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(d1, d1), __ddiv_rn(p1, ps)));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(d2, d2), __ddiv_rn(p2, ps)));
    dc = __dadd_rn(dc, __ddiv_rn(t, __dmul_rn(tf, __dsub_rn(1.0, f))));
  }
  atomicAdd(&Dc[i], dc);
}

// col-major (rows x cols, ld = rows) -> row-major [rows][ldd], 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) k_dense_transpose_in(const double* __restrict__ src, uint64_t rows, uint64_t cols,
                                                             double* __restrict__ dst, uint32_t ldd) {
  __shared__ double tile[32][33];
  const uint64_t r0 = (uint64_t)blockIdx.x * 32, c0 = (uint64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const uint64_t r = r0 + tx, c = c0 + j;
    tile[j][tx] = (r < rows && c < cols) ? src[c * rows + r] : 0.0;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const uint64_t r = r0 + j, c = c0 + tx;
    if (r < rows && c < ldd) dst[r * ldd + c] = tile[tx][j];
  }
}

}  // namespace pcaone
