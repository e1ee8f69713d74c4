// pcaone_b200 — engine + C-ABI (include/pcaone_b200.h).
//
// Host-side state machine of the randomized-SVD hot path, driving the sm_100a kernels:
//   RsvdOpData::computeUSV                 reference src/Halko.cpp:46-97
//   NormalRsvdOpData::computeGandH         reference src/Halko.cpp:99-153
//   FancyRsvdOpData::computeGandH          reference src/Halko.cpp:155-269
//   run_pca_with_halko EM loop             reference src/Halko.cpp:290-319
//   FileBed::read_all / read_block_*       reference src/FilePlink.cpp:26-298
// Nothing here falls back to the CPU: every arithmetic step is a kernel launch.
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "../../include/pcaone_b200.h"
#include "common.cuh"
#include "decode.cuh"
#include "gemm_fp64.cuh"
#include "small_dense.cuh"
#include "tall_skinny.cuh"
#include "orth_fused.cuh"
#include "tc_gemm.cuh"
#include "dense_gemm.cuh"
#include "ld.cuh"

using namespace pcaone;

namespace {
thread_local std::string g_create_err;

struct EvPair {
  cudaEvent_t a, b;
  int kind;  // 0 gemm_g, 1 gemm_h, 2 orth, 3 small, 4 h2d, 5 allreduce
};
}  // namespace

struct pcaone_ctx {
  pcaone_config cfg{};
  std::string err;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  int sms = 148;

  uint64_t N = 0, M = 0;
  int k = 0, l = 0, NT = 0, lp = 0;
  uint32_t bpr = 0, pitch = 0;
  LutParams lut{};
  int update = 0, standardize = 0;

  // genotype source
  int source = -1;
  double* d_dense = nullptr;  // PCAONE_SRC_DENSE: tall orientation of a generic matrix, row-major [M][ldd]
  uint32_t ldd = 0;
  double* d_P = nullptr;      // PCAONE_SRC_GL: genotype likelihoods [M][2N]; the expected genotypes E live in d_dense
  float* d_dos = nullptr;     // PCAONE_SRC_DOSAGE: float dosages, row-major [M][ldf], NaN = missing
  uint32_t ldf = 0;
  uint8_t* d_packed = nullptr;  // resident, M x pitch
  const uint8_t* h_packed = nullptr;
  pcaone_read_block_fn reader = nullptr;
  void* reader_user = nullptr;
  FILE* bed_file = nullptr;
  uint64_t bed_snp_offset = 0;
  std::vector<uint64_t> blk_start, blk_stop;
  uint32_t band_factor = 1;
  uint64_t max_block = 0;
  uint8_t* d_raw[2] = {nullptr, nullptr};
  uint8_t* d_blk[2] = {nullptr, nullptr};
  uint8_t* h_pin[2] = {nullptr, nullptr};
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  bool af_done = false;

  // per-SNP
  double* d_F = nullptr;
  uint32_t* d_nmiss = nullptr;

  // tall matrices, row-major [rows][lp]
  double *d_Omg0 = nullptr, *d_Omg = nullptr, *d_Omg2 = nullptr, *d_H = nullptr, *d_H1 = nullptr, *d_H2 = nullptr, *d_Bt = nullptr,
         *d_Ucur = nullptr, *d_Upre = nullptr, *d_U = nullptr;
  double *d_G = nullptr, *d_V = nullptr, *d_Vpre = nullptr;
  double* d_S = nullptr;
  double* d_Hpart = nullptr;
  uint32_t max_splits = 1;
  bool have_usv = false, have_omg0 = false;

  // small l x l (ld = lp)
  double *d_W = nullptr, *d_R = nullptr, *d_Rinv = nullptr, *d_T1 = nullptr, *d_T2 = nullptr, *d_T = nullptr,
         *d_Vr = nullptr, *d_Z = nullptr, *d_sigma = nullptr, *d_sign = nullptr, *d_hsign = nullptr, *d_scal = nullptr;
  int* d_status = nullptr;
  int* h_status = nullptr;    // pinned
  double* h_scal = nullptr;   // pinned
  double* d_part = nullptr;   // partial workspace for two-stage reductions
  size_t part_doubles = 0;
  unsigned long long* d_pidx = nullptr;
  double* d_stage = nullptr;  // col-major staging for host transfers
  size_t stage_doubles = 0;

  // winSVD state (FancyRsvdOpData members, Halko.hpp:66-68)
  uint64_t bandsize = 1;

  // tensor-core (int8 Ozaki) path, tc_gemm.cuh. slices == 0 -> FP64 DMMA only.
  int slices = 0, NP = 0, RT = 1;
  uint8_t *d_PG = nullptr, *d_PH = nullptr;            // resident tiled operands (rows = SNPs / rows = samples)
  uint8_t *d_PGb[2] = {nullptr, nullptr}, *d_PHb[2] = {nullptr, nullptr};  // per streamed block
  bool tiles_valid = false;
  int8_t *d_BimgO = nullptr, *d_BimgW = nullptr;       // B operand images: Omega, W = s o G of the current range
  size_t bimgW_kb = 0;
  long long* d_Racc = nullptr;                         // int64 accumulators
  long long* d_Racc2 = nullptr;                        // int64 accumulators of the missing-mask products
  int8_t* d_BimgD = nullptr;                           // B image of D = (f - 1) o W (mask operand of the H pass)
  size_t R2_rows = 0, bimgD_kb = 0;
  size_t R_rows = 0;
  unsigned long long* d_tcs = nullptr;                 // [5][lp] + 1: Omega colmax, Omega Csum, W colmax, W Csum, Fw, block counter
  double* d_Fpart = nullptr;
  bool omega_img_valid = false;
  bool omega_colmax_valid = false;                     // d_tcs colmax of Omega was produced by the orth kernel
  const double* sum_other = nullptr;                   // winSVD: the next finish_h also writes sum_out = Hacc + sum_other
  double* sum_out = nullptr;
  bool sum_done = false;
  std::vector<uint32_t> h_nmiss;                       // per local SNP; UINT32_MAX = not known yet
  std::vector<uint64_t> nmiss_prefix;
  uint64_t tc_ranges = 0, fp64_ranges = 0, tc_miss_ranges = 0;
  int half = 3;                                        // which products a range runs: 1 = G rows only, 2 = H only, 3 = both
  bool g_is_q = false;                                 // d_G holds Q = G T after small_stage (else raw G)
  double* d_jscratch = nullptr;                        // eigen-fallback scratch of k_orth_fused
  int fused_orth = 1;                                  // PCAONE_FUSED_ORTH=0 selects the multi-kernel path

  pcaone_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;

  // measurement
  bool timing = false;
  std::vector<EvPair> evs;
  pcaone_timers tm{};
  double last_diff = 0.0;
  int last_epochs = 0;
};

namespace {

#define CTX_GUARD(ctx, ...)                     \
  if (!(ctx)) return 1;                         \
  try {                                         \
    PCA_CUDA(cudaSetDevice((ctx)->cfg.device)); \
    __VA_ARGS__;                                \
    return 0;                                   \
  } catch (const std::exception& e) {           \
    (ctx)->err = e.what();                      \
    return 1;                                   \
  }

template <class T>
void dmalloc(T** p, size_t n) {
  PCA_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
}

int grid_for(uint64_t work, int threads, int sms) {
  uint64_t b = (work + threads - 1) / threads;
  uint64_t cap = (uint64_t)sms * 16;
  return (int)std::max<uint64_t>(1, std::min(b, cap));
}

int supported_nt(int l) {
  static const int opts[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  const int need = (l + 7) / 8;
  for (int o : opts)
    if (o >= need) return o;
  return -1;
}

struct Timed {
  pcaone_ctx* c;
  EvPair ev{};
  bool on;
  Timed(pcaone_ctx* c_, int kind) : c(c_), on(c_->timing) {
    if (on) {
      PCA_CUDA(cudaEventCreate(&ev.a));
      PCA_CUDA(cudaEventCreate(&ev.b));
      ev.kind = kind;
      PCA_CUDA(cudaEventRecord(ev.a, c->stream));
    }
  }
  ~Timed() {
    if (on) {
      cudaEventRecord(ev.b, c->stream);
      c->evs.push_back(ev);
    }
  }
};

void resolve_timers(pcaone_ctx* c) {
  if (c->evs.empty()) return;
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& e : c->evs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.a, e.b);
    switch (e.kind) {
      case 0: c->tm.gemm_g_ms += ms; break;
      case 1: c->tm.gemm_h_ms += ms; break;
      case 2: c->tm.orth_ms += ms; break;
      case 3: c->tm.small_ms += ms; break;
      case 4: c->tm.h2d_ms += ms; break;
      case 5: c->tm.allreduce_ms += ms; break;
      case 6: c->tm.decode_ms += ms; break;
      case 7: c->tm.tc_g_ms += ms; break;
      case 8: c->tm.tc_h_ms += ms; break;
      case 9: c->tm.ld_ms += ms; break;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  c->evs.clear();
}

// ---------------------------------------------------------------- kernel dispatch on NT
template <int NT>
void gemm_g_nt(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, const double* F, double* G, const double* Vrows) {
  const size_t smem = 2 * GemmSmem<NT>::kStageG;
  const int grid = ceil_div(nrows, kTileRows);
  if (c->update && c->cfg.emu) {
    static bool attr = false;
    if (!attr) {
      PCA_CUDA(cudaFuncSetAttribute(k_gemm_g<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    k_gemm_g<NT, true><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, c->d_Omg, F,
                                                                c->lut, G, c->d_U, c->lp, c->d_S, Vrows, c->lp, c->k);
  } else {
    static bool attr = false;
    if (!attr) {
      PCA_CUDA(cudaFuncSetAttribute(k_gemm_g<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
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
    static bool attr = false;
    if (!attr) {
      PCA_CUDA(cudaFuncSetAttribute(k_gemm_h<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    k_gemm_h<NT, true><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, G, F, c->lut,
                                                                c->d_Hpart, rows_per_split, c->d_U, c->lp, c->d_S,
                                                                Vrows, c->lp, c->k);
  } else {
    static bool attr = false;
    if (!attr) {
      PCA_CUDA(cudaFuncSetAttribute(k_gemm_h<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    k_gemm_h<NT, false><<<grid, kGemmThreads, smem, c->stream>>>(P, c->pitch, nrows, (uint32_t)c->N, G, F, c->lut,
                                                                 c->d_Hpart, rows_per_split, nullptr, 0, nullptr,
                                                                 nullptr, 0, 0);
  }
  PCA_CHECK_LAUNCH();
}

#define NT_DISPATCH(fn, ...)                                   \
  switch (c->NT) {                                             \
    case 1: fn<1>(__VA_ARGS__); break;                         \
    case 2: fn<2>(__VA_ARGS__); break;                         \
    case 3: fn<3>(__VA_ARGS__); break;                         \
    case 4: fn<4>(__VA_ARGS__); break;                         \
    case 5: fn<5>(__VA_ARGS__); break;                         \
    case 6: fn<6>(__VA_ARGS__); break;                         \
    case 8: fn<8>(__VA_ARGS__); break;                         \
    case 10: fn<10>(__VA_ARGS__); break;                       \
    case 12: fn<12>(__VA_ARGS__); break;                       \
    case 16: fn<16>(__VA_ARGS__); break;                       \
    default: throw std::runtime_error("unsupported NT");       \
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

// ---------------------------------------------------------------- tensor-core (int8 Ozaki) path
uint64_t tc_nkb_samples(const pcaone_ctx* c) { return (c->N + tc::kKB - 1) / tc::kKB; }
uint64_t tc_nrt_samples(const pcaone_ctx* c) { return (c->N + tc::kRowTile - 1) / tc::kRowTile; }
size_t tc_pg_bytes(const pcaone_ctx* c, uint64_t rows) {
  return (size_t)((rows + tc::kRowTile - 1) / tc::kRowTile) * tc_nkb_samples(c) * tc::kChunkBytes;
}
size_t tc_ph_bytes(const pcaone_ctx* c, uint64_t rows) {
  return (size_t)((rows + tc::kKB - 1) / tc::kKB) * tc_nrt_samples(c) * tc::kChunkBytes;
}

// tiled copies of `rows` packed SNP rows at P: PG (rows = SNPs) and PH (rows = samples)
void tc_build_tiles(pcaone_ctx* c, const uint8_t* P, uint64_t rows, uint8_t* PG, uint8_t* PH, cudaStream_t st) {
  const uint32_t nkb = (uint32_t)tc_nkb_samples(c), nrt = (uint32_t)tc_nrt_samples(c);
  const uint64_t work = (uint64_t)((rows + tc::kRowTile - 1) / tc::kRowTile) * nkb * tc::kRowTile;
  tc::k_tile_rows<<<grid_for(work, 256, c->sms), 256, 0, st>>>(P, c->pitch, rows, (uint32_t)c->N, nkb, PG);
  PCA_CHECK_LAUNCH();
  const uint64_t nkbh = (rows + tc::kKB - 1) / tc::kKB;
  tc::k_tile_transpose<<<(unsigned)(nkbh * nrt), 128, 0, st>>>(P, c->pitch, rows, (uint32_t)c->N, nrt, PH);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches += 2;
}

void tc_alloc(pcaone_ctx* c, uint64_t max_range_rows, bool miss) {
  if (!c->d_tcs) {
    dmalloc(&c->d_tcs, (size_t)5 * c->lp + 1);  // + the block counter of the slice kernel's Fw reduction
    PCA_CUDA(cudaMemsetAsync(c->d_tcs, 0, ((size_t)5 * c->lp + 1) * sizeof(unsigned long long), c->stream));
    dmalloc(&c->d_BimgO, (size_t)(tc_nkb_samples(c) + 1) * tc::kKB * c->NP);
  }
  const size_t need_rows = std::max<uint64_t>(tc_nrt_samples(c) * tc::kRowTile, max_range_rows + 2 * tc::kRowTile);
  if (need_rows > c->R_rows) {
    if (c->d_Racc) cudaFree(c->d_Racc);
    dmalloc(&c->d_Racc, need_rows * c->lp);
    PCA_CUDA(cudaMemsetAsync(c->d_Racc, 0, need_rows * c->lp * sizeof(long long), c->stream));
    c->R_rows = need_rows;
  }
  const size_t need_kb = max_range_rows / tc::kKB + 4;
  if (miss && need_rows > c->R2_rows) {
    if (c->d_Racc2) cudaFree(c->d_Racc2);
    dmalloc(&c->d_Racc2, need_rows * c->lp);
    PCA_CUDA(cudaMemsetAsync(c->d_Racc2, 0, need_rows * c->lp * sizeof(long long), c->stream));
    c->R2_rows = need_rows;
  }
  if (miss && need_kb > c->bimgD_kb) {
    if (c->d_BimgD) cudaFree(c->d_BimgD);
    dmalloc(&c->d_BimgD, need_kb * tc::kKB * c->NP);
    c->bimgD_kb = need_kb;
  }
  if (need_kb > c->bimgW_kb) {
    if (c->d_BimgW) cudaFree(c->d_BimgW);
    if (c->d_Fpart) cudaFree(c->d_Fpart);
    dmalloc(&c->d_BimgW, need_kb * tc::kKB * c->NP);
    dmalloc(&c->d_Fpart, need_kb * c->lp);
    c->bimgW_kb = need_kb;
  }
}

// missing genotypes among local SNPs [s, s+n): UINT64_MAX if not known on the host yet
uint64_t tc_missing_in(pcaone_ctx* c, uint64_t s, uint64_t n) {
  if (c->h_nmiss.size() != c->M) return UINT64_MAX;
  if (c->nmiss_prefix.size() == c->M + 1) return c->nmiss_prefix[s + n] - c->nmiss_prefix[s];
  uint64_t tot = 0;
  for (uint64_t j = s; j < s + n; ++j) {
    if (c->h_nmiss[j] == UINT32_MAX) return UINT64_MAX;
    tot += c->h_nmiss[j];
  }
  return tot;
}
void tc_fetch_nmiss(pcaone_ctx* c, uint64_t s, uint64_t n) {
  if (c->h_nmiss.size() != c->M) c->h_nmiss.assign(c->M, UINT32_MAX);
  PCA_CUDA(cudaMemcpyAsync(c->h_nmiss.data() + s, c->d_nmiss + s, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  c->nmiss_prefix.clear();
  bool all = true;
  for (uint64_t j = 0; j < c->M && all; ++j) all = c->h_nmiss[j] != UINT32_MAX;
  if (all) {
    c->nmiss_prefix.resize(c->M + 1);
    c->nmiss_prefix[0] = 0;
    for (uint64_t j = 0; j < c->M; ++j) c->nmiss_prefix[j + 1] = c->nmiss_prefix[j] + c->h_nmiss[j];
  }
}

template <int S, int RT, int MODE>
void tc_launch_st(pcaone_ctx* c, const tc::TcGemmArgs& a, int grid) {
  const size_t smem = tc::tc_smem_bytes(RT, c->NP);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(tc::k_tc_gemm<S, RT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  tc::k_tc_gemm<S, RT, MODE><<<grid, tc::tc_threads(RT), smem, c->stream>>>(a);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// mode: tc::kPlain / kNonMiss / kMask (what the packed operand decodes to); R: int64 accumulators
void tc_launch(pcaone_ctx* c, tc::TcGemmArgs a, int mode, long long* R) {
  // split-K so that (row-tile groups x splits) fills the SMs in whole waves
  const uint32_t n_rtp = (a.nrt + c->RT - 1) / c->RT;
  const uint32_t epi_cost = 24;  // epilogue + pipeline fill, in k-block units
  uint32_t best_ns = 1;
  uint64_t best = UINT64_MAX;
  const uint32_t max_ns = std::max<uint32_t>(1, a.nkb / 8);
  for (uint32_t ns = 1; ns <= std::min<uint32_t>(max_ns, 4u * c->sms); ++ns) {
    const uint64_t per = ((a.nkb + ns - 1) / ns + 1) & ~1ull;
    const uint64_t waves = ((uint64_t)n_rtp * ns + c->sms - 1) / c->sms;
    const uint64_t cost = waves * (per + epi_cost);
    if (cost < best) {
      best = cost;
      best_ns = ns;
    }
  }
  a.kb_per_split = ((a.nkb + best_ns - 1) / best_ns + 1) & ~1u;
  a.nsplit = (a.nkb + a.kb_per_split - 1) / a.kb_per_split;
  if ((uint64_t)a.kb_per_split * tc::kKB >= (1ull << 22)) throw std::runtime_error("tc_gemm: contraction too long for exact s32 sums");
  const int grid = (int)std::min<uint64_t>((uint64_t)n_rtp * a.nsplit, (uint64_t)c->sms);
  a.NP = c->NP;
  a.l = c->l;
  a.lp = c->lp;
  a.R = R;
#define TC_CASE(S_, RT_)                                                                        \
  if (c->slices == S_ && c->RT == RT_) {                                                        \
    if (mode == tc::kPlain) tc_launch_st<S_, RT_, tc::kPlain>(c, a, grid);                      \
    else if (mode == tc::kNonMiss) tc_launch_st<S_, RT_, tc::kNonMiss>(c, a, grid);             \
    else tc_launch_st<S_, RT_, tc::kMask>(c, a, grid);                                          \
    return;                                                                                     \
  }
  TC_CASE(2, 1) TC_CASE(2, 2) TC_CASE(3, 1) TC_CASE(3, 2) TC_CASE(4, 1) TC_CASE(4, 2)
#undef TC_CASE
  throw std::runtime_error("tc_gemm: unsupported slice count");
}

constexpr uint32_t kFoldFwMaxParts = 512;

void tc_slice(pcaone_ctx* c, double* X, uint64_t r0, uint64_t r1, const unsigned long long* colmax, const double* F,
              int writeback, int8_t* Bimg, long long* Csum, double* Fpart, uint32_t* nkb_out, int dmode = 0) {
  tc::TcSliceArgs a{};
  a.X = X;
  a.lp = c->lp;
  a.l = c->l;
  a.S = c->slices;
  a.NP = c->NP;
  a.r0 = r0;
  a.r1 = r1;
  a.kb0 = (uint32_t)(r0 / tc::kKB);
  a.colmax = colmax;
  a.F = F;
  a.lut = c->lut;
  a.writeback = writeback;
  a.dmode = dmode;
  a.Bimg = Bimg;
  a.Csum = Csum;
  a.Fpart = Fpart;
  // an even number of k-block images: a pipeline stage of k_tc_gemm is two k-blocks (the pad image is zero)
  const uint32_t nkb = ((uint32_t)((r1 - 1) / tc::kKB) - a.kb0 + 2) & ~1u;
  // window-sized launches fold the Fw reduction into their last block; with thousands of partials
  // (merged ranges of the late epochs) the single block would be a long tail: separate kernel
  const bool fold_fw = Fpart && nkb <= kFoldFwMaxParts;
  a.Fw = fold_fw ? reinterpret_cast<double*>(c->d_tcs + 4 * c->lp) : nullptr;
  a.done = reinterpret_cast<unsigned int*>(c->d_tcs + 5 * c->lp);
  const size_t smem = (size_t)tc::kKB * c->NP;
  tc::k_tc_slice<<<nkb, tc::tc_flat_threads(c->lp), smem, c->stream>>>(a);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
  if (nkb_out) *nkb_out = nkb;
}

// tensor-core version of range_gemms. PG/PH: tiled operands in which the range starts at local
// row / contraction index `loc0`; snp0 = first SNP of the range in d_G / d_F.
// `miss`: the range contains missing calls -> every product is run as (non-missing counts, mask) pair.
void range_gemms_tc(pcaone_ctx* c, const uint8_t* PG, const uint8_t* PH, uint64_t loc0, uint32_t nrows, uint64_t snp0,
                    double* Hacc, bool accumulate, bool miss) {
  const int mode = miss ? tc::kNonMiss : tc::kPlain;
  unsigned long long* o_colmax = c->d_tcs;
  long long* o_csum = reinterpret_cast<long long*>(c->d_tcs + c->lp);
  unsigned long long* w_colmax = c->d_tcs + 2 * c->lp;
  long long* w_csum = reinterpret_cast<long long*>(c->d_tcs + 3 * c->lp);
  const uint32_t nkb_s = (uint32_t)tc_nkb_samples(c), nrt_s = (uint32_t)tc_nrt_samples(c);
  {
    Timed t(c, 0);
    if (!c->omega_img_valid) {
      if (!c->omega_colmax_valid) {
        PCA_CUDA(cudaMemsetAsync(c->d_tcs, 0, (size_t)2 * c->lp * sizeof(unsigned long long), c->stream));
        tc::k_tc_colmax<<<grid_for(c->N * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_Omg, c->lp, c->l, 0, c->N, o_colmax);
        PCA_CHECK_LAUNCH();
        c->tm.kernel_launches++;
      }
      c->omega_colmax_valid = false;
      tc_slice(c, c->d_Omg, 0, c->N, o_colmax, nullptr, 0, c->d_BimgO, o_csum, nullptr, nullptr);
      c->omega_img_valid = true;
    }
    tc::TcGemmArgs a{};
    a.zero_ptr = w_colmax;  // W column maxima + column sums of this range (finish_g / slice accumulate into them)
    a.zero_n = 2 * (uint32_t)c->lp;
    a.PA = PG;
    a.stride_rt = (uint64_t)nkb_s * tc::kChunkBytes;
    a.stride_kb = tc::kChunkBytes;
    a.Bimg = c->d_BimgO;
    a.rt0 = (uint32_t)(loc0 / tc::kRowTile);
    a.nrt = (uint32_t)((loc0 + nrows - 1) / tc::kRowTile) - a.rt0 + 1;
    a.kb0 = 0;
    a.nkb = (nkb_s + 1) & ~1u;
    a.kb_valid_last = nkb_s - 1;
    a.row_begin = (long long)loc0;
    a.row_end = (long long)(loc0 + nrows);
    a.row_r0 = (long long)a.rt0 * tc::kRowTile;
    {
      Timed tk(c, 7);
      tc_launch(c, a, mode, c->d_Racc);
      a.zero_ptr = nullptr;
      if (miss) tc_launch(c, a, tc::kMask, c->d_Racc2);
    }
    const uint64_t roff = (loc0 - (uint64_t)a.row_r0) * c->lp;
    tc::k_tc_finish_g<<<(unsigned)std::min<uint64_t>((nrows + tc::kKB - 1) / tc::kKB, (uint64_t)c->sms * 8),
                        tc::tc_pair_threads(c->lp), 0, c->stream>>>(
        c->d_Racc + roff, miss ? c->d_Racc2 + roff : nullptr, nrows, c->l, c->lp, c->slices, c->d_F + snp0, c->lut, o_csum,
        o_colmax, c->d_G + snp0 * c->lp, w_colmax);
    PCA_CHECK_LAUNCH();
    c->tm.gemm_g_launches++;
    c->tm.kernel_launches++;
  }
  {
    Timed t(c, 1);
    uint32_t nkb_w = 0;
    // contraction index = loc0 + (row of d_G - snp0): hand the slice kernel pointers to index 0
    double* X0 = c->d_G + snp0 * c->lp - loc0 * c->lp;
    const double* F0 = c->d_F + snp0 - loc0;
    tc_slice(c, X0, loc0, loc0 + nrows, w_colmax, F0, 1, c->d_BimgW, w_csum, c->d_Fpart, &nkb_w);
    if (miss) tc_slice(c, X0, loc0, loc0 + nrows, w_colmax, F0, 0, c->d_BimgD, nullptr, nullptr, nullptr, 1);
    tc::TcGemmArgs a{};
    a.PA = PH;
    a.stride_rt = tc::kChunkBytes;
    a.stride_kb = (uint64_t)nrt_s * tc::kChunkBytes;
    a.Bimg = c->d_BimgW;
    a.rt0 = 0;
    a.nrt = nrt_s;
    a.kb0 = (uint32_t)(loc0 / tc::kKB);
    a.nkb = nkb_w;
    a.kb_valid_last = (uint32_t)((loc0 + nrows - 1) / tc::kKB);
    a.row_begin = 0;
    a.row_end = (long long)c->N;
    a.row_r0 = 0;
    {
      Timed tk(c, 8);
      tc_launch(c, a, mode, c->d_Racc);
      if (miss) {
        a.Bimg = c->d_BimgD;
        tc_launch(c, a, tc::kMask, c->d_Racc2);
      }
    }
    double* Fw = reinterpret_cast<double*>(c->d_tcs + 4 * c->lp);
    // Fw: reduced by the last block of the slice kernel for window-sized launches, by its own
    // kernel otherwise (folding the sum into every finish block was tried: the serial chain of a
    // window's ~250 partials per block cost ~30 us per launch)
    const bool fold_fw = false;
    if (nkb_w > kFoldFwMaxParts) {
      tc::k_tc_reduce_fpart<<<c->l, 256, 0, c->stream>>>(c->d_Fpart, nkb_w, c->l, c->lp, Fw);
      PCA_CHECK_LAUNCH();
      c->tm.kernel_launches++;
    }
    const bool fuse_sum = c->sum_out != nullptr && c->sum_other != nullptr;
    tc::k_tc_finish_h<<<grid_for((c->N * c->lp + 3) / 4, 256, c->sms), 256, 0, c->stream>>>(  // 4 elements per thread
        c->d_Racc, miss ? c->d_Racc2 : nullptr, c->N, c->l, c->lp, c->slices, w_csum, w_colmax, Fw,
        fold_fw ? c->d_Fpart : nullptr, nkb_w, Hacc, accumulate ? 1 : 0, fuse_sum ? c->sum_other : nullptr,
        fuse_sum ? c->sum_out : nullptr);
    PCA_CHECK_LAUNCH();
    if (fuse_sum) c->sum_done = true;
    c->tm.kernel_launches++;
    c->tm.gemm_h_launches++;
  }
  c->tc_ranges++;
  if (miss) c->tc_miss_ranges++;
}

// ---------------------------------------------------------------- generic dense matrix (RsvdOpOnePass)
template <int NT>
void dense_g_nt(pcaone_ctx* c, const double* D, uint32_t nrows, double* G) {
  const size_t smem = 2 * DenseSmem<NT>::kStageG;
  static bool attr = false;
  if (!attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_dense_g<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  k_dense_g<NT><<<ceil_div(nrows, kDenseRows), kDenseThreads, smem, c->stream>>>(D, c->ldd, nrows, (uint32_t)c->N,
                                                                                   c->d_Omg, G);
  PCA_CHECK_LAUNCH();
}
template <int NT>
void dense_h_nt(pcaone_ctx* c, const double* D, uint32_t nrows, const double* G, uint32_t splits, uint32_t rps) {
  const size_t smem = 2 * DenseSmem<NT>::kStageH;
  static bool attr = false;
  if (!attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_dense_h<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
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
  static bool attr = false;
  if (!attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_dos_g<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  k_dos_g<NT><<<ceil_div(nrows, kDenseRows), kDenseThreads, smem, c->stream>>>(D, c->ldf, nrows, (uint32_t)c->N, F,
                                                                                 c->lut, c->d_Omg, G);
  PCA_CHECK_LAUNCH();
}
template <int NT>
void dos_h_nt(pcaone_ctx* c, const float* D, uint32_t nrows, const double* F, const double* G, uint32_t splits,
              uint32_t rps) {
  const size_t smem = 2 * DenseSmem<NT>::kDosStageH;
  static bool attr = false;
  if (!attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_dos_h<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
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

// G rows of the range = X^T Omega ; Hacc (+)= X G. `buf` = streamed block buffer holding P, or -1
// when P points into the resident shard. Ranges without missing genotypes (and no EMU fill) run
// on the int8 tensor-core kernels when the context was created with a PCAONE_PREC_INT8* mode.
void range_gemms(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, uint64_t snp0, double* Hacc, bool accumulate,
                 int buf) {
  if (nrows == 0) return;
  if (c->source == PCAONE_SRC_DENSE || c->source == PCAONE_SRC_GL) {
    range_gemms_dense(c, snp0, nrows, Hacc, accumulate);
    return;
  }
  if (c->source == PCAONE_SRC_DOSAGE) {
    range_gemms_dosage(c, snp0, nrows, Hacc, accumulate);
    return;
  }
  // EMU update passes fill every missing entry with its own FP64 value: FP64 kernels
  const bool use_tc = c->slices > 0 && !(c->update && c->cfg.emu) && c->half == 3;  // half passes: FP64 kernels
  bool has_miss = false;
  if (use_tc) {
    uint64_t miss = tc_missing_in(c, snp0, nrows);
    if (miss == UINT64_MAX) {
      tc_fetch_nmiss(c, snp0, nrows);
      miss = tc_missing_in(c, snp0, nrows);
    }
    has_miss = miss != 0;
  }
  if (!use_tc) {
    range_gemms_fp64(c, P, nrows, snp0, Hacc, accumulate);
    c->fp64_ranges++;
    return;
  }
  tc_alloc(c, std::max<uint64_t>(nrows, c->max_block), has_miss);
  if (buf < 0) {
    if (!c->tiles_valid) {
      if (!c->d_PG) {
        PCA_CUDA(cudaMalloc((void**)&c->d_PG, tc_pg_bytes(c, c->M)));
        PCA_CUDA(cudaMalloc((void**)&c->d_PH, tc_ph_bytes(c, c->M)));
      }
      tc_build_tiles(c, c->d_packed, c->M, c->d_PG, c->d_PH, c->stream);
      c->tiles_valid = true;
    }
    range_gemms_tc(c, c->d_PG, c->d_PH, snp0, nrows, snp0, Hacc, accumulate, has_miss);
  } else {
    if (!c->d_PGb[buf]) {
      PCA_CUDA(cudaMalloc((void**)&c->d_PGb[buf], tc_pg_bytes(c, c->max_block)));
      PCA_CUDA(cudaMalloc((void**)&c->d_PHb[buf], tc_ph_bytes(c, c->max_block)));
    }
    tc_build_tiles(c, P, nrows, c->d_PGb[buf], c->d_PHb[buf], c->stream);
    range_gemms_tc(c, c->d_PGb[buf], c->d_PHb[buf], 0, nrows, snp0, Hacc, accumulate, has_miss);
  }
}

// ---------------------------------------------------------------- tall-skinny helpers
template <int R>
void ts_gemm_r(pcaone_ctx* c, const double* A, int l1, const double* B, int l2, uint64_t rows, int nparts,
               uint64_t rpc) {
  k_ts_gemm_tn<R, R><<<nparts, kTsThreads, 0, c->stream>>>(A, c->lp, l1, B, c->lp, l2, rows, rpc, c->d_part, c->lp);
}

// C (l1 x l2, ld lp) = A^T B over `rows` rows (both [rows][lp]); optional allreduce for sharded rows
void ts_gemm_tn(pcaone_ctx* c, const double* A, int l1, const double* B, int l2, uint64_t rows, double* C,
                bool sharded_rows) {
  const int R = (std::max(l1, l2) + 15) / 16;
  uint64_t rpc = std::max<uint64_t>(kTsKR, round_up((rows + 2 * c->sms - 1) / (2 * c->sms), kTsKR));
  int nparts = (int)std::max<uint64_t>(1, (rows + rpc - 1) / rpc);
  const size_t need = (size_t)nparts * 16 * R * c->lp;
  if (need > c->part_doubles) throw std::runtime_error("partial workspace too small");
  switch (R) {
    case 1: ts_gemm_r<1>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 2: ts_gemm_r<2>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 3: ts_gemm_r<3>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 4: ts_gemm_r<4>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 5: ts_gemm_r<5>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 6: ts_gemm_r<6>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 7: ts_gemm_r<7>(c, A, l1, B, l2, rows, nparts, rpc); break;
    case 8: ts_gemm_r<8>(c, A, l1, B, l2, rows, nparts, rpc); break;
    default: throw std::runtime_error("l too large for ts_gemm");
  }
  PCA_CHECK_LAUNCH();
  k_reduce_small<<<ceil_div(l1 * l2, 256), 256, 0, c->stream>>>(c->d_part, nparts, 16 * R * c->lp, l1, l2, c->lp, C);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches += 2;
  if (sharded_rows && c->cfg.world > 1) {
    if (!c->allreduce) throw std::runtime_error("world > 1 but no allreduce hook installed");
    Timed t(c, 5);
    if (c->allreduce(c->allreduce_user, C, (uint64_t)c->lp * c->lp, c->stream))
      throw std::runtime_error("allreduce hook failed");
  }
}

template <int RN>
void rightmult_r(pcaone_ctx* c, const double* A, int l1, const double* T, int l2, uint64_t rows, double* Out) {
  const size_t smem = ((size_t)l1 * 16 * RN + (size_t)64 * (l1 + 1)) * sizeof(double);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_ts_rightmult<RN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int grid = (int)std::min<uint64_t>((rows + 63) / 64, (uint64_t)c->sms * 4);
  k_ts_rightmult<RN><<<grid, kTsThreads, smem, c->stream>>>(A, c->lp, l1, T, c->lp, l2, rows, Out, c->lp);
}

// Out[rows][lp] = A[rows][:l1] * T[l1 x l2]
void ts_rightmult(pcaone_ctx* c, const double* A, int l1, const double* T, int l2, uint64_t rows, double* Out) {
  const int RN = (l2 + 15) / 16;
  switch (RN) {
    case 1: rightmult_r<1>(c, A, l1, T, l2, rows, Out); break;
    case 2: rightmult_r<2>(c, A, l1, T, l2, rows, Out); break;
    case 3: rightmult_r<3>(c, A, l1, T, l2, rows, Out); break;
    case 4: rightmult_r<4>(c, A, l1, T, l2, rows, Out); break;
    case 5: rightmult_r<5>(c, A, l1, T, l2, rows, Out); break;
    case 6: rightmult_r<6>(c, A, l1, T, l2, rows, Out); break;
    case 7: rightmult_r<7>(c, A, l1, T, l2, rows, Out); break;
    case 8: rightmult_r<8>(c, A, l1, T, l2, rows, Out); break;
    default: throw std::runtime_error("l too large for rightmult");
  }
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

int read_status(pcaone_ctx* c) {
  PCA_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  return c->h_status[0];
}

void small_matmul(pcaone_ctx* c, const double* A, int tA, const double* B, int tB, int m, int p, int n, double* C) {
  k_small_matmul<<<1, 1024, 0, c->stream>>>(A, tA, B, tB, m, p, n, c->lp, C);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

void jacobi(pcaone_ctx* c, const double* A, int sym, double* sigma, double* V) {
  const size_t smem = 2 * (size_t)c->l * c->l * sizeof(double);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_jacobi_svd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  static int sw_left = getenv("PCAONE_SMALL_PROF") ? atoi(getenv("PCAONE_SMALL_PROF")) : 0;
  k_jacobi_svd<<<1, kSmallThreads, smem, c->stream>>>(A, c->l, c->lp, sym, sigma, V, sw_left > 0 ? c->d_status + 2 : nullptr);
  if (sw_left > 0) {
    --sw_left;
    int sw = 0;
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    PCA_CUDA(cudaMemcpy(&sw, c->d_status + 2, sizeof(int), cudaMemcpyDeviceToHost));
    fprintf(stderr, "jacobi sweeps: %d\n", sw);
  }
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

void launch_chol(pcaone_ctx* c, const double* W, double* R, double* Rinv) {
  const size_t smem = (size_t)c->l * c->l * sizeof(double);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_chol_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  k_chol_inv<<<1, kSmallThreads, smem, c->stream>>>(W, c->l, c->lp, R, Rinv, c->d_status);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// One orthonormalising factor from the Gram W of A: Tout (l x l) with A*Tout having orthonormal
// columns. Cholesky (CholeskyQR) when W is numerically full rank, else the eigen route (SVQB)
// which zeroes the null directions.
void gram_factor(pcaone_ctx* c, const double* W, double* Tout) {
  launch_chol(c, W, c->d_R, Tout);
  if (read_status(c) != 0) {
    jacobi(c, W, 1, c->d_sigma, c->d_Vr);
    k_svqb_factor<<<1, 1024, 0, c->stream>>>(c->d_Vr, c->d_sigma, c->l, c->lp, Tout);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
  }
}

template <int R>
void orth_fused_r(pcaone_ctx* c, OrthArgs& a) {
  const size_t smem = orth_smem_bytes(c->l, R);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_orth_fused<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  void* args[] = {(void*)&a};
  PCA_CUDA(cudaLaunchCooperativeKernel((void*)k_orth_fused<R>, dim3(c->sms), dim3(kOrthThreads), args, smem, c->stream));
  c->tm.kernel_launches++;
}

bool orth_fused_ok(const pcaone_ctx* c) { return c->fused_orth && c->l <= kOrthMaxL; }

// One cooperative launch: Q = orth(A) (CholeskyQR2) [+ Householder signs] [+ flipOmg against Q2].
void orth_fused(pcaone_ctx* c, const double* A, uint64_t rows, double* Q, double* Q2, double* Ttot, bool signs,
                bool flip, int phases = 7, unsigned long long* colmax_out = nullptr) {
  OrthArgs a{};
  a.phases = phases;
  a.colmax_out = colmax_out;
  a.A = A;
  a.Q = Q;
  a.Q2 = Q2;
  a.rows = rows;
  a.l = c->l;
  a.lp = c->lp;
  a.want_signs = signs ? 1 : 0;
  a.want_flip = flip ? 1 : 0;
  a.part = c->d_part;
  a.Wg = c->d_W;
  a.T1g = c->d_T1;
  a.T2g = c->d_T2;
  a.Ttot = Ttot;
  a.hsign = c->d_hsign;
  a.fsign = c->d_sign;
  a.jscratch = c->d_jscratch;
  a.status = c->d_status + 1;
  // QR(G) of the dense stage (factors only; single launch or the row-sharded three-launch form): the second Cholesky pass is
  // dropped when the first one shows cond_F(G)^2 <= 1e5 (PCAONE_QR2_ALWAYS=1 keeps it)
  static const bool qr2_always = getenv("PCAONE_QR2_ALWAYS") && atoi(getenv("PCAONE_QR2_ALWAYS")) != 0;
  a.skip2 = (!Q && (phases == 7 || phases == 2) && !qr2_always) ? c->d_status + 3 : nullptr;
  a.skip_diag = c->cfg.rank == 0 ? 1.0 : 0.0;
  static unsigned long long* d_prof = nullptr;
  static int prof_left = getenv("PCAONE_ORTH_PROF") ? atoi(getenv("PCAONE_ORTH_PROF")) : 0;
  if (prof_left > 0) {
    if (!d_prof) PCA_CUDA(cudaMalloc((void**)&d_prof, 64 * sizeof(unsigned long long)));
    PCA_CUDA(cudaMemsetAsync(d_prof, 0, 64 * sizeof(unsigned long long), c->stream));
    a.prof = d_prof;
  }
  if ((size_t)c->sms * c->l * c->lp > c->part_doubles) throw std::runtime_error("partial workspace too small");
  switch ((c->l + 15) / 16) {
    case 1: orth_fused_r<1>(c, a); break;
    case 2: orth_fused_r<2>(c, a); break;
    case 3: orth_fused_r<3>(c, a); break;
    case 4: orth_fused_r<4>(c, a); break;
    case 5: orth_fused_r<5>(c, a); break;
    default: throw std::runtime_error("orth_fused: l too large");
  }
  if (prof_left > 0) {
    --prof_left;
    unsigned long long h[64];
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    PCA_CUDA(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "orth_fused rows=%llu phases(us):", (unsigned long long)rows);
    const int np = (int)std::min<unsigned long long>(h[63], 62);
    for (int i = 1; i < np; ++i) fprintf(stderr, " %.1f", (double)(h[i] - h[i - 1]) * 1e-3);
    fprintf(stderr, "\n");
  }
}

// Q = orth(A) in two passes (CholeskyQR2); Q may alias A. Ttot (optional) = T1*T2, Q = A*Ttot.
// Q == nullptr: only Ttot is wanted (the caller applies it later); returns false if Q was not formed.
bool orth2(pcaone_ctx* c, const double* A, uint64_t rows, double* Q, double* Ttot, bool sharded_rows,
           bool factors_only = false) {
  if (orth_fused_ok(c) && !(sharded_rows && c->cfg.world > 1)) {
    orth_fused(c, A, rows, factors_only ? nullptr : Q, nullptr, Ttot, false, false);
    return !factors_only;
  }
  if (orth_fused_ok(c)) {
    // rows sharded across ranks: the same kernel in three launches, the two l x l Gram matrices
    // summed over the ranks in between (the allreduce hook cannot be called from inside a kernel)
    auto reduce_W = [&]() {
      if (!c->allreduce) throw std::runtime_error("world > 1 but no allreduce hook installed");
      Timed t(c, 5);
      if (c->allreduce(c->allreduce_user, c->d_W, (uint64_t)c->l * c->lp, c->stream))
        throw std::runtime_error("allreduce hook failed");
    };
    double* Qo = factors_only ? nullptr : Q;
    orth_fused(c, A, rows, Qo, nullptr, Ttot, false, false, 1);
    reduce_W();
    orth_fused(c, A, rows, Qo, nullptr, Ttot, false, false, 2);
    reduce_W();
    orth_fused(c, A, rows, Qo, nullptr, Ttot, false, false, 4);
    return !factors_only;
  }
  ts_gemm_tn(c, A, c->l, A, c->l, rows, c->d_W, sharded_rows);
  gram_factor(c, c->d_W, c->d_T1);
  ts_rightmult(c, A, c->l, c->d_T1, c->l, rows, Q);
  ts_gemm_tn(c, Q, c->l, Q, c->l, rows, c->d_W, sharded_rows);
  gram_factor(c, c->d_W, c->d_T2);
  ts_rightmult(c, Q, c->l, c->d_T2, c->l, rows, Q);
  if (Ttot) small_matmul(c, c->d_T1, 0, c->d_T2, 0, c->l, c->l, c->l, Ttot);
  return true;
}

void flip_omg(pcaone_ctx* c, const double* pre) {
  uint64_t rpc = std::max<uint64_t>(8, (c->N + c->sms - 1) / c->sms);
  int nparts = (int)((c->N + rpc - 1) / rpc);
  if ((size_t)nparts * 2 * c->l > c->part_doubles) throw std::runtime_error("partial workspace too small");
  k_flip_partial<<<nparts, 256, 0, c->stream>>>(c->d_Omg2, c->d_Omg, c->lp, c->l, c->N, rpc, pre, c->d_part);
  PCA_CHECK_LAUNCH();
  k_flip_sign<<<1, 128, 0, c->stream>>>(c->d_part, nparts, c->l, pre, c->d_sign);
  PCA_CHECK_LAUNCH();
  k_flip_apply<<<grid_for(c->N * c->lp, 256, c->sms), 256, 0, c->stream>>>(c->d_Omg, c->d_Omg2, c->lp, c->l, c->N,
                                                                          c->d_sign);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches += 3;
}

void allreduce_H(pcaone_ctx* c, double* H) {
  if (c->cfg.world > 1) {
    if (!c->allreduce) throw std::runtime_error("world > 1 but no allreduce hook installed");
    Timed t(c, 5);
    if (c->allreduce(c->allreduce_user, H, c->N * c->lp, c->stream)) throw std::runtime_error("allreduce hook failed");
  }
}

// Omega = thinQ(H) (+ flipOmg)   Halko.cpp:120-124 / 208-213
void update_omega(pcaone_ctx* c, const double* H, bool flip) {
  Timed t(c, 2);
  if (orth_fused_ok(c)) {
    unsigned long long* cm = nullptr;
    if (c->slices > 0 && c->d_tcs) {  // int8 route: the kernel also leaves max |Omega| per column for the slicing
      cm = c->d_tcs;  // [0, 2 lp): column maxima + column sums of Omega, cleared by the kernel itself
    }
    orth_fused(c, H, c->N, c->d_Omg, flip ? c->d_Omg2 : nullptr, nullptr, true, flip, 7, cm);
    c->tm.omega_updates++;
    c->omega_img_valid = false;
    c->omega_colmax_valid = cm != nullptr;
    return;
  }
  orth2(c, H, c->N, c->d_Omg, nullptr, false);
  // give the CholeskyQR basis the column signs of the reference's Householder thin Q
  const size_t smem = (size_t)c->l * c->l * sizeof(double);
  static size_t attr = 0;
  if (smem > attr) {
    PCA_CUDA(cudaFuncSetAttribute(k_householder_signs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  k_householder_signs<<<1, kSmallThreads, smem, c->stream>>>(c->d_Omg, c->l, c->lp, c->d_hsign);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
  if (flip) {
    flip_omg(c, c->d_hsign);
  } else {
    k_scale_cols<<<grid_for(c->N * c->l, 256, c->sms), 256, 0, c->stream>>>(c->d_Omg, c->lp, c->l, c->N, c->d_hsign);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
  }
  c->tm.omega_updates++;
  c->omega_img_valid = c->omega_colmax_valid = false;
}

// ---------------------------------------------------------------- host <-> device matrices
void ensure_stage(pcaone_ctx* c, size_t doubles) {
  if (doubles > c->stage_doubles) {
    if (c->d_stage) cudaFree(c->d_stage);
    dmalloc(&c->d_stage, doubles);
    c->stage_doubles = doubles;
  }
}
void upload_colmajor(pcaone_ctx* c, const double* h, uint64_t rows, int cols, double* d) {
  ensure_stage(c, rows * cols);
  PCA_CUDA(cudaMemcpyAsync(c->d_stage, h, rows * cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->tm.h2d_bytes += rows * cols * sizeof(double);
  dim3 blk(32, 8);
  k_colmajor_to_rowmajor<<<ceil_div(rows, 32), blk, 0, c->stream>>>(c->d_stage, rows, cols, d, c->lp);
  PCA_CHECK_LAUNCH();
  PCA_CUDA(cudaStreamSynchronize(c->stream));
}
void download_colmajor(pcaone_ctx* c, const double* d, uint64_t rows, int cols, double* h) {
  ensure_stage(c, rows * cols);
  dim3 blk(32, 8);
  k_rowmajor_to_colmajor<<<ceil_div(rows, 32), blk, 0, c->stream>>>(d, c->lp, rows, cols, c->d_stage);
  PCA_CHECK_LAUNCH();
  PCA_CUDA(cudaMemcpyAsync(h, c->d_stage, rows * cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  c->tm.d2h_bytes += rows * cols * sizeof(double);
}

// ---------------------------------------------------------------- block streaming
void alloc_stream_buffers(pcaone_ctx* c) {
  if (c->d_blk[0] || c->max_block == 0) return;
  for (int i = 0; i < 2; ++i) {
    dmalloc(&c->d_blk[i], c->max_block * c->pitch);
    if (c->pitch != c->bpr) dmalloc(&c->d_raw[i], c->max_block * c->bpr);
    if (c->source == PCAONE_SRC_FILE)
      PCA_CUDA(cudaHostAlloc((void**)&c->h_pin[i], c->max_block * c->bpr, cudaHostAllocDefault));
    PCA_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    PCA_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  }
}

// enqueue the H2D of block b into buffer `buf`; returns the device pointer (pitch layout)
const uint8_t* stage_block(pcaone_ctx* c, uint32_t b, int buf) {
  const uint64_t s = c->blk_start[b], e = c->blk_stop[b];
  const uint64_t nrows = e - s + 1;
  const size_t bytes = nrows * c->bpr;
  PCA_CUDA(cudaEventSynchronize(c->ev_done[buf]));  // previous user of this buffer finished
  const uint8_t* src;
  if (c->source == PCAONE_SRC_HOST) {
    src = c->h_packed + s * c->bpr;
  } else {
    if (c->reader) {
      if (c->reader(c->reader_user, s, e, c->h_pin[buf])) throw std::runtime_error("block reader failed");
    } else {
      const long long off = 3 + (long long)(c->bed_snp_offset + s) * c->bpr;
      if (fseeko(c->bed_file, off, SEEK_SET) != 0 || fread(c->h_pin[buf], 1, bytes, c->bed_file) != bytes)
        throw std::runtime_error("read_block: short read from bed file");
    }
    src = c->h_pin[buf];
  }
  uint8_t* dst = (c->pitch != c->bpr) ? c->d_raw[buf] : c->d_blk[buf];
  PCA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
  c->tm.h2d_bytes += bytes;
  if (c->pitch != c->bpr) {
    k_repitch<<<grid_for(nrows * (c->pitch >> 4), 256, c->sms), 256, 0, c->copy_stream>>>(c->d_raw[buf], c->d_blk[buf],
                                                                                         nrows, c->bpr, c->pitch);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
  }
  PCA_CUDA(cudaEventRecord(c->ev_copied[buf], c->copy_stream));
  PCA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
  return c->d_blk[buf];
}

void block_af_if_needed(pcaone_ctx* c, const uint8_t* P, uint64_t s, uint64_t nrows) {
  if (c->af_done) return;
  k_allele_freq<<<grid_for(nrows * 32, 256, c->sms), 256, 0, c->stream>>>(P, c->pitch, (uint32_t)c->N, nrows,
                                                                          c->d_F + s, c->d_nmiss + s);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// ---------------------------------------------------------------- the passes
struct WinStep {
  uint64_t start, stop;  // inclusive SNP range (local), empty if stop < start
  int target;            // 1 -> H1, 2 -> H2
  bool update;
  bool zero_h1;  // after the update: i == bandsize -> zero H1 (and i = 0), else zero H2
};

// FancyRsvdOpData::computeGandH window state machine (Halko.cpp:188-222 in-core, :228-267 OOC)
std::vector<WinStep> winsvd_schedule(pcaone_ctx* c, int pi, const std::vector<uint64_t>& ws,
                                     const std::vector<uint64_t>& we, bool ooc) {
  const uint64_t bands = c->cfg.bands;
  const uint64_t nwin = ws.size();
  if (pi == 0) c->bandsize = ooc ? c->band_factor : 1;
  c->bandsize = std::min<uint64_t>(c->bandsize * 2, ooc ? nwin : bands);
  const uint64_t bandsize = c->bandsize;
  std::vector<WinStep> steps;
  uint64_t i = 1;
  for (uint64_t b = 0; b < nwin; ++b, ++i) {
    WinStep st{ws[b], we[b], (i <= bandsize / 2) ? 1 : 2, false, false};
    const double adj_at = pi > 0 ? std::pow(2.0, pi - 1) * (ooc ? c->band_factor : 1) : -1.0;
    const bool adjacent = (pi > 0 && (double)(b + 1) == adj_at && std::pow(2.0, pi) < (double)bands);
    if (!((b + 1) < bandsize && !adjacent)) {
      if ((i == bandsize) || (i == bandsize / 2) || adjacent) {
        st.update = true;
        st.zero_h1 = (i == bandsize);
        if (i == bandsize) i = 0;
      }
    }
    steps.push_back(st);
  }
  return steps;
}

void incore_windows(pcaone_ctx* c, std::vector<uint64_t>& ws, std::vector<uint64_t>& we) {
  // Halko.cpp:180,192-194: blocksize = ceil(M / bands) on the WHOLE job's SNP count; a shard
  // walks its 1/world slice of every window, i.e. the same formula on local M.
  const uint64_t bands = c->cfg.bands;
  const uint64_t bs = (c->M + bands - 1) / bands;
  for (uint64_t b = 0; b < bands; ++b) {
    uint64_t s = b * bs;
    uint64_t e = ((b + 1) * bs >= c->M) ? c->M - 1 : (b + 1) * bs - 1;
    if (s >= c->M) {  // empty trailing window (M < bands * blocksize)
      s = 1;
      e = 0;
    }
    ws.push_back(s);
    we.push_back(e);
  }
}

void zero_async(pcaone_ctx* c, double* p, uint64_t n) { PCA_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), c->stream)); }

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

void compute_gandh(pcaone_ctx* c, int pi) {
  if (c->source < 0) throw std::runtime_error("no genotype source set");
  if (c->update && c->cfg.emu && !c->have_usv) throw std::runtime_error("EMU update pass without U,S,V");
  const bool ooc = c->source == PCAONE_SRC_HOST || c->source == PCAONE_SRC_FILE;
  const bool win = c->cfg.svd == PCAONE_SVD_WINSVD;
  c->lut.standardize = (c->standardize && c->cfg.scale == -9) ? 1 : 0;
  const uint64_t HN = c->N * c->lp;
  if (pi == 0) {  // initOmg(): every computeUSV restarts from the same seeded Omega (Halko.cpp:105,160)
    if (!c->have_omg0) throw std::runtime_error("call pcaone_set_omega before the first pass");
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg, c->d_Omg0, HN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg2, c->d_Omg0, HN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->omega_img_valid = c->omega_colmax_valid = false;
  }
  if (ooc) {
    if (c->blk_start.empty()) throw std::runtime_error("out-of-core source needs pcaone_set_blocks");
    alloc_stream_buffers(c);
  } else if (!c->af_done && c->source != PCAONE_SRC_DENSE) {
    throw std::runtime_error("call pcaone_allele_freq before the first pass");
  }
  if (c->source == PCAONE_SRC_GL && pi == 0) {
    Timed t(c, 6);
    gl_refresh(c, 0, c->M, c->update != 0, c->d_dense, c->ldd);
  }

  if (!win) {
    // ---- sSVD, Halko.cpp:99-153
    if (pi > 0) {
      // in-core flips (Halko.cpp:123); the block version does not (Halko.cpp:133-136).
      // A resident shard walked with a block plan follows cfg.out_of_core.
      update_omega(c, c->d_H, !c->cfg.out_of_core);
    }
    if (!ooc) {
      if (c->blk_start.empty()) {
        range_gemms(c, c->d_packed, (uint32_t)c->M, 0, c->d_H, false, -1);
      } else {
        zero_async(c, c->d_H, HN);
        for (size_t b = 0; b < c->blk_start.size(); ++b)
          range_gemms(c, c->d_packed + c->blk_start[b] * c->pitch, (uint32_t)(c->blk_stop[b] - c->blk_start[b] + 1),
                      c->blk_start[b], c->d_H, true, -1);
      }
    } else {
      zero_async(c, c->d_H, HN);
      for (uint32_t b = 0; b < c->blk_start.size(); ++b) {
        const int buf = b & 1;
        const uint8_t* P = stage_block(c, b, buf);
        const uint64_t nrows = c->blk_stop[b] - c->blk_start[b] + 1;
        block_af_if_needed(c, P, c->blk_start[b], nrows);
        range_gemms(c, P, (uint32_t)nrows, c->blk_start[b], c->d_H, true, buf);
        PCA_CUDA(cudaEventRecord(c->ev_done[buf], c->stream));
      }
      c->af_done = true;
    }
    allreduce_H(c, c->d_H);
    return;
  }

  // ---- winSVD, Halko.cpp:155-269
  if (std::pow(2.0, pi) >= (double)c->cfg.bands) {
    zero_async(c, c->d_H1, HN);
    zero_async(c, c->d_H2, HN);
  }
  std::vector<uint64_t> ws, we;
  const bool block_walk = ooc || !c->blk_start.empty();
  if (block_walk) {
    ws = c->blk_start;
    we = c->blk_stop;
  } else {
    incore_windows(c, ws, we);
  }
  auto steps = winsvd_schedule(c, pi, ws, we, c->cfg.out_of_core != 0);
  size_t b = 0;
  while (b < steps.size()) {
    // merge resident windows that share a target and have no Omega update between them
    size_t e = b;
    if (!ooc) {
      while (!steps[e].update && e + 1 < steps.size() && steps[e + 1].target == steps[b].target &&
             steps[e + 1].stop >= steps[e + 1].start && steps[e].stop >= steps[e].start &&
             steps[e + 1].start == steps[e].stop + 1)
        ++e;
    }
    double* Hacc = steps[b].target == 1 ? c->d_H1 : c->d_H2;
    const WinStep& last = steps[e];
    if (last.update) {  // the range's H finish may also form H = H1 + H2 for the update (int8 route)
      c->sum_other = steps[b].target == 1 ? c->d_H2 : c->d_H1;
      c->sum_out = c->d_H;
    }
    c->sum_done = false;
    if (steps[b].stop >= steps[b].start) {
      const uint64_t s0 = steps[b].start, nrows = steps[e].stop - s0 + 1;
      if (!ooc) {
        range_gemms(c, c->d_packed + s0 * c->pitch, (uint32_t)nrows, s0, Hacc, true, -1);
      } else {
        const int buf = (int)(b & 1);
        const uint8_t* P = stage_block(c, (uint32_t)b, buf);
        block_af_if_needed(c, P, s0, nrows);
        range_gemms(c, P, (uint32_t)nrows, s0, Hacc, true, buf);
        PCA_CUDA(cudaEventRecord(c->ev_done[buf], c->stream));
      }
    }
    c->sum_other = nullptr;
    c->sum_out = nullptr;
    if (last.update) {
      if (!c->sum_done) {
        k_add2<<<grid_for(HN, 256, c->sms), 256, 0, c->stream>>>(c->d_H1, c->d_H2, c->d_H, HN);
        PCA_CHECK_LAUNCH();
        c->tm.kernel_launches++;
      }
      allreduce_H(c, c->d_H);
      update_omega(c, c->d_H, true);
      zero_async(c, last.zero_h1 ? c->d_H1 : c->d_H2, HN);
    }
    b = e + 1;
  }
  if (ooc) c->af_done = true;
}

// Halko.cpp:55-70 on the device. Leaves: G <- Q2, d_Ucur (N x k), d_sigma (l), d_Vr = U_B (l x l)
void small_stage(pcaone_ctx* c) {
  Timed t(c, 3);
  // optional per-step breakdown (debug aid): PCAONE_SMALL_PROF=n prints the first n calls
  static int prof_left = getenv("PCAONE_SMALL_PROF") ? atoi(getenv("PCAONE_SMALL_PROF")) : 0;
  cudaEvent_t ev[10];
  int nev = 0;
  const bool prof = prof_left > 0;
  auto mark = [&]() {
    if (!prof) return;
    PCA_CUDA(cudaEventCreate(&ev[nev]));
    PCA_CUDA(cudaEventRecord(ev[nev], c->stream));
    ++nev;
  };
  mark();
  // G = Q R twice (CholeskyQR2); T = R^-1 so that Q = G T and B^T = H R^-1 = H T
  // Only T is needed per epoch; Q itself enters the result once, as V = Q U_B (Halko.cpp:89), which
  // finalize_usv forms as G (T U_B): the M x l matrix Q is never written.
  c->g_is_q = orth2(c, c->d_G, c->M, c->d_G, c->d_T, true, true);
  mark();
  ts_rightmult(c, c->d_H, c->l, c->d_T, c->l, c->N, c->d_Bt);
  // SVD of B^T (N x l): Gram -> Cholesky -> one-sided Jacobi on the triangular factor
  ts_gemm_tn(c, c->d_Bt, c->l, c->d_Bt, c->l, c->N, c->d_W, false);
  launch_chol(c, c->d_W, c->d_R, c->d_Rinv);
  mark();
  const int st = read_status(c);
  mark();
  if (st == 0)
    jacobi(c, c->d_R, 0, c->d_sigma, c->d_Vr);
  else
    jacobi(c, c->d_W, 1, c->d_sigma, c->d_Vr);
  mark();
  k_scale_v_by_inv_sigma<<<1, 1024, 0, c->stream>>>(c->d_Vr, c->d_sigma, c->l, c->k, c->lp, c->d_Z);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
  ts_rightmult(c, c->d_Bt, c->l, c->d_Z, c->k, c->N, c->d_Ucur);
  mark();
  if (prof) {
    --prof_left;
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    fprintf(stderr, "small_stage (ms): orth(G)");
    const char* names[] = {"", " Bt+Gram+chol", " status-sync", " jacobi", " scale+Ucur"};
    for (int i = 1; i < nev; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, "%s %.3f", names[i - 1], ms);
    }
    fprintf(stderr, "\n");
    for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
  }
}

double device_mev(pcaone_ctx* c, const double* X, const double* Y, uint64_t rows, bool sharded) {
  ts_gemm_tn(c, X, c->k, Y, c->k, rows, c->d_W, sharded);
  k_mev_from_xty<<<1, 32, 0, c->stream>>>(c->d_W, c->k, c->lp, c->d_scal);
  PCA_CHECK_LAUNCH();
  PCA_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  c->tm.kernel_launches += 1;
  return c->h_scal[0];
}

void finalize_usv(pcaone_ctx* c) {
  const uint64_t bytes = c->N * c->lp * sizeof(double);
  PCA_CUDA(cudaMemcpyAsync(c->d_U, c->d_Ucur, bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (c->g_is_q) {
    ts_rightmult(c, c->d_G, c->l, c->d_Vr, c->k, c->M, c->d_V);
  } else {
    small_matmul(c, c->d_T, 0, c->d_Vr, 0, c->l, c->l, c->k, c->d_Z);
    ts_rightmult(c, c->d_G, c->l, c->d_Z, c->k, c->M, c->d_V);
  }
  PCA_CUDA(cudaMemcpyAsync(c->d_S, c->d_sigma, c->k * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  c->have_usv = true;
}

// RsvdOpData::computeUSV, Halko.cpp:46-97
void compute_usv(pcaone_ctx* c, int p, double tol) {
  double diff = 0.0;
  int epochs = 0;
  const uint64_t ubytes = c->N * c->lp * sizeof(double);
  for (int pi = 0; pi <= p; ++pi) {
    compute_gandh(c, pi);
    small_stage(c);
    epochs = pi + 1;
    if (pi > 0) {
      diff = 1.0 - device_mev(c, c->d_Ucur, c->d_Upre, c->N, false);
      if (diff < tol || pi == p) {
        if (c->cfg.svd == PCAONE_SVD_WINSVD && std::pow(2.0, pi) < (double)c->cfg.bands) {
          p = (int)std::log2((double)c->cfg.bands);
        } else {
          finalize_usv(c);
          break;
        }
      } else {
        PCA_CUDA(cudaMemcpyAsync(c->d_Upre, c->d_Ucur, ubytes, cudaMemcpyDeviceToDevice, c->stream));
      }
    } else {
      PCA_CUDA(cudaMemcpyAsync(c->d_Upre, c->d_Ucur, ubytes, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  PCA_CUDA(cudaMemcpyAsync(c->d_U, c->d_Ucur, ubytes, cudaMemcpyDeviceToDevice, c->stream));
  c->last_diff = diff;
  c->last_epochs = epochs;
  PCA_CUDA(cudaStreamSynchronize(c->stream));
}

// One pass over every SNP of the source in plan order with the CURRENT Omega / d_G: G rows of each
// range from Omega, d_H = sum over ranges (the sSVD pass of compute_gandh without the Omega update).
void walk_ranges(pcaone_ctx* c) {
  const bool ooc = c->source == PCAONE_SRC_HOST || c->source == PCAONE_SRC_FILE;
  const uint64_t HN = c->N * c->lp;
  if (!ooc) {
    if (!c->af_done && c->source != PCAONE_SRC_DENSE) throw std::runtime_error("call pcaone_allele_freq first");
    if (c->blk_start.empty()) {
      range_gemms(c, c->d_packed, (uint32_t)c->M, 0, c->d_H, false, -1);
    } else {
      zero_async(c, c->d_H, HN);
      for (size_t b = 0; b < c->blk_start.size(); ++b)
        range_gemms(c, c->d_packed + c->blk_start[b] * c->pitch, (uint32_t)(c->blk_stop[b] - c->blk_start[b] + 1),
                    c->blk_start[b], c->d_H, true, -1);
    }
  } else {
    if (c->blk_start.empty()) throw std::runtime_error("out-of-core source needs pcaone_set_blocks");
    alloc_stream_buffers(c);
    zero_async(c, c->d_H, HN);
    for (uint32_t b = 0; b < c->blk_start.size(); ++b) {
      const int buf = b & 1;
      const uint8_t* P = stage_block(c, b, buf);
      const uint64_t nrows = c->blk_stop[b] - c->blk_start[b] + 1;
      block_af_if_needed(c, P, c->blk_start[b], nrows);
      range_gemms(c, P, (uint32_t)nrows, c->blk_start[b], c->d_H, true, buf);
      PCA_CUDA(cudaEventRecord(c->ev_done[buf], c->stream));
    }
    c->af_done = true;
  }
}

// ArnoldiOpData::perform_op (Arnoldi.cpp:18-46): y = sum over blocks G_b (G_b^T x), the operator the
// IRAM solver (Spectra) iterates. One decode + GEMM pass with x in column 0 of Omega; the current
// update / standardize flags apply exactly as in computeGandH.
void perform_op(pcaone_ctx* c, const double* x_in, double* y_out) {
  if (c->source < 0) throw std::runtime_error("no genotype source set");
  if (c->update && c->cfg.emu && !c->have_usv) throw std::runtime_error("perform_op(update) without U,S,V");
  c->lut.standardize = (c->standardize && c->cfg.scale == -9) ? 1 : 0;
  const uint64_t HN = c->N * c->lp;
  ensure_stage(c, c->N);
  PCA_CUDA(cudaMemcpyAsync(c->d_stage, x_in, c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  zero_async(c, c->d_Omg, HN);
  PCA_CUDA(cudaMemcpy2DAsync(c->d_Omg, (size_t)c->lp * sizeof(double), c->d_stage, sizeof(double), sizeof(double), c->N,
                             cudaMemcpyDeviceToDevice, c->stream));
  c->omega_img_valid = c->omega_colmax_valid = false;
  walk_ranges(c);
  allreduce_H(c, c->d_H);
  PCA_CUDA(cudaMemcpy2DAsync(c->d_stage, sizeof(double), c->d_H, (size_t)c->lp * sizeof(double), sizeof(double), c->N,
                             cudaMemcpyDeviceToDevice, c->stream));
  PCA_CUDA(cudaMemcpyAsync(y_out, c->d_stage, c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  c->tm.h2d_bytes += c->N * sizeof(double);
  c->tm.d2h_bytes += c->N * sizeof(double);
}

// V = X^T A per SNP (+ the squared norm of every decoded SNP column): the device part of
// run_selection (Selection.cpp:16-34, `V.row(j) = U^T G.col(j); y_norm2(j) = G.col(j).squaredNorm()`).
// A: N x ncols (column-major, host), out: M x ncols (column-major), sqnorm: M or NULL.
void xt_times(pcaone_ctx* c, const double* A, uint32_t ncols, double* out, double* sqnorm) {
  if (c->source < 0) throw std::runtime_error("no genotype source set");
  if (ncols == 0 || (int)ncols > c->l) throw std::runtime_error("xt_times: ncols must be in [1, k + oversamples]");
  if (c->cfg.world > 1) throw std::runtime_error("xt_times: single-GPU only");
  if (c->update && c->cfg.emu && !c->have_usv) throw std::runtime_error("xt_times(update) without U,S,V");
  c->lut.standardize = (c->standardize && c->cfg.scale == -9) ? 1 : 0;
  zero_async(c, c->d_Omg, c->N * c->lp);
  upload_colmajor(c, A, c->N, (int)ncols, c->d_Omg);
  c->omega_img_valid = c->omega_colmax_valid = false;
  c->half = 1;
  try {
    walk_ranges(c);
  } catch (...) {
    c->half = 3;
    throw;
  }
  c->half = 3;
  download_colmajor(c, c->d_G, c->M, (int)ncols, out);
  if (sqnorm) {
    ensure_stage(c, c->M);
    if (c->source == PCAONE_SRC_RESIDENT) {
      k_snp_sqnorm<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_packed, c->pitch, (uint32_t)c->N, c->M,
                                                                           c->d_F, c->lut, c->d_stage);
    } else if (c->source == PCAONE_SRC_DOSAGE) {
      k_dosage_sqnorm<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_dos, c->ldf, (uint32_t)c->N, c->M,
                                                                              c->d_F, c->lut, c->d_stage);
    } else {
      throw std::runtime_error("xt_times: squared norms need a resident genotype or dosage source");
    }
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaMemcpyAsync(sqnorm, c->d_stage, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
  }
}

// U = X B: the device part of run_projection option 1 (Projection.cpp:236-241, `U = G * V` with
// V already scaled by 1 / S on the host side of the call). B: M x ncols, out: N x ncols (column-major).
void x_times(pcaone_ctx* c, const double* B, uint32_t ncols, double* out) {
  if (c->source < 0) throw std::runtime_error("no genotype source set");
  if (ncols == 0 || (int)ncols > c->l) throw std::runtime_error("x_times: ncols must be in [1, k + oversamples]");
  if (c->cfg.world > 1) throw std::runtime_error("x_times: single-GPU only");
  if (c->update && c->cfg.emu && !c->have_usv) throw std::runtime_error("x_times(update) without U,S,V");
  c->lut.standardize = (c->standardize && c->cfg.scale == -9) ? 1 : 0;
  zero_async(c, c->d_G, c->M * c->lp);
  upload_colmajor(c, B, c->M, (int)ncols, c->d_G);
  c->half = 2;
  try {
    walk_ranges(c);
  } catch (...) {
    c->half = 3;
    throw;
  }
  c->half = 3;
  download_colmajor(c, c->d_H, c->N, (int)ncols, out);
}

// RsvdOpOnePass::computeGandH (RSVD.hpp:137-166 plain, :168-252 windows) followed by
// RsvdOnePass::computeUSV (RSVD.hpp:281-313) on the dense source: a fixed number of power
// iterations, no convergence test. Result: d_U (ncol side, = svd.matrixV()), d_V (nrow side,
// = G * svd.matrixU()), d_S.
void dense_onepass(pcaone_ctx* c, uint32_t p, uint32_t windows, int finder) {
  if (c->source != PCAONE_SRC_DENSE) throw std::runtime_error("dense_rsvd: call pcaone_upload_dense first");
  if (finder != 1)
    throw std::runtime_error("dense_rsvd: only the QR range finder (finder = 1, RSVD.hpp:147-149) is implemented");
  if (!c->have_omg0) throw std::runtime_error("call pcaone_set_omega before dense_rsvd");
  const uint64_t HN = c->N * c->lp;
  PCA_CUDA(cudaMemcpyAsync(c->d_Omg, c->d_Omg0, HN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  PCA_CUDA(cudaMemcpyAsync(c->d_Omg2, c->d_Omg0, HN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  c->omega_img_valid = c->omega_colmax_valid = false;
  auto full_pass = [&]() { range_gemms(c, nullptr, (uint32_t)c->M, 0, c->d_H, false, -1); };
  full_pass();
  if (windows == 0) {
    for (uint32_t pi = 0; pi < p; ++pi) {
      update_omega(c, c->d_H, false);  // Omg = householderQ(H) * I, no flipOmg (RSVD.hpp:146-149)
      full_pass();
    }
  } else {
    if (windows % 2 != 0) throw std::runtime_error("windows must be a power of 2, ie. windows=2^x.");
    if (std::pow(2.0, (double)p) < (double)windows) throw std::runtime_error("pow(2, p) >= windows has to be met");
    const uint64_t bs = (c->M + windows - 1) / windows;
    if (bs < windows || (uint64_t)(windows - 1) * bs >= c->M)
      throw std::runtime_error("window size is smaller than number of windows because given matrix is too small");
    if (!c->d_H1) throw std::runtime_error("dense_rsvd with windows needs a context created with svd = PCAONE_SVD_WINSVD");
    zero_async(c, c->d_H1, HN);
    zero_async(c, c->d_H2, HN);
    auto update = [&](bool zero_h1) {
      k_add2<<<grid_for(HN, 256, c->sms), 256, 0, c->stream>>>(c->d_H1, c->d_H2, c->d_H, HN);
      PCA_CHECK_LAUNCH();
      c->tm.kernel_launches++;
      update_omega(c, c->d_H, true);
      zero_async(c, zero_h1 ? c->d_H1 : c->d_H2, HN);
    };
    uint64_t band = 1;
    for (uint32_t pi = 0; pi <= p; ++pi) {
      if (std::pow(2.0, (double)pi) >= (double)windows) {
        zero_async(c, c->d_H1, HN);
        zero_async(c, c->d_H2, HN);
      }
      band = std::min<uint64_t>(band * 2, windows);
      const double half_prev = pi > 0 ? std::pow(2.0, (double)pi - 1.0) : 0.0;
      const bool early = pi > 0 && std::pow(2.0, (double)pi) < (double)windows;
      uint64_t i = 1, j = 1;
      for (uint64_t b = 0; b < windows; ++b, ++i, ++j) {
        const uint64_t start = b * bs, stop = std::min<uint64_t>((b + 1) * bs, c->M) - 1;
        const uint32_t n = (uint32_t)(stop - start + 1);
        if (early && (double)j <= half_prev) {
          range_gemms(c, nullptr, n, start, c->d_H1, true, -1);
          if ((double)j == half_prev) update(false);  // complementary power iteration (RSVD.hpp:207-214)
        } else if (i <= band / 2) {
          range_gemms(c, nullptr, n, start, c->d_H1, true, -1);
        } else {
          range_gemms(c, nullptr, n, start, c->d_H2, true, -1);
        }
        if (b + 1 >= band) {
          if (i == band) {
            update(true);
            i = 0;
          } else if (i == band / 2) {
            update(false);
          }
        }
      }
    }
  }
  small_stage(c);
  finalize_usv(c);
  PCA_CUDA(cudaMemcpyAsync(c->d_U, c->d_Ucur, HN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
}

// flip_UV(U, V, false), Utils.cpp:136-143. V rows may be sharded: every rank then writes its
// column maxima into its own slot of a zeroed buffer and the sum-allreduce hook acts as an
// all-gather (k_flip_slot_write / k_flip_slot_pick).
void flip_uv(pcaone_ctx* c) {
  uint64_t rpc = std::max<uint64_t>(256, (c->M + c->sms - 1) / c->sms);
  int nparts = (int)((c->M + rpc - 1) / rpc);
  double* pval = c->d_part;
  double* psgn = c->d_part + (size_t)nparts * c->k;
  if ((size_t)nparts * 2 * c->k > c->part_doubles) throw std::runtime_error("partial workspace too small");
  k_colabsmax_partial<<<nparts, 256, 0, c->stream>>>(c->d_V, c->lp, c->k, c->M, rpc, pval, psgn, c->d_pidx);
  PCA_CHECK_LAUNCH();
  k_colabsmax_final<<<1, 128, 0, c->stream>>>(pval, psgn, c->d_pidx, nparts, c->k, c->d_scal + 8, c->d_sign);
  PCA_CHECK_LAUNCH();
  if (c->cfg.world > 1) {
    if (!c->allreduce) throw std::runtime_error("world > 1 but no allreduce hook installed");
    double* slots = c->d_part;  // the partials above are consumed
    k_flip_slot_write<<<1, 256, 0, c->stream>>>(c->d_scal + 8, c->d_sign, c->k, c->cfg.rank, c->cfg.world, slots);
    PCA_CHECK_LAUNCH();
    if (c->allreduce(c->allreduce_user, slots, (uint64_t)c->cfg.world * 2 * c->k, c->stream))
      throw std::runtime_error("allreduce hook failed");
    k_flip_slot_pick<<<1, 128, 0, c->stream>>>(slots, c->k, c->cfg.world, c->d_sign);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches += 2;
  }
  k_scale_cols<<<grid_for(c->M * c->k, 256, c->sms), 256, 0, c->stream>>>(c->d_V, c->lp, c->k, c->M, c->d_sign);
  k_scale_cols<<<grid_for(c->N * c->k, 256, c->sms), 256, 0, c->stream>>>(c->d_U, c->lp, c->k, c->N, c->d_sign);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches += 4;
}

// Halko.cpp:290-319 (EMU branch)
void run_em(pcaone_ctx* c, int* iters_out) {
  const int maxp = (int)c->cfg.maxp;
  const double tol = c->cfg.tol;
  c->update = 0;
  c->standardize = 0;
  compute_usv(c, maxp, tol);
  flip_uv(c);
  int iters = 0;
  const uint64_t vbytes = c->M * c->lp * sizeof(double);
  if (!c->d_Vpre) dmalloc(&c->d_Vpre, c->M * c->lp);  // PCAngsd EM on a context created with emu = 0
  for (uint32_t i = 0; i < c->cfg.maxiter; ++i) {
    c->update = 1;
    c->standardize = 0;
    PCA_CUDA(cudaMemcpyAsync(c->d_Vpre, c->d_V, vbytes, cudaMemcpyDeviceToDevice, c->stream));
    compute_usv(c, maxp, tol);
    flip_uv(c);
    const double diff = 1.0 - device_mev(c, c->d_V, c->d_Vpre, c->M, true);
    iters = (int)i + 1;
    if (diff < c->cfg.tolem) break;
  }
  if (c->cfg.emu) {
    c->update = 1;
    c->standardize = 1;
    compute_usv(c, maxp, tol);
    flip_uv(c);
  }
  if (iters_out) *iters_out = iters;
  PCA_CUDA(cudaStreamSynchronize(c->stream));
}

void set_blocks(pcaone_ctx* c, const uint64_t* start, const uint64_t* stop, uint32_t nblocks, uint32_t band_factor) {
  c->blk_start.assign(start, start + nblocks);
  c->blk_stop.assign(stop, stop + nblocks);
  c->band_factor = band_factor ? band_factor : 1;
  uint64_t mb = 0;
  for (uint32_t i = 0; i < nblocks; ++i) {
    if (stop[i] >= c->M || start[i] > stop[i]) throw std::runtime_error("set_blocks: block out of range");
    mb = std::max(mb, stop[i] - start[i] + 1);
  }
  if (mb > c->max_block && c->d_blk[0]) throw std::runtime_error("set_blocks: cannot grow blocks after streaming began");
  c->max_block = std::max(c->max_block, mb);
}

// ---------------------------------------------------------------- LD r2 (LD.cpp:450-473)
// Banded tile Gram on the FP64 tensor cores (ld.cuh). The SNP axis is walked in chunks of lead
// SNPs (+ a halo of the widest window) sized to the free HBM, so M x N need not fit at once.
// With `keep_out` the r^2 values stay on the device and feed the greedy pruning kernel chunk by
// chunk (ld_prune_big, LD.cpp:240-268); r2_out may then be NULL.
void ld_r2(pcaone_ctx* c, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
           double* r2_out, const double* af = nullptr, double r2_tol = 0.0, unsigned char* keep_out = nullptr) {
  const uint64_t N = c->N;
  if (N < 2) throw std::runtime_error("ld_r2: needs at least two samples");
  if (!G) {
    if (c->source != PCAONE_SRC_RESIDENT || !c->af_done)
      throw std::runtime_error("ld_r2: G == NULL needs a resident packed shard with allele frequencies");
    if (nsnps != c->M) throw std::runtime_error("ld_r2: nsnps must equal the resident shard size");
  }
  if (nwin == 0) {
    if (keep_out) memset(keep_out, 1, nsnps);
    return;
  }
  std::vector<uint64_t> offs(nwin + 1, 0);
  uint64_t maxwe = 1;
  for (uint64_t w = 0; w < nwin; ++w) {
    if (ws[w] < 0 || we[w] < 1 || (uint64_t)ws[w] + (uint64_t)we[w] > nsnps || (w && ws[w] <= ws[w - 1]))
      throw std::runtime_error("ld_r2: windows must be ascending and inside [0, nsnps)");
    offs[w + 1] = offs[w] + (uint64_t)(we[w] - 1);
    maxwe = std::max<uint64_t>(maxwe, (uint64_t)we[w]);
  }
  const uint32_t Np = (uint32_t)round_up(N, 16);
  const double df = 1.0 / (double)(N - 1);
  // chunk plan: leads per chunk (multiple of the tile) from the free memory
  size_t free_b = 0, total_b = 0;
  PCA_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const double budget = std::min<double>(0.5 * (double)free_b, 48.0 * (1ull << 30));
  const double per_row = (double)Np * 8 + (G ? (double)N * 8 : 0.0) + 16.0;
  const double per_lead = per_row + (double)(maxwe - 1) * 8;
  const uint64_t halo = maxwe;  // rows beyond the last lead of a chunk
  double leads_d = (budget - (double)halo * per_row) / per_lead;
  if (const char* e = getenv("PCAONE_LD_CHUNK")) leads_d = atof(e);  // test hook: force small chunks
  if (leads_d < ld::kTile) throw std::runtime_error("ld_r2: not enough device memory for one tile row of this window width");
  const uint64_t leads = std::min<uint64_t>(round_up(nsnps, ld::kTile), (uint64_t)leads_d / ld::kTile * ld::kTile);
  const uint64_t max_rows = std::min<uint64_t>(nsnps, leads + halo);

  double *d_Gs = nullptr, *d_raw = nullptr, *d_isd = nullptr, *d_out = nullptr;
  int32_t *d_winof = nullptr, *d_we = nullptr, *d_ws = nullptr;
  unsigned char* d_keep = nullptr;
  double* d_af = nullptr;
  uint64_t* d_offs = nullptr;
  int2* d_tiles = nullptr;
  size_t out_cap = 0, tiles_cap = 0;
  auto cleanup = [&]() {
    cudaFree(d_Gs); cudaFree(d_raw); cudaFree(d_isd); cudaFree(d_out);
    cudaFree(d_winof); cudaFree(d_we); cudaFree(d_offs); cudaFree(d_tiles);
    cudaFree(d_ws); cudaFree(d_keep); cudaFree(d_af);
  };
  try {
    dmalloc(&d_Gs, max_rows * Np);
    if (G) dmalloc(&d_raw, max_rows * N);
    dmalloc(&d_isd, max_rows);
    dmalloc(&d_winof, max_rows);
    dmalloc(&d_we, nwin);
    dmalloc(&d_offs, nwin + 1);
    PCA_CUDA(cudaMemcpyAsync(d_we, we, nwin * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaMemcpyAsync(d_offs, offs.data(), (nwin + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    if (keep_out) {
      dmalloc(&d_ws, nwin);
      dmalloc(&d_keep, nsnps);
      PCA_CUDA(cudaMemcpyAsync(d_ws, ws, nwin * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      PCA_CUDA(cudaMemsetAsync(d_keep, 1, nsnps, c->stream));
      if (af) {
        dmalloc(&d_af, nsnps);
        PCA_CUDA(cudaMemcpyAsync(d_af, af, nsnps * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      }
    }
    static bool attr = false;
    if (!attr) {
      PCA_CUDA(cudaFuncSetAttribute(ld::k_ld_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ld::kSmemBytes));
      attr = true;
    }
    LutParams lut = c->lut;
    lut.standardize = 0;  // --ld runs centred, unscaled genotypes (Halko.cpp:283-288)
    std::vector<int32_t> winof;
    std::vector<int2> tiles;
    uint64_t w_lo = 0;
    for (uint64_t c0 = 0; c0 < nsnps && w_lo < nwin; c0 += leads) {
      const uint64_t c1 = std::min<uint64_t>(nsnps, c0 + leads);
      uint64_t w_hi = w_lo;
      while (w_hi < nwin && (uint64_t)ws[w_hi] < c1) ++w_hi;
      if (w_hi == w_lo) continue;
      const uint64_t r1 = std::min<uint64_t>(nsnps, c1 + halo), rows = r1 - c0;
      // ---- operand chunk: padded SNP-major doubles
      if (G) {
        PCA_CUDA(cudaMemcpyAsync(d_raw, G + c0 * N, rows * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->tm.h2d_bytes += rows * N * sizeof(double);
        ld::k_pad_rows<<<grid_for(rows * Np, 256, c->sms), 256, 0, c->stream>>>(d_raw, rows, (uint32_t)N, Np, d_Gs);
      } else {
        ld::k_decode_rows<<<grid_for(rows * (Np >> 2), 256, c->sms), 256, 0, c->stream>>>(
            c->d_packed + c0 * c->pitch, c->pitch, (uint32_t)N, Np, rows, c->d_F + c0, lut, d_Gs);
      }
      PCA_CHECK_LAUNCH();
      ld::k_inv_sd<<<grid_for(rows * 32, 256, c->sms), 256, 0, c->stream>>>(d_Gs, rows, Np, df, d_isd);
      PCA_CHECK_LAUNCH();
      // ---- windows and tile list of the chunk
      winof.assign(rows, -1);
      const uint64_t nlt = (c1 - c0 + ld::kTile - 1) / ld::kTile;
      std::vector<int64_t> maxk(nlt, -1);
      for (uint64_t w = w_lo; w < w_hi; ++w) {
        const uint64_t i = (uint64_t)ws[w] - c0;
        winof[i] = (int32_t)w;
        if (we[w] > 1) maxk[i / ld::kTile] = std::max<int64_t>(maxk[i / ld::kTile], (int64_t)(i + we[w] - 1));
      }
      tiles.clear();
      for (uint64_t lt = 0; lt < nlt; ++lt)
        for (int64_t tk = (int64_t)lt; maxk[lt] >= 0 && tk <= maxk[lt] / ld::kTile; ++tk)
          tiles.push_back(make_int2((int)lt, (int)tk));
      const uint64_t nout = offs[w_hi] - offs[w_lo];
      if (!tiles.empty() && nout > 0) {
        if (tiles.size() > tiles_cap) {
          cudaFree(d_tiles);
          d_tiles = nullptr;
          dmalloc(&d_tiles, tiles.size());
          tiles_cap = tiles.size();
        }
        if (nout > out_cap) {
          cudaFree(d_out);
          d_out = nullptr;
          dmalloc(&d_out, nout);
          out_cap = nout;
        }
        PCA_CUDA(cudaMemcpyAsync(d_winof, winof.data(), rows * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        PCA_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        ld::LdArgs a{};
        a.Gs = d_Gs;
        a.Np = Np;
        a.rows = rows;
        a.inv_sd = d_isd;
        a.df = df;
        a.tiles = d_tiles;
        a.win_of = d_winof;
        a.we = d_we;
        a.offs = d_offs;
        a.out0 = offs[w_lo];
        a.out = d_out;
        {
          Timed t(c, 9);
          ld::k_ld_tiles<<<(unsigned)tiles.size(), ld::kThreads, ld::kSmemBytes, c->stream>>>(a);
          PCA_CHECK_LAUNCH();
        }
        c->tm.ld_tiles += tiles.size();
        c->tm.ld_pairs += nout;
        c->tm.kernel_launches += 3;
        if (keep_out) {
          ld::k_ld_prune<<<1, 1024, 0, c->stream>>>(d_out, offs[w_lo], d_offs, d_ws, d_we, w_lo, w_hi, d_af, r2_tol,
                                                     d_keep);
          PCA_CHECK_LAUNCH();
          c->tm.kernel_launches++;
        }
        if (r2_out) {
          PCA_CUDA(cudaMemcpyAsync(r2_out + offs[w_lo], d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
          c->tm.d2h_bytes += nout * sizeof(double);
        }
      }
      PCA_CUDA(cudaStreamSynchronize(c->stream));  // winof / tiles host vectors are reused
      w_lo = w_hi;
    }
    if (keep_out) {
      PCA_CUDA(cudaMemcpyAsync(keep_out, d_keep, nsnps, cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      c->tm.d2h_bytes += nsnps;
    }
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
}

}  // namespace

// =============================================================================== C-ABI
extern "C" {

int pcaone_abi_version(void) { return 1; }

const char* pcaone_last_error(const pcaone_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pcaone_create(const pcaone_config* cfg, pcaone_ctx** out) {
  if (!cfg || !out) return 1;
  pcaone_ctx* c = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw std::runtime_error(std::string("pcaone_b200 needs a CUDA device (no CPU fallback): ") +
                               cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) throw std::runtime_error("invalid CUDA device ordinal");
    PCA_CUDA(cudaSetDevice(cfg->device));
    c = new pcaone_ctx();
    c->cfg = *cfg;
    if (c->cfg.world < 1) c->cfg.world = 1;
    if (c->cfg.nsnps_total == 0) c->cfg.nsnps_total = c->cfg.nsnps;
    if (c->cfg.bands == 0) c->cfg.bands = 64;
    c->N = cfg->nsamples;
    c->M = cfg->nsnps;
    c->k = (int)cfg->k;
    c->l = (int)(cfg->k + cfg->oversamples);
    if (c->N == 0 || c->M == 0 || c->k == 0) throw std::runtime_error("nsamples, nsnps and k must be positive");
    if (c->l > kMaxL) throw std::runtime_error("k + oversamples must be <= 112");
    if ((uint64_t)c->l > c->N || (uint64_t)c->l > c->M) throw std::runtime_error("k + oversamples exceeds the matrix size");
    if (cfg->precision != PCAONE_PREC_FP64 && cfg->precision != PCAONE_PREC_INT8X2 &&
        cfg->precision != PCAONE_PREC_INT8X3 && cfg->precision != PCAONE_PREC_INT8X4)
      throw std::runtime_error("precision must be PCAONE_PREC_FP64 or PCAONE_PREC_INT8X2/3/4");
    if (cfg->svd != PCAONE_SVD_SSVD && cfg->svd != PCAONE_SVD_WINSVD) throw std::runtime_error("svd must be 1 or 2");
    c->NT = supported_nt(c->l);
    c->lp = c->NT * 8;
    if (cfg->precision != PCAONE_PREC_FP64) {
      c->slices = cfg->precision;
      c->NP = (int)round_up((size_t)c->slices * c->l, 16);
      if (c->NP > tc::kMaxNP)
        throw std::runtime_error("slices * (k + oversamples) must be <= 256 for the tensor-core path");
      c->RT = c->NP <= 128 ? 2 : 1;
    }
    c->bpr = (uint32_t)((c->N + 3) >> 2);
    c->pitch = (uint32_t)round_up(c->bpr, 16);
    c->lut.sqrt_ploidy = sqrt((double)cfg->ploidy);
    c->lut.standardize = 0;
    cudaDeviceProp prop;
    PCA_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    c->sms = prop.multiProcessorCount;
    PCA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PCA_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    const size_t NL = c->N * c->lp, ML = c->M * c->lp, LL = (size_t)c->lp * c->lp;
    dmalloc(&c->d_Omg0, NL);
    dmalloc(&c->d_Omg, NL);
    dmalloc(&c->d_Omg2, NL);
    dmalloc(&c->d_H, NL);
    dmalloc(&c->d_Bt, NL);
    dmalloc(&c->d_Ucur, NL);
    dmalloc(&c->d_Upre, NL);
    dmalloc(&c->d_U, NL);
    if (cfg->svd == PCAONE_SVD_WINSVD) {
      dmalloc(&c->d_H1, NL);
      dmalloc(&c->d_H2, NL);
      PCA_CUDA(cudaMemset(c->d_H1, 0, NL * sizeof(double)));
      PCA_CUDA(cudaMemset(c->d_H2, 0, NL * sizeof(double)));
    }
    dmalloc(&c->d_G, ML);
    dmalloc(&c->d_V, ML);
    if (cfg->emu) dmalloc(&c->d_Vpre, ML);
    PCA_CUDA(cudaMemset(c->d_V, 0, ML * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_U, 0, NL * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_H, 0, NL * sizeof(double)));
    dmalloc(&c->d_S, c->lp);
    dmalloc(&c->d_F, c->M);
    dmalloc(&c->d_nmiss, c->M);
    PCA_CUDA(cudaMemset(c->d_F, 0, c->M * sizeof(double)));
    PCA_CUDA(cudaMemset(c->d_nmiss, 0xff, c->M * sizeof(uint32_t)));  // unknown -> treated as "has missing"
    const uint32_t tiles = ceil_div(c->N, kTileRows);
    c->max_splits = std::max<uint32_t>(1, std::min<uint32_t>(64, (2u * c->sms + tiles - 1) / tiles));
    dmalloc(&c->d_Hpart, (size_t)c->max_splits * NL);
    for (double** p : {&c->d_W, &c->d_R, &c->d_Rinv, &c->d_T1, &c->d_T2, &c->d_T, &c->d_Vr, &c->d_Z}) dmalloc(p, LL);
    dmalloc(&c->d_sigma, c->lp);
    dmalloc(&c->d_sign, c->lp);
    dmalloc(&c->d_hsign, c->lp);
    dmalloc(&c->d_scal, 64);
    dmalloc(&c->d_status, 4);
    PCA_CUDA(cudaMemset(c->d_status, 0, 4 * sizeof(int)));
    dmalloc(&c->d_jscratch, (size_t)2 * c->l * c->l + 2 * c->l + 8);
    if (const char* e = getenv("PCAONE_FUSED_ORTH")) c->fused_orth = atoi(e);
    PCA_CUDA(cudaHostAlloc((void**)&c->h_status, 4 * sizeof(int), cudaHostAllocDefault));
    PCA_CUDA(cudaHostAlloc((void**)&c->h_scal, 64 * sizeof(double), cudaHostAllocDefault));
    c->part_doubles = (size_t)(2 * c->sms + 8) * 128 * c->lp;
    dmalloc(&c->d_part, c->part_doubles);
    dmalloc(&c->d_pidx, (size_t)(2 * c->sms + 8) * 128);
    PCA_CUDA(cudaDeviceSynchronize());
    *out = c;
    return 0;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    delete c;
    return 1;
  }
}

void pcaone_destroy(pcaone_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : {(void*)c->d_Omg0, (void*)c->d_packed, (void*)c->d_F, (void*)c->d_nmiss, (void*)c->d_Omg, (void*)c->d_Omg2,
                  (void*)c->d_H, (void*)c->d_H1, (void*)c->d_H2, (void*)c->d_Bt, (void*)c->d_Ucur, (void*)c->d_Upre,
                  (void*)c->d_U, (void*)c->d_G, (void*)c->d_V, (void*)c->d_Vpre, (void*)c->d_S, (void*)c->d_Hpart,
                  (void*)c->d_W, (void*)c->d_R, (void*)c->d_Rinv, (void*)c->d_T1, (void*)c->d_T2, (void*)c->d_T,
                  (void*)c->d_Vr, (void*)c->d_Z, (void*)c->d_sigma, (void*)c->d_sign, (void*)c->d_hsign, (void*)c->d_scal,
                  (void*)c->d_status, (void*)c->d_part, (void*)c->d_pidx, (void*)c->d_stage, (void*)c->d_raw[0],
                  (void*)c->d_raw[1], (void*)c->d_blk[0], (void*)c->d_blk[1], (void*)c->d_PG, (void*)c->d_PH, (void*)c->d_PGb[0],
                  (void*)c->d_PGb[1], (void*)c->d_PHb[0], (void*)c->d_PHb[1], (void*)c->d_BimgO, (void*)c->d_BimgW,
                  (void*)c->d_dense, (void*)c->d_P, (void*)c->d_dos, (void*)c->d_Racc, (void*)c->d_Racc2, (void*)c->d_BimgD, (void*)c->d_tcs, (void*)c->d_Fpart, (void*)c->d_jscratch})
    if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (c->h_pin[i]) cudaFreeHost(c->h_pin[i]);
    if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  }
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  for (auto& e : c->evs) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (c->bed_file) fclose(c->bed_file);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

void* pcaone_stream(pcaone_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pcaone_alloc_pinned(void** out, size_t bytes) {
  if (!out) return 1;
  *out = nullptr;
  return cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable) == cudaSuccess ? 0 : 1;
}
void pcaone_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
int pcaone_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int pcaone_sync(pcaone_ctx* c) { CTX_GUARD(c, PCA_CUDA(cudaStreamSynchronize(c->stream))); }
int pcaone_set_allreduce(pcaone_ctx* c, pcaone_allreduce_fn fn, void* user) {
  CTX_GUARD(c, {
    c->allreduce = fn;
    c->allreduce_user = user;
  });
}

int pcaone_upload_bed(pcaone_ctx* c, const uint8_t* packed, uint64_t nsnps, int device_ptr) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("upload_bed: nsnps does not match the context");
    if (!c->d_packed) dmalloc(&c->d_packed, c->M * (size_t)c->pitch);
    const cudaMemcpyKind kind = device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (c->pitch == c->bpr) {
      PCA_CUDA(cudaMemcpyAsync(c->d_packed, packed, c->M * (size_t)c->bpr, kind, c->stream));
    } else if (device_ptr) {
      k_repitch<<<grid_for(c->M * (c->pitch >> 4), 256, c->sms), 256, 0, c->stream>>>(packed, c->d_packed, c->M, c->bpr,
                                                                                     c->pitch);
      PCA_CHECK_LAUNCH();
    } else {
      // chunked: stage raw rows then repitch on the device
      const uint64_t chunk = std::max<uint64_t>(1, (256ull << 20) / c->bpr);
      uint8_t* raw = nullptr;
      dmalloc(&raw, std::min(chunk, c->M) * (size_t)c->bpr);
      for (uint64_t s = 0; s < c->M; s += chunk) {
        const uint64_t n = std::min(chunk, c->M - s);
        PCA_CUDA(cudaMemcpyAsync(raw, packed + s * c->bpr, n * (size_t)c->bpr, kind, c->stream));
        k_repitch<<<grid_for(n * (c->pitch >> 4), 256, c->sms), 256, 0, c->stream>>>(
            raw, c->d_packed + s * c->pitch, n, c->bpr, c->pitch);
        PCA_CHECK_LAUNCH();
      }
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(raw);
    }
    if (!device_ptr) c->tm.h2d_bytes += c->M * (size_t)c->bpr;
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->source = PCAONE_SRC_RESIDENT;
    c->af_done = false;
    c->tiles_valid = false;
    c->h_nmiss.clear();
    c->nmiss_prefix.clear();
  });
}

int pcaone_set_host_source(pcaone_ctx* c, const uint8_t* packed, uint64_t nsnps) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("set_host_source: nsnps does not match the context");
    c->h_packed = packed;
    c->source = PCAONE_SRC_HOST;
    c->af_done = false;
  });
}

int pcaone_set_reader_source(pcaone_ctx* c, pcaone_read_block_fn fn, void* user) {
  CTX_GUARD(c, {
    c->reader = fn;
    c->reader_user = user;
    c->source = PCAONE_SRC_FILE;
    c->af_done = false;
  });
}

int pcaone_open_bed(pcaone_ctx* c, const char* path, uint64_t snp_offset) {
  CTX_GUARD(c, {
    if (c->bed_file) fclose(c->bed_file);
    c->bed_file = fopen(path, "rb");
    if (!c->bed_file) throw std::runtime_error("Cannot open bed file.");
    unsigned char hdr[3];
    if (fread(hdr, 1, 3, c->bed_file) != 3 || hdr[0] != 0x6c || hdr[1] != 0x1b || hdr[2] != 0x01)
      throw std::runtime_error("Incorrect magic number in plink bed file.");
    c->bed_snp_offset = snp_offset;
    c->reader = nullptr;
    c->source = PCAONE_SRC_FILE;
    c->af_done = false;
  });
}

int pcaone_set_blocks(pcaone_ctx* c, const uint64_t* start, const uint64_t* stop, uint32_t nblocks,
                      uint32_t band_factor) {
  CTX_GUARD(c, set_blocks(c, start, stop, nblocks, band_factor));
}

int pcaone_permute_resident(pcaone_ctx* c, const uint32_t* indices) {
  CTX_GUARD(c, {
    if (c->source != PCAONE_SRC_RESIDENT && c->source != PCAONE_SRC_DOSAGE && c->source != PCAONE_SRC_GL)
      throw std::runtime_error("permute_resident needs a resident shard");
    const bool gl = c->source == PCAONE_SRC_GL;
    const bool dos = c->source == PCAONE_SRC_DOSAGE || gl;  // rows of d_dos / d_P instead of d_packed
    const uint32_t row_bytes = gl ? (uint32_t)(16 * c->N) : dos ? c->ldf * (uint32_t)sizeof(float) : c->pitch;  // multiples of 16
    uint32_t* d_idx = nullptr;
    uint8_t* d_new = nullptr;
    double* d_Fn = nullptr;
    dmalloc(&d_idx, c->M);
    dmalloc(&d_new, c->M * (size_t)row_bytes);
    dmalloc(&d_Fn, c->M);
    PCA_CUDA(cudaMemcpyAsync(d_idx, indices, c->M * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    k_gather_rows<<<grid_for(c->M * (row_bytes >> 4), 256, c->sms), 256, 0, c->stream>>>(
        gl ? reinterpret_cast<const uint8_t*>(c->d_P) : dos ? reinterpret_cast<const uint8_t*>(c->d_dos) : c->d_packed,
        d_new, d_idx, c->M, row_bytes);
    k_gather_f64<<<grid_for(c->M, 256, c->sms), 256, 0, c->stream>>>(c->d_F, d_Fn, d_idx, c->M);
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    uint32_t* d_nm = nullptr;
    dmalloc(&d_nm, c->M);
    k_gather_u32<<<grid_for(c->M, 256, c->sms), 256, 0, c->stream>>>(c->d_nmiss, d_nm, d_idx, c->M);
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(gl ? (void*)c->d_P : dos ? (void*)c->d_dos : (void*)c->d_packed);
    cudaFree(c->d_F);
    cudaFree(c->d_nmiss);
    cudaFree(d_idx);
    if (gl)
      c->d_P = reinterpret_cast<double*>(d_new);
    else if (dos)
      c->d_dos = reinterpret_cast<float*>(d_new);
    else
      c->d_packed = d_new;
    c->d_F = d_Fn;
    c->d_nmiss = d_nm;
    c->tiles_valid = false;
    c->h_nmiss.clear();
    c->nmiss_prefix.clear();
  });
}

int pcaone_allele_freq(pcaone_ctx* c) {
  CTX_GUARD(c, {
    if (c->source == PCAONE_SRC_RESIDENT) {
      Timed t(c, 6);
      k_allele_freq<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_packed, c->pitch, (uint32_t)c->N,
                                                                             c->M, c->d_F, c->d_nmiss);
      PCA_CHECK_LAUNCH();
      c->tm.kernel_launches++;
    } else if (c->source == PCAONE_SRC_DOSAGE) {
      Timed t(c, 6);
      k_dosage_af<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_dos, c->ldf, (uint32_t)c->N, c->M,
                                                                          c->d_F, c->d_nmiss);
      PCA_CHECK_LAUNCH();
      c->tm.kernel_launches++;
    } else if (c->source == PCAONE_SRC_GL) {
      throw std::runtime_error("allele_freq: genotype likelihoods use pcaone_gl_em_maf");
    } else if (c->source == PCAONE_SRC_DENSE) {
      throw std::runtime_error("allele_freq: a dense matrix has no allele frequencies");
    } else if (c->source >= 0) {
      if (c->blk_start.empty()) throw std::runtime_error("allele_freq on a streamed source needs pcaone_set_blocks");
      alloc_stream_buffers(c);
      for (uint32_t b = 0; b < c->blk_start.size(); ++b) {
        const int buf = b & 1;
        const uint8_t* P = stage_block(c, b, buf);
        block_af_if_needed(c, P, c->blk_start[b], c->blk_stop[b] - c->blk_start[b] + 1);
        PCA_CUDA(cudaEventRecord(c->ev_done[buf], c->stream));
      }
    } else {
      throw std::runtime_error("no genotype source set");
    }
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->af_done = true;
  });
}

int pcaone_get_F(pcaone_ctx* c, double* F) {
  CTX_GUARD(c, {
    PCA_CUDA(cudaMemcpyAsync(F, c->d_F, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
  });
}
int pcaone_set_F(pcaone_ctx* c, const double* F) {
  CTX_GUARD(c, {
    PCA_CUDA(cudaMemcpyAsync(c->d_F, F, c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->af_done = true;
  });
}
int pcaone_get_lookup(pcaone_ctx* c, double* lut) {
  CTX_GUARD(c, {
    ensure_stage(c, 4 * c->M);
    k_lookup_scale<<<grid_for(c->M, 256, c->sms), 256, 0, c->stream>>>(c->d_F, c->M, c->lut, c->d_stage, nullptr);
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaMemcpyAsync(lut, c->d_stage, 4 * c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
  });
}
int pcaone_get_scale(pcaone_ctx* c, double* s) {
  CTX_GUARD(c, {
    ensure_stage(c, c->M);
    LutParams p = c->lut;
    p.standardize = c->cfg.scale == -9 ? 1 : 0;
    k_lookup_scale<<<grid_for(c->M, 256, c->sms), 256, 0, c->stream>>>(c->d_F, c->M, p, nullptr, c->d_stage);
    PCA_CHECK_LAUNCH();
    PCA_CUDA(cudaMemcpyAsync(s, c->d_stage, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
  });
}
int pcaone_missing_count(pcaone_ctx* c, uint64_t* n) {
  CTX_GUARD(c, {
    std::vector<uint32_t> h(c->M);
    PCA_CUDA(cudaMemcpyAsync(h.data(), c->d_nmiss, c->M * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    uint64_t s = 0;
    for (auto v : h) s += v;
    *n = s;
  });
}

int pcaone_decode_block(pcaone_ctx* c, uint64_t start, uint64_t stop, int standardize, int update, double* out) {
  CTX_GUARD(c, {
    if (stop < start || stop >= c->M) throw std::runtime_error("decode_block: range out of bounds");
    const uint64_t B = stop - start + 1;
    if (c->source == PCAONE_SRC_DOSAGE) {  // FileBgen::read_block_initial, FileBgen.cpp:96-110
      if (!c->af_done) throw std::runtime_error("decode_block: call pcaone_allele_freq first");
      if (update && c->cfg.emu) throw std::runtime_error("--emu on a dosage source is not implemented");
      LutParams p = c->lut;
      p.standardize = (standardize && c->cfg.scale == -9) ? 1 : 0;
      ensure_stage(c, c->N * B);
      k_dosage_decode<<<grid_for(c->N * B, 256, c->sms), 256, 0, c->stream>>>(c->d_dos + start * c->ldf, c->ldf,
                                                                             (uint32_t)c->N, B, c->d_F + start, p,
                                                                             c->d_stage);
      PCA_CHECK_LAUNCH();
      PCA_CUDA(cudaMemcpyAsync(out, c->d_stage, c->N * B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      return 0;
    }
    if (c->source == PCAONE_SRC_GL) {  // E block: initial (FileBeagle.cpp:57-66) or fit_with_pi (Data.cpp:296-316)
      ensure_stage(c, c->N * B);
      gl_refresh(c, start, B, update != 0, c->d_stage, (uint32_t)c->N);
      PCA_CUDA(cudaMemcpyAsync(out, c->d_stage, c->N * B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      return 0;
    }
    if (c->source == PCAONE_SRC_DENSE) throw std::runtime_error("decode_block: not a genotype source");
    const uint8_t* P;
    if (c->source == PCAONE_SRC_RESIDENT) {
      P = c->d_packed + start * c->pitch;
    } else {
      // stage the requested range through buffer 0 as a one-off block
      std::vector<uint64_t> s0 = c->blk_start, e0 = c->blk_stop;
      const uint64_t mb = c->max_block;
      if (B > c->max_block && c->d_blk[0]) throw std::runtime_error("decode_block: range larger than the block plan");
      c->max_block = std::max(c->max_block, B);
      alloc_stream_buffers(c);
      c->blk_start = {start};
      c->blk_stop = {stop};
      P = stage_block(c, 0, 0);
      c->blk_start = s0;
      c->blk_stop = e0;
      c->max_block = std::max(mb, c->max_block);
    }
    if (!c->af_done) block_af_if_needed(c, P, start, B);
    LutParams p = c->lut;
    p.standardize = (standardize && c->cfg.scale == -9) ? 1 : 0;
    ensure_stage(c, c->N * B);
    const int emu = (update && c->cfg.emu) ? 1 : 0;
    if (emu && !c->have_usv) throw std::runtime_error("decode_block(update) without U,S,V");
    Timed t(c, 6);
    k_decode_block<<<grid_for(((c->N + 3) / 4) * B, 256, c->sms), 256, 0, c->stream>>>(
        P, c->pitch, (uint32_t)c->N, (uint32_t)B, c->d_F + start, p, emu, c->d_U, c->lp, c->d_S,
        c->d_V + start * c->lp, c->lp, c->k, c->d_stage);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
    if (c->source != PCAONE_SRC_RESIDENT) PCA_CUDA(cudaEventRecord(c->ev_done[0], c->stream));
    PCA_CUDA(cudaMemcpyAsync(out, c->d_stage, c->N * B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.d2h_bytes += c->N * B * sizeof(double);
  });
}

int pcaone_set_flags(pcaone_ctx* c, int update, int standardize) {
  CTX_GUARD(c, {
    c->update = update;
    c->standardize = standardize;
  });
}
int pcaone_set_omega(pcaone_ctx* c, const double* Omg) {
  CTX_GUARD(c, {
    upload_colmajor(c, Omg, c->N, c->l, c->d_Omg0);
    const size_t nb = c->N * c->lp * sizeof(double);
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg, c->d_Omg0, nb, cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->d_Omg2, c->d_Omg0, nb, cudaMemcpyDeviceToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->have_omg0 = true;
  });
}
int pcaone_get_omega(pcaone_ctx* c, double* Omg) { CTX_GUARD(c, download_colmajor(c, c->d_Omg, c->N, c->l, Omg)); }
int pcaone_set_usv(pcaone_ctx* c, const double* U, const double* S, const double* V) {
  CTX_GUARD(c, {
    upload_colmajor(c, U, c->N, c->k, c->d_U);
    upload_colmajor(c, V, c->M, c->k, c->d_V);
    PCA_CUDA(cudaMemcpyAsync(c->d_S, S, c->k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->have_usv = true;
  });
}
int pcaone_get_usv(pcaone_ctx* c, double* U, double* S, double* V) {
  CTX_GUARD(c, {
    if (U) download_colmajor(c, c->d_U, c->N, c->k, U);
    if (V) download_colmajor(c, c->d_V, c->M, c->k, V);
    if (S) {
      PCA_CUDA(cudaMemcpyAsync(S, c->d_S, c->k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
    }
  });
}
int pcaone_get_GH(pcaone_ctx* c, double* G, double* H) {
  CTX_GUARD(c, {
    if (G) download_colmajor(c, c->d_G, c->M, c->l, G);
    if (H) download_colmajor(c, c->d_H, c->N, c->l, H);
  });
}
int pcaone_set_H(pcaone_ctx* c, const double* H) { CTX_GUARD(c, upload_colmajor(c, H, c->N, c->l, c->d_H)); }

int pcaone_compute_gandh(pcaone_ctx* c, int pi) { CTX_GUARD(c, compute_gandh(c, pi)); }
int pcaone_small_stage(pcaone_ctx* c) { CTX_GUARD(c, small_stage(c)); }
int pcaone_compute_usv(pcaone_ctx* c, int maxp, double tol, double* diff_out, int* epochs_out) {
  CTX_GUARD(c, {
    compute_usv(c, maxp, tol);
    if (diff_out) *diff_out = c->last_diff;
    if (epochs_out) *epochs_out = c->last_epochs;
  });
}
int pcaone_run_em(pcaone_ctx* c, int* iters_out) { CTX_GUARD(c, run_em(c, iters_out)); }
int pcaone_orth_omega(pcaone_ctx* c, int flip) { CTX_GUARD(c, update_omega(c, c->d_H, flip != 0)); }

int pcaone_mev(pcaone_ctx* c, const double* X, const double* Y, uint64_t rows, uint32_t cols, double* out) {
  CTX_GUARD(c, {
    if ((int)cols > c->lp) throw std::runtime_error("mev: too many columns");
    double *dx = nullptr, *dy = nullptr;
    dmalloc(&dx, rows * c->lp);
    dmalloc(&dy, rows * c->lp);
    ensure_stage(c, rows * cols);
    dim3 blk(32, 8);
    for (int i = 0; i < 2; ++i) {
      PCA_CUDA(cudaMemcpyAsync(c->d_stage, i ? Y : X, rows * cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      k_colmajor_to_rowmajor<<<ceil_div(rows, 32), blk, 0, c->stream>>>(c->d_stage, rows, (int)cols, i ? dy : dx, c->lp);
      PCA_CHECK_LAUNCH();
    }
    const int ksave = c->k;
    c->k = (int)cols;
    double r = 0.0;
    try {
      r = device_mev(c, dx, dy, rows, false);
    } catch (...) {
      c->k = ksave;
      throw;
    }
    c->k = ksave;
    cudaFree(dx);
    cudaFree(dy);
    *out = r;
  });
}

int pcaone_upload_dense(pcaone_ctx* c, const double* A, uint64_t rows, uint64_t cols) {
  CTX_GUARD(c, {
    const bool trans = rows < cols;  // RSVD.hpp:113-121: a wide matrix is used transposed
    const uint64_t nrow = trans ? cols : rows, ncol = trans ? rows : cols;
    if (nrow != c->M || ncol != c->N)
      throw std::runtime_error("upload_dense: context must be created with nsnps = max(rows, cols), nsamples = min(rows, cols)");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_dense: the dense source runs in FP64");
    if (c->cfg.world > 1) throw std::runtime_error("upload_dense: single-GPU only");
    c->ldd = (uint32_t)round_up(c->N, 8);
    if (!c->d_dense) dmalloc(&c->d_dense, c->M * (size_t)c->ldd);
    if (trans) {
      // A^T in row-major is A in column-major: rows of length N, re-pitched to ldd
      PCA_CUDA(cudaMemsetAsync(c->d_dense, 0, c->M * (size_t)c->ldd * sizeof(double), c->stream));
      PCA_CUDA(cudaMemcpy2DAsync(c->d_dense, (size_t)c->ldd * sizeof(double), A, c->N * sizeof(double),
                                 c->N * sizeof(double), c->M, cudaMemcpyHostToDevice, c->stream));
    } else {
      double* stage = nullptr;
      dmalloc(&stage, c->M * c->N);
      PCA_CUDA(cudaMemcpyAsync(stage, A, c->M * c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      dim3 grid((unsigned)ceil_div(c->M, 32), (unsigned)ceil_div(c->ldd, 32));
      k_dense_transpose_in<<<grid, 256, 0, c->stream>>>(stage, c->M, c->N, c->d_dense, c->ldd);
      PCA_CHECK_LAUNCH();
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(stage);
    }
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.h2d_bytes += c->M * c->N * sizeof(double);
    c->source = PCAONE_SRC_DENSE;
  });
}

int pcaone_ld_prune(pcaone_ctx* c, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
                    const double* af, double r2_tol, uint8_t* keep_out) {
  CTX_GUARD(c, {
    if (!keep_out) throw std::runtime_error("ld_prune: keep_out is NULL");
    ld_r2(c, G, nsnps, ws, we, nwin, nullptr, af, r2_tol, keep_out);
  });
}

int pcaone_xt_times(pcaone_ctx* c, const double* A, uint32_t ncols, double* out, double* sqnorm) {
  CTX_GUARD(c, xt_times(c, A, ncols, out, sqnorm));
}
int pcaone_x_times(pcaone_ctx* c, const double* B, uint32_t ncols, double* out) { CTX_GUARD(c, x_times(c, B, ncols, out)); }

int pcaone_perform_op(pcaone_ctx* c, const double* x_in, double* y_out) { CTX_GUARD(c, perform_op(c, x_in, y_out)); }

int pcaone_upload_dosage(pcaone_ctx* c, const float* dosage, uint64_t nsnps, int device_ptr) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("upload_dosage: nsnps does not match the context");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_dosage: the dosage source runs in FP64");
    if (c->cfg.emu) throw std::runtime_error("--emu on a dosage source is not implemented");
    c->ldf = (uint32_t)round_up(c->N, 8);
    if (!c->d_dos) dmalloc(&c->d_dos, c->M * (size_t)c->ldf);
    PCA_CUDA(cudaMemsetAsync(c->d_dos, 0, c->M * (size_t)c->ldf * sizeof(float), c->stream));
    PCA_CUDA(cudaMemcpy2DAsync(c->d_dos, (size_t)c->ldf * sizeof(float), dosage, c->N * sizeof(float),
                               c->N * sizeof(float), c->M, device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    if (!device_ptr) c->tm.h2d_bytes += c->M * c->N * sizeof(float);
    c->source = PCAONE_SRC_DOSAGE;
    c->af_done = false;
  });
}

int pcaone_upload_gl(pcaone_ctx* c, const double* P, uint64_t nsnps, int device_ptr) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("upload_gl: nsnps does not match the context");
    if (c->cfg.precision != PCAONE_PREC_FP64) throw std::runtime_error("upload_gl: genotype likelihoods run in FP64");
    if (c->cfg.emu) throw std::runtime_error("upload_gl: --emu does not apply to genotype likelihoods (PCAngsd EM is pcaone_run_em with emu = 0)");
    if (c->cfg.world > 1) throw std::runtime_error("upload_gl: single-GPU only");
    c->ldd = (uint32_t)round_up(c->N, 8);
    if (!c->d_P) dmalloc(&c->d_P, c->M * 2 * c->N);
    if (!c->d_dense) dmalloc(&c->d_dense, c->M * (size_t)c->ldd);
    PCA_CUDA(cudaMemsetAsync(c->d_dense, 0, c->M * (size_t)c->ldd * sizeof(double), c->stream));
    PCA_CUDA(cudaMemcpyAsync(c->d_P, P, c->M * 2 * c->N * sizeof(double),
                             device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    PCA_CUDA(cudaStreamSynchronize(c->stream));
    if (!device_ptr) c->tm.h2d_bytes += c->M * 2 * c->N * sizeof(double);
    c->source = PCAONE_SRC_GL;
    c->af_done = false;
  });
}

int pcaone_gl_em_maf(pcaone_ctx* c, uint32_t maxiter, double tolmaf, int* iters_out) {
  CTX_GUARD(c, {
    if (c->source != PCAONE_SRC_GL) throw std::runtime_error("gl_em_maf: call pcaone_upload_gl first");
    // emMAF_with_GL (Utils.cpp:745-775): F = 0.25, EM steps until the RMS change over all variants < tolmaf
    std::vector<double> f0(c->M, 0.25);
    PCA_CUDA(cudaMemcpyAsync(c->d_F, f0.data(), c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
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
    if (iters_out) *iters_out = it;
    PCA_CUDA(cudaMemsetAsync(c->d_nmiss, 0, c->M * sizeof(uint32_t), c->stream));
    c->af_done = true;
  });
}

int pcaone_dense_rsvd(pcaone_ctx* c, uint32_t p, uint32_t windows, int finder) {
  CTX_GUARD(c, dense_onepass(c, p, windows, finder));
}

int pcaone_ld_r2(pcaone_ctx* c, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
                 double* r2_out) {
  CTX_GUARD(c, ld_r2(c, G, nsnps, ws, we, nwin, r2_out));
}

int pcaone_get_timers(pcaone_ctx* c, pcaone_timers* out, int reset) {
  CTX_GUARD(c, {
    resolve_timers(c);
    c->tm.tc_ranges = c->tc_ranges;
    c->tm.fp64_ranges = c->fp64_ranges;
    c->tm.tc_miss_ranges = c->tc_miss_ranges;
    if (out) *out = c->tm;
    if (reset) {
      c->tm = pcaone_timers{};
      c->tc_ranges = c->fp64_ranges = c->tc_miss_ranges = 0;
    }
  });
}
int pcaone_enable_timing(pcaone_ctx* c, int on) { CTX_GUARD(c, c->timing = on != 0); }

// ---- host helpers that must match the reference's libstdc++ streams bit for bit -------------
// RsvdOpData::initOmg (Halko.cpp:15-23) with StandardNormalRandom / UniformRandom
// (RSVD.hpp:20-59): std::default_random_engine seeded with `seed`, values drawn in
// column-major order (Eigen NullaryExpr evaluation order for a column-major MatrixXd).
int pcaone_init_omega(uint64_t rows, uint32_t cols, int seed, int gaussian, double* out) {
  auto rng = std::default_random_engine{};
  rng.seed(seed);
  const uint64_t n = rows * cols;
  if (gaussian) {
    std::normal_distribution<double> dist{0, 1};
    for (uint64_t i = 0; i < n; ++i) out[i] = dist(rng);
  } else {
    std::uniform_real_distribution<double> dist{-1, 1};
    for (uint64_t i = 0; i < n; ++i) out[i] = dist(rng);
  }
  return 0;
}
// permute_matrix (RSVD.hpp:61-71): std::shuffle of 0..n-1 with an UNSEEDED default engine
int pcaone_shuffle_indices(uint64_t n, uint32_t* out) {
  std::vector<int> idx(n);
  for (uint64_t i = 0; i < n; ++i) idx[i] = (int)i;
  auto rng = std::default_random_engine{};
  std::shuffle(idx.data(), idx.data() + n, rng);
  for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t)idx[i];
  return 0;
}

}  // extern "C"
