// pcaone_b200 — engine state shared by the translation units behind the C-ABI.
//
//   engine.cu      schedules (computeGandH / computeUSV / EM), block streaming, the HBM tile cache
//   abi.cu         extern "C" entry points of include/pcaone_b200.h
//   launch_tc.cu   int8 tensor-core products (tc_gemm.cuh)
//   launch_fp64.cu FP64 DMMA products on packed / dense / dosage operands (gemm_fp64.cuh, dense_gemm.cuh)
//   launch_orth.cu tall-skinny orthonormalisation + the l x l dense stage (orth_fused.cuh, ...)
//   launch_ld.cu   LD r2 tiles and pruning (ld.cuh)
//   comm.cu        collectives between the ranks of a sharded job (NCCL / host hook)
// Every kernel header is included by exactly ONE of them; the others call the launchers
// declared here.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pcaone_b200.h"
#include "common.cuh"

namespace pcaone {

// tile geometry the host planning needs (the kernels static_assert against these)
constexpr int kFp64TileRows = 128;  // gemm_fp64.cuh kTileRows
constexpr int kTcRowTile = 128, kTcKB = 64, kTcMaxNP = 256;
constexpr uint32_t kTcChunkBytes = kTcRowTile * 16;

struct EvPair {
  cudaEvent_t a, b;
  int kind;  // 0 gemm_g, 1 gemm_h, 2 orth, 3 small, 4 h2d, 5 allreduce, 6 decode, 7 tc_g, 8 tc_h, 9 ld
};

}  // namespace pcaone

struct pcaone_comm;  // comm.cu

struct pcaone_ctx {
  pcaone_config cfg{};
  std::string err;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  int sms = 148;

  uint64_t N = 0, M = 0;
  int k = 0, l = 0, NT = 0, lp = 0;
  uint32_t bpr = 0, pitch = 0;
  pcaone::LutParams lut{};
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
  uint64_t h_row_stride = 0;    // bytes between SNP rows of h_packed (>= bpr: a sample shard of a wider bed)
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
  int64_t staged_blk[2] = {-1, -1};  // block of the plan whose copy is enqueued in d_blk[i] (-1: none)
  bool af_done = false;

  // per-SNP
  double* d_F = nullptr;
  uint32_t* d_nmiss = nullptr;
  uint32_t* d_cnt = nullptr;    // sample-sharded jobs: (c01, c10, c11) per SNP, summed over the ranks
  uint64_t cnt_rows = 0;

  // tall matrices, row-major [rows][lp]
  double *d_Omg0 = nullptr, *d_Omg = nullptr, *d_Omg2 = nullptr, *d_H = nullptr, *d_H1 = nullptr, *d_H2 = nullptr, *d_Bt = nullptr,
         *d_Ucur = nullptr, *d_Upre = nullptr, *d_U = nullptr;
  double *d_G = nullptr, *d_V = nullptr, *d_Vpre = nullptr;
  double* d_S = nullptr;
  double* d_emu_us = nullptr;  // U o S of the EMU fill (N x lp), emu_fix.cuh
  double* d_emu_part = nullptr;  // per-slice partial sums of k_emu_fix_g on short ranges
  size_t emu_part_cap = 0;
  int emu_split = 1;  // slice the sample axis of k_emu_fix_g on short ranges (PCAONE_EMU_SPLIT=0: never)
  double* d_Hpart = nullptr;
  uint32_t max_splits = 1;
  bool have_usv = false, have_omg0 = false;

  // small l x l (ld = lp)
  double *d_W = nullptr, *d_R = nullptr, *d_Rinv = nullptr, *d_T1 = nullptr, *d_T2 = nullptr, *d_T = nullptr,
         *d_Vr = nullptr, *d_Z = nullptr, *d_sigma = nullptr, *d_sign = nullptr, *d_hsign = nullptr, *d_scal = nullptr,
         *d_flipbuf = nullptr;
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
  // HBM cache of the tiled operands of streamed blocks (out-of-core sources): block b of the plan
  // keeps its PG / PH tiles after the first pass if they fit, so later passes read HBM, not the host
  // The cache is ONE tiling of SNPs [0, cache_rows) — blocks write their rows / k-block bits at
  // their global offsets — so cached blocks behave like a resident shard: consecutive blocks with no
  // Omega update between them run as one range.
  uint8_t *d_cache_pg = nullptr, *d_cache_ph = nullptr;
  size_t cache_bytes = 0;
  uint64_t cache_rows = 0;                             // SNPs [0, cache_rows) have a place in the cache
  std::vector<uint8_t> cache_filled;                   // per block of the plan: its tiles are in the cache
  int cache_mode = -1;                                 // -1 undecided, 0 off, 1 on
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
  uint64_t tc_ranges = 0, fp64_ranges = 0, tc_miss_ranges = 0, tc_emu_ranges = 0;
  int half = 3;                                        // which products a range runs: 1 = G rows only, 2 = H only, 3 = both
  bool g_is_q = false;                                 // d_G holds Q = G T after small_stage (else raw G)
  double* d_jscratch = nullptr;                        // eigen-fallback scratch of k_orth_fused
  int fused_orth = 1;                                  // PCAONE_FUSED_ORTH=0 selects the multi-kernel path
  int omega_skip2 = 1;                                 // int8 route: Omega updates may skip the second CholeskyQR pass (PCAONE_OMEGA_SKIP2=0: never)
  int omega_force_full = 0;
  uint64_t omega_update_no = 0;
  int emu_tc = 1;                                      // EMU update passes on the int8 route + FP64 correction over the missing calls (PCAONE_EMU_TC=0: FP64 DMMA kernels)
  int one_shot_q = 1;                                  // int8 route: Omega = H (T1 T2) in one tile product (PCAONE_ORTH_ONE_SHOT=0: two)

  // sharded jobs
  pcaone_allreduce_fn allreduce = nullptr;             // host hook (double sums only)
  void* allreduce_user = nullptr;
  pcaone_allreduce2_fn allreduce2 = nullptr;           // typed host hook (any transport)
  void* allreduce2_user = nullptr;
  pcaone_comm* comm = nullptr;                         // in-library NCCL communicator (comm.cu)
  // peer-memory mailboxes for the in-kernel exchanges of the row-sharded Omega update (orth_fused.cuh)
  double* d_mbox = nullptr;                            // this rank's mailbox [slots][world][l * lp] + flags behind it
  size_t mbox_bytes = 0, mbox_flag_off = 0;
  std::vector<void*> peer_opened;                      // cudaIpcOpenMemHandle'd bases (closed at destroy)
  double** d_peer_mbox = nullptr;                      // device array [world] of mailbox pointers
  unsigned long long** d_peer_flag = nullptr;          // device array [world] of flag-array pointers
  unsigned long long peer_seq = 1;                     // next exchange sequence number
  bool peer_ready = false;
  bool shard_samples = false;                          // rows of X^T (samples) sharded instead of SNPs
  uint64_t N_total = 0;                                // samples of the whole job (== N unless shard_samples)
  uint64_t samp0 = 0;                                  // first global sample of this rank (shard_samples)

  // per-context cache of cudaFuncAttributeMaxDynamicSharedMemorySize (the attribute is per device)
  std::unordered_map<const void*, size_t> smem_attr;

  // measurement
  bool timing = false;
  std::vector<pcaone::EvPair> evs;
  pcaone_timers tm{};
  double last_diff = 0.0;
  int last_epochs = 0;
};

namespace pcaone {

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
inline void dmalloc(T** p, size_t n) {
  PCA_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
}

inline int grid_for(uint64_t work, int threads, int sms) {
  uint64_t b = (work + threads - 1) / threads;
  uint64_t cap = (uint64_t)sms * 16;
  return (int)std::max<uint64_t>(1, std::min(b, cap));
}

// Raise a kernel's dynamic shared-memory limit once per (context, kernel): the attribute lives in
// the device's context, so a process that drives several GPUs must set it on each of them.
template <class K>
inline void ensure_smem(pcaone_ctx* c, K kernel, size_t smem) {
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = c->smem_attr.find(key);
  if (it != c->smem_attr.end() && it->second >= smem) return;
  PCA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  c->smem_attr[key] = smem;
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

inline void zero_async(pcaone_ctx* c, double* p, uint64_t n) { PCA_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), c->stream)); }

// ---- comm.cu: collectives between the ranks of a sharded job (no-ops when world == 1)
enum CommOp { kCommSum = 0, kCommMax = 1 };
void comm_allreduce_f64(pcaone_ctx* c, double* buf, uint64_t count);
void comm_allreduce_i64(pcaone_ctx* c, long long* buf, uint64_t count);
void comm_allreduce_u64_max(pcaone_ctx* c, unsigned long long* buf, uint64_t count);
void comm_allreduce_u32(pcaone_ctx* c, uint32_t* buf, uint64_t count);
void comm_group_begin(pcaone_ctx* c);
void comm_group_end(pcaone_ctx* c);
void comm_destroy(pcaone_ctx* c);

// ---- launch_fp64.cu
void range_gemms_fp64(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, uint64_t snp0, double* Hacc, bool accumulate);
void range_gemms_dense(pcaone_ctx* c, uint64_t r0, uint32_t nrows, double* Hacc, bool accumulate);
void range_gemms_dosage(pcaone_ctx* c, uint64_t r0, uint32_t nrows, double* Hacc, bool accumulate);
void gl_refresh(pcaone_ctx* c, uint64_t r0, uint64_t nrows, bool update, double* E, uint32_t ldd);
void dosage_allele_freq(pcaone_ctx* c);
void dosage_decode(pcaone_ctx* c, uint64_t start, uint64_t B, const LutParams& p, double* out);
void dosage_sqnorm(pcaone_ctx* c, double* out);
int gl_em_maf(pcaone_ctx* c, uint32_t maxiter, double tolmaf);
void dense_transpose_in(pcaone_ctx* c, const double* stage);

// ---- launch_tc.cu
size_t tc_pg_bytes(const pcaone_ctx* c, uint64_t rows);
size_t tc_ph_bytes(const pcaone_ctx* c, uint64_t rows);
void tc_build_tiles(pcaone_ctx* c, const uint8_t* P, uint64_t rows, uint8_t* PG, uint8_t* PH, cudaStream_t st,
                    uint64_t row0 = 0);
void tc_alloc(pcaone_ctx* c, uint64_t max_range_rows, bool miss);
bool emu_tc_supported(const pcaone_ctx* c);
void range_gemms_tc(pcaone_ctx* c, const uint8_t* PG, const uint8_t* PH, uint64_t loc0, uint32_t nrows, uint64_t snp0,
                    double* Hacc, bool accumulate, bool miss);

// ---- launch_orth.cu
void ts_gemm_tn(pcaone_ctx* c, const double* A, int l1, const double* B, int l2, uint64_t rows, double* C, bool sharded_rows);
void ts_rightmult(pcaone_ctx* c, const double* A, int l1, const double* T, int l2, uint64_t rows, double* Out);
void small_matmul(pcaone_ctx* c, const double* A, int tA, const double* B, int tB, int m, int p, int n, double* C);
void update_omega(pcaone_ctx* c, const double* H, bool flip);
void small_stage(pcaone_ctx* c);
double device_mev(pcaone_ctx* c, const double* X, const double* Y, uint64_t rows, bool sharded);
void finalize_usv(pcaone_ctx* c);
void flip_uv(pcaone_ctx* c);
void add2(pcaone_ctx* c, const double* A, const double* B, double* Out, uint64_t n);
void ensure_stage(pcaone_ctx* c, size_t doubles);
void upload_colmajor(pcaone_ctx* c, const double* h, uint64_t rows, int cols, double* d);
void download_colmajor(pcaone_ctx* c, const double* d, uint64_t rows, int cols, double* h);
void colmajor_to_rowmajor(pcaone_ctx* c, const double* src, uint64_t rows, int cols, double* dst);

// ---- launch_ld.cu
void ld_r2(pcaone_ctx* c, const pcaone_ld_source& src, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
           double* r2_out, const double* af, double r2_tol, unsigned char* keep_out);
void residuals_block(pcaone_ctx* c, uint64_t start, uint64_t stop, int ld_stats, float* out);

// ---- engine.cu / sources.cu
void resolve_timers(pcaone_ctx* c);
void compute_gandh(pcaone_ctx* c, int pi);
void compute_usv(pcaone_ctx* c, int p, double tol);
void run_em(pcaone_ctx* c, int* iters_out);
void set_blocks(pcaone_ctx* c, const uint64_t* start, const uint64_t* stop, uint32_t nblocks, uint32_t band_factor);
void perform_op(pcaone_ctx* c, const double* x_in, double* y_out);
void walk_ranges(pcaone_ctx* c);
void allreduce_H(pcaone_ctx* c, double* H);
void gl_grm_standardize(pcaone_ctx* c, double* d_Dc);                                       // launch_fp64.cu
void gl_grm(pcaone_ctx* c, double* C_out, double* Dc_out);                                  // launch_cov.cu
void sample_covariance(pcaone_ctx* c, double* K_out);                                       // launch_cov.cu
int sym_svd(pcaone_ctx* c, const double* A, uint64_t n, double* U_out, double* S_out);     // launch_cov.cu
void xt_times(pcaone_ctx* c, const double* A, uint32_t ncols, double* out, double* sqnorm);
void x_times(pcaone_ctx* c, const double* B, uint32_t ncols, double* out);
void dense_onepass(pcaone_ctx* c, uint32_t p, uint32_t windows, int finder);
void alloc_stream_buffers(pcaone_ctx* c);
const uint8_t* stage_block(pcaone_ctx* c, uint32_t b, int buf, bool wait = true);
void block_af_if_needed(pcaone_ctx* c, const uint8_t* P, uint64_t s, uint64_t nrows);
void allele_freq_rows(pcaone_ctx* c, const uint8_t* P, uint64_t s, uint64_t nrows);
void snp_sqnorm(pcaone_ctx* c, double* out);
uint64_t stage_range_max(pcaone_ctx* c, uint64_t want);
const uint8_t* stage_range(pcaone_ctx* c, uint64_t s0, uint64_t nb);
void cache_plan(pcaone_ctx* c);
void cache_release(pcaone_ctx* c);
void cache_invalidate(pcaone_ctx* c);
bool cache_covers(const pcaone_ctx* c, uint32_t b);

}  // namespace pcaone
