// pcaone_b200 — launchers of the int8 tensor-core products (tc_gemm.cuh).
#include "ctx.hpp"
#include "tc_gemm.cuh"
#include "emu_fix.cuh"

namespace pcaone {

static_assert(tc::kRowTile == kTcRowTile && tc::kKB == kTcKB && tc::kMaxNP == kTcMaxNP && tc::kChunkBytes == kTcChunkBytes,
              "ctx.hpp mirrors the tile geometry of tc_gemm.cuh");

// ---------------------------------------------------------------- tensor-core (int8 Ozaki) path
uint64_t tc_nkb_samples(const pcaone_ctx* c) { return (c->N + tc::kKB - 1) / tc::kKB; }
uint64_t tc_nrt_samples(const pcaone_ctx* c) { return (c->N + tc::kRowTile - 1) / tc::kRowTile; }
size_t tc_pg_bytes(const pcaone_ctx* c, uint64_t rows) {
  return (size_t)((rows + tc::kRowTile - 1) / tc::kRowTile) * tc_nkb_samples(c) * tc::kChunkBytes;
}
size_t tc_ph_bytes(const pcaone_ctx* c, uint64_t rows) {
  return (size_t)((rows + tc::kKB - 1) / tc::kKB) * tc_nrt_samples(c) * tc::kChunkBytes;
}

// tiled copies of `rows` packed SNP rows at P: PG (rows = SNPs) and PH (rows = samples). `row0`: SNP
// index of P's first row inside the tiling PG / PH point at (0 for a stand-alone block)
void tc_build_tiles(pcaone_ctx* c, const uint8_t* P, uint64_t rows, uint8_t* PG, uint8_t* PH, cudaStream_t st,
                    uint64_t row0) {
  const uint32_t nkb = (uint32_t)tc_nkb_samples(c), nrt = (uint32_t)tc_nrt_samples(c);
  const uint64_t work = (uint64_t)((rows + tc::kRowTile - 1) / tc::kRowTile) * nkb * tc::kRowTile;
  tc::k_tile_rows<<<grid_for(work, 256, c->sms), 256, 0, st>>>(P, c->pitch, rows, (uint32_t)c->N, nkb, PG, row0);
  PCA_CHECK_LAUNCH();
  const uint64_t nkbh = (row0 + rows - 1) / tc::kKB - row0 / tc::kKB + 1;
  tc::k_tile_transpose<<<(unsigned)(nkbh * nrt), 128, 0, st>>>(P, c->pitch, rows, (uint32_t)c->N, nrt, PH, row0);
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

template <int S, int RT, int MODE>
void tc_launch_st(pcaone_ctx* c, const tc::TcGemmArgs& a, int grid) {
  const size_t smem = tc::tc_smem_bytes(RT, c->NP);
  ensure_smem(c, tc::k_tc_gemm<S, RT, MODE>, smem);
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

// ---------------------------------------------------------------- EMU fill on the int8 route (emu_fix.cuh)
bool emu_tc_supported(const pcaone_ctx* c) { return c->emu_tc && c->k <= emu::kMaxK; }

// column tile of the correction kernels: l split into equal tiles of at most 24 columns, rounded up to 4
static int emu_lt(int l) {
  const int nt = (l + emu::kMaxLT - 1) / emu::kMaxLT;
  return std::max(12, ((l + nt - 1) / nt + 3) / 4 * 4);
}

template <int KR, int LT>
void emu_fix_g_t(pcaone_ctx* c, const uint8_t* PG, uint64_t loc0, uint32_t nrows, uint64_t snp0, unsigned long long* w_colmax) {
  constexpr int SB = emu::sb_of(KR, LT);
  constexpr size_t smem = emu::smem_bytes(KR, LT, SB, false);
  const uint32_t rt0 = (uint32_t)(loc0 / tc::kRowTile), rt1 = (uint32_t)((loc0 + nrows - 1) / tc::kRowTile);
  const uint32_t nkb = (uint32_t)tc_nkb_samples(c);
  const uint32_t nblk = (rt1 - rt0 + 1) * (uint32_t)((c->l + LT - 1) / LT);
  // a range of a few row tiles (a winSVD window) would run on a few SMs, every block walking the whole sample axis:
  // slice that axis until about two waves of blocks exist, at least four barrier groups per slice
  uint32_t nsplit = 1;
  if (c->emu_split && nblk < 2u * c->sms) nsplit = std::min<uint32_t>((2u * c->sms + nblk - 1) / nblk, std::max<uint32_t>(1, nkb / (4 * SB)));
  uint32_t kb_per = (nkb + nsplit - 1) / nsplit;
  kb_per = (kb_per + SB - 1) / SB * SB;
  nsplit = (nkb + kb_per - 1) / kb_per;
  double* part = nullptr;
  if (nsplit > 1) {
    const size_t need = (size_t)nsplit * nrows * c->lp;
    if (need > c->emu_part_cap) {
      if (c->d_emu_part) cudaFree(c->d_emu_part);
      c->d_emu_part = nullptr;
      dmalloc(&c->d_emu_part, need);
      c->emu_part_cap = need;
    }
    part = c->d_emu_part;
  }
  const dim3 grid(rt1 - rt0 + 1, (unsigned)((c->l + LT - 1) / LT), nsplit);
  ensure_smem(c, emu::k_emu_fix_g<KR, LT, SB>, smem);
  emu::k_emu_fix_g<KR, LT, SB><<<grid, emu::kThreads, smem, c->stream>>>(
      PG, (uint64_t)nkb * tc::kChunkBytes, nkb, (uint32_t)c->N, loc0, nrows, c->d_emu_us, c->lp, c->k, c->d_V + snp0 * c->lp,
      c->lp, c->d_Omg, c->lp, c->l, c->d_F + snp0, c->lut, c->d_G + snp0 * c->lp, w_colmax, kb_per, part);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
  if (part) {
    const int rpp = 256 / c->lp;
    emu::k_emu_reduce_g<<<(unsigned)std::min<uint64_t>((nrows + rpp - 1) / rpp, (uint64_t)c->sms * 8), 256, 0, c->stream>>>(
        part, nsplit, nrows, c->lp, c->l, c->d_F + snp0, c->lut, c->d_G + snp0 * c->lp, w_colmax);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
  }
}

template <int KR, int LT>
void emu_fix_h_t(pcaone_ctx* c, const uint8_t* PH, uint64_t loc0, uint32_t nrows, uint64_t snp0, double* Hacc, double* Hsum) {
  constexpr int SB = emu::sb_of(KR, LT);
  constexpr size_t smem = emu::smem_bytes(KR, LT, SB, true);
  const uint32_t nrt = (uint32_t)tc_nrt_samples(c);
  const dim3 grid(nrt, (unsigned)((c->l + LT - 1) / LT));
  ensure_smem(c, emu::k_emu_fix_h<KR, LT, SB>, smem);
  emu::k_emu_fix_h<KR, LT, SB><<<grid, emu::kThreads, smem, c->stream>>>(
      PH, nrt, (uint32_t)c->N, loc0, nrows, c->d_emu_us, c->lp, c->k, c->d_V + snp0 * c->lp, c->lp, c->d_G + snp0 * c->lp,
      c->lp, c->l, c->d_F + snp0, c->lut, Hacc, Hsum);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// register widths matched to k and to the column tile (the kernels are bound by the shared-memory reads per missing call)
#define EMU_LT_DISPATCH(FN, KR_, ...)                        \
  switch (emu_lt(c->l)) {                                    \
    case 12: FN<KR_, 12>(__VA_ARGS__); break;                \
    case 16: FN<KR_, 16>(__VA_ARGS__); break;                \
    case 20: FN<KR_, 20>(__VA_ARGS__); break;                \
    default: FN<KR_, 24>(__VA_ARGS__); break;                \
  }
#define EMU_DISPATCH(FN, ...)                                \
  do {                                                       \
    if (c->k <= 6) { EMU_LT_DISPATCH(FN, 6, __VA_ARGS__) }   \
    else if (c->k <= 10) { EMU_LT_DISPATCH(FN, 10, __VA_ARGS__) } \
    else if (c->k <= 16) { EMU_LT_DISPATCH(FN, 16, __VA_ARGS__) } \
    else if (c->k <= 24) { EMU_LT_DISPATCH(FN, 24, __VA_ARGS__) } \
    else if (c->k <= 32) { EMU_LT_DISPATCH(FN, 32, __VA_ARGS__) } \
    else { EMU_LT_DISPATCH(FN, 56, __VA_ARGS__) }            \
  } while (0)

// U o S of the current fill, once per range (U, S are fixed while a computeUSV runs its epochs; N x k: negligible)
static void emu_prepare(pcaone_ctx* c) {
  if (!c->d_emu_us) dmalloc(&c->d_emu_us, c->N * c->lp);
  emu::k_emu_scale_u<<<grid_for(c->N * c->lp, 256, c->sms), 256, 0, c->stream>>>(c->d_U, c->d_S, c->N, c->k, c->lp, c->d_emu_us);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// tensor-core version of range_gemms. PG/PH: tiled operands in which the range starts at local
// row / contraction index `loc0`; snp0 = first SNP of the range in d_G / d_F.
// `miss`: the range contains missing calls -> every product is run as (non-missing counts, mask) pair.
void range_gemms_tc(pcaone_ctx* c, const uint8_t* PG, const uint8_t* PH, uint64_t loc0, uint32_t nrows, uint64_t snp0,
                    double* Hacc, bool accumulate, bool miss) {
  const int mode = miss ? tc::kNonMiss : tc::kPlain;
  const bool emu_fill = miss && c->update && c->cfg.emu;  // (a range without missing calls has nothing to fill)
  unsigned long long* o_colmax = c->d_tcs;
  long long* o_csum = reinterpret_cast<long long*>(c->d_tcs + c->lp);
  unsigned long long* w_colmax = c->d_tcs + 2 * c->lp;
  long long* w_csum = reinterpret_cast<long long*>(c->d_tcs + 3 * c->lp);
  const uint32_t nkb_s = (uint32_t)tc_nkb_samples(c), nrt_s = (uint32_t)tc_nrt_samples(c);
  // sample-sharded job: this rank contracts over ITS samples only; the exact int64 partial sums of
  // the range (and, once per Omega image, the column sums of the Omega integers) are summed over the
  // ranks before the finish kernel, which then writes the same G rows on every rank
  const bool shard = c->shard_samples && c->cfg.world > 1;
  bool omega_csum_local = false;
  {
    Timed t(c, 0);
    if (!c->omega_img_valid) {
      if (!c->omega_colmax_valid) {
        PCA_CUDA(cudaMemsetAsync(c->d_tcs, 0, (size_t)2 * c->lp * sizeof(unsigned long long), c->stream));
        tc::k_tc_colmax<<<grid_for(c->N * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_Omg, c->lp, c->l, 0, c->N, o_colmax);
        PCA_CHECK_LAUNCH();
        c->tm.kernel_launches++;
        if (shard) comm_allreduce_u64_max(c, o_colmax, (uint64_t)c->l);  // one scale per column on every rank
      }
      c->omega_colmax_valid = false;
      tc_slice(c, c->d_Omg, 0, c->N, o_colmax, nullptr, 0, c->d_BimgO, o_csum, nullptr, nullptr);
      c->omega_img_valid = true;
      omega_csum_local = shard;
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
    if (shard) {
      comm_group_begin(c);
      comm_allreduce_i64(c, c->d_Racc + roff, (uint64_t)nrows * c->lp);
      if (miss) comm_allreduce_i64(c, c->d_Racc2 + roff, (uint64_t)nrows * c->lp);
      if (omega_csum_local) comm_allreduce_i64(c, o_csum, (uint64_t)c->l);
      comm_group_end(c);
    }
    tc::k_tc_finish_g<<<(unsigned)std::min<uint64_t>((nrows + tc::kKB - 1) / tc::kKB, (uint64_t)c->sms * 8),
                        tc::tc_pair_threads(c->lp), 0, c->stream>>>(
        c->d_Racc + roff, miss ? c->d_Racc2 + roff : nullptr, nrows, c->l, c->lp, c->slices, c->d_F + snp0, c->lut, o_csum,
        o_colmax, c->d_G + snp0 * c->lp, w_colmax);
    PCA_CHECK_LAUNCH();
    c->tm.gemm_g_launches++;
    c->tm.kernel_launches++;
    // EMU update pass: the missing calls hold clamp(U S V^T) instead of 0 — their FP64 terms, before W is sliced
    if (emu_fill) {
      Timed te(c, 10);
      emu_prepare(c);
      EMU_DISPATCH(emu_fix_g_t, c, PG, loc0, nrows, snp0, w_colmax);
    }
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
    if (emu_fill) {
      Timed te(c, 10);
      EMU_DISPATCH(emu_fix_h_t, c, PH, loc0, nrows, snp0, Hacc, fuse_sum ? c->sum_out : nullptr);
    }
  }
  c->tc_ranges++;
  if (miss) c->tc_miss_ranges++;
  if (emu_fill) c->tc_emu_ranges++;
}

}  // namespace pcaone
