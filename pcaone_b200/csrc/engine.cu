// pcaone_b200 — engine: schedules of the randomized-SVD hot path, block streaming, HBM tile cache.
//
// Host-side state machine of the randomized-SVD hot path, driving the sm_100a kernels:
//   RsvdOpData::computeUSV                 reference src/Halko.cpp:46-97
//   NormalRsvdOpData::computeGandH         reference src/Halko.cpp:99-153
//   FancyRsvdOpData::computeGandH          reference src/Halko.cpp:155-269
//   run_pca_with_halko EM loop             reference src/Halko.cpp:290-319
//   FileBed::read_all / read_block_*       reference src/FilePlink.cpp:26-298
// Nothing here falls back to the CPU: every arithmetic step is a kernel launch.
#include "ctx.hpp"

namespace pcaone {

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
      case 10: c->tm.emu_fix_ms += ms; break;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  c->evs.clear();
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

// does the current pass run its products on the int8 tensor-core kernels? Half passes are GEMV-shaped: FP64
// kernels. EMU update passes fill every missing entry with its own FP64 value: tensor-core products of the
// mean-imputed block plus the FP64 correction over the missing calls (emu_fix.cuh), unless switched off
// (PCAONE_EMU_TC=0) or k is beyond what the correction kernels hold in registers — then the FP64 DMMA kernels.
bool pass_uses_tc(const pcaone_ctx* c) {
  if (c->slices <= 0 || c->half != 3) return false;
  if (c->update && c->cfg.emu) return emu_tc_supported(c);
  return true;
}

// G rows of the range = X^T Omega ; Hacc (+)= X G. `buf` = streamed block buffer holding P, or -1
// when P points into the resident shard; `blk` = block of the plan the range is (streamed sources;
// its tiles live in the HBM cache when they fit), else -1. Ranges without EMU fill run on the int8
// tensor-core kernels when the context was created with a PCAONE_PREC_INT8* mode.
void range_gemms(pcaone_ctx* c, const uint8_t* P, uint32_t nrows, uint64_t snp0, double* Hacc, bool accumulate,
                 int buf, int64_t blk = -1) {
  if (nrows == 0) return;
  if (c->source == PCAONE_SRC_DENSE || c->source == PCAONE_SRC_GL) {
    range_gemms_dense(c, snp0, nrows, Hacc, accumulate);
    return;
  }
  if (c->source == PCAONE_SRC_DOSAGE) {
    range_gemms_dosage(c, snp0, nrows, Hacc, accumulate);
    return;
  }
  const bool use_tc = pass_uses_tc(c);
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
    if (!P) throw std::runtime_error("range_gemms: the FP64 route needs the packed rows");
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
    return;
  }
  uint8_t *PG, *PH;
  if (blk >= 0 && cache_covers(c, (uint32_t)blk)) {
    // the cache is one tiling of SNPs [0, cache_rows): the block sits at its own SNP offset, and a
    // range may span several cached blocks (blk = the first)
    if (!c->cache_filled[blk]) {
      if (!P) throw std::runtime_error("range_gemms: cached block without its packed rows");
      tc_build_tiles(c, P, nrows, c->d_cache_pg, c->d_cache_ph, c->stream, snp0);
      c->cache_filled[blk] = 1;
    }
    range_gemms_tc(c, c->d_cache_pg, c->d_cache_ph, snp0, nrows, snp0, Hacc, accumulate, has_miss);
    return;
  } else {
    if (!P) throw std::runtime_error("range_gemms: streamed block without its packed rows");
    if (!c->d_PGb[buf]) {
      PCA_CUDA(cudaMalloc((void**)&c->d_PGb[buf], tc_pg_bytes(c, c->max_block)));
      PCA_CUDA(cudaMalloc((void**)&c->d_PHb[buf], tc_ph_bytes(c, c->max_block)));
    }
    PG = c->d_PGb[buf];
    PH = c->d_PHb[buf];
    tc_build_tiles(c, P, nrows, PG, PH, c->stream);
  }
  range_gemms_tc(c, PG, PH, 0, nrows, snp0, Hacc, accumulate, has_miss);
}

// does block b of the plan have to come from the host on this pass?
bool ooc_needs_stage(pcaone_ctx* c, uint32_t b) {
  if (c->blk_stop[b] + 1 == c->blk_start[b]) return false;  // empty placeholder
  return !(pass_uses_tc(c) && c->af_done && cache_covers(c, b) && c->cache_filled[b]);
}

// enqueue the host->device copy of block b into its buffer (b & 1) unless it is already there
const uint8_t* ooc_stage(pcaone_ctx* c, uint32_t b) {
  const int buf = (int)(b & 1);
  if (c->staged_blk[buf] != (int64_t)b) {
    stage_block(c, b, buf, false);
    c->staged_blk[buf] = b;
  }
  return c->d_blk[buf];
}

// One block of an out-of-core plan: streamed from the host (double-buffered: the copy of the NEXT
// block is enqueued before this block's products, which may synchronise with the host) unless its
// tiles are already in the HBM cache.
void ooc_block(pcaone_ctx* c, uint32_t b, double* Hacc) {
  const uint64_t s0 = c->blk_start[b], nrows = c->blk_stop[b] + 1 - s0;
  if (nrows == 0) return;
  if (pass_uses_tc(c) && c->cache_mode < 0) {
    // the working set first (accumulators, operand images, stream buffers), the cache takes what is left
    alloc_stream_buffers(c);
    tc_alloc(c, c->max_block, false);
    cache_plan(c);
  }
  const uint32_t nb = (uint32_t)c->blk_start.size();
  const bool need = ooc_needs_stage(c, b);
  const uint8_t* P = need ? ooc_stage(c, b) : nullptr;
  if (b + 1 < nb && ooc_needs_stage(c, b + 1)) ooc_stage(c, b + 1);
  if (!need) {
    range_gemms(c, nullptr, (uint32_t)nrows, s0, Hacc, true, 0, b);
    c->tm.cache_hits++;
    return;
  }
  const int buf = (int)(b & 1);
  PCA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
  block_af_if_needed(c, P, s0, nrows);
  range_gemms(c, P, (uint32_t)nrows, s0, Hacc, true, buf, b);
  PCA_CUDA(cudaEventRecord(c->ev_done[buf], c->stream));
  c->staged_blk[buf] = -1;
}

// blocks [b, e] of the plan as one step: a span of cached blocks runs as ONE range of the cache's
// tiling (like merged windows of a resident shard), anything else block by block
bool ooc_cached(pcaone_ctx* c, size_t b) {
  return pass_uses_tc(c) && c->af_done && cache_covers(c, (uint32_t)b) && c->cache_filled[b];
}
void ooc_span(pcaone_ctx* c, size_t b, size_t e, double* Hacc) {
  if (e > b) {
    const uint64_t s0 = c->blk_start[b], nrows = c->blk_stop[e] + 1 - s0;
    range_gemms(c, nullptr, (uint32_t)nrows, s0, Hacc, true, 0, (int64_t)b);
    c->tm.cache_hits += e - b + 1;
    return;
  }
  ooc_block(c, (uint32_t)b, Hacc);
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

// Sum of the SNP-sharded partial H over the ranks. Sample-sharded jobs own disjoint
// rows of H: nothing to exchange.
void allreduce_H(pcaone_ctx* c, double* H) {
  if (c->cfg.world > 1 && !c->shard_samples) comm_allreduce_f64(c, H, c->N * c->lp);
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
      const size_t nb = c->blk_start.size();
      for (size_t b = 0; b < nb;) {
        size_t e2 = b;
        while (ooc_cached(c, b) && e2 + 1 < nb && ooc_cached(c, e2 + 1) && c->blk_start[e2 + 1] == c->blk_stop[e2] + 1) ++e2;
        ooc_span(c, b, e2, c->d_H);
        b = e2 + 1;
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
    if (!ooc || ooc_cached(c, b)) {  // (streamed source: blocks whose tiles sit in the HBM cache)
      while (!steps[e].update && e + 1 < steps.size() && steps[e + 1].target == steps[b].target &&
             steps[e + 1].stop >= steps[e + 1].start && steps[e].stop >= steps[e].start &&
             steps[e + 1].start == steps[e].stop + 1 && (!ooc || ooc_cached(c, e + 1)))
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
        ooc_span(c, b, e, Hacc);
      }
    }
    c->sum_other = nullptr;
    c->sum_out = nullptr;
    if (last.update) {
      if (!c->sum_done) {
        add2(c, c->d_H1, c->d_H2, c->d_H, HN);
      }
      allreduce_H(c, c->d_H);
      update_omega(c, c->d_H, true);
      zero_async(c, last.zero_h1 ? c->d_H1 : c->d_H2, HN);
    }
    b = e + 1;
  }
  if (ooc) c->af_done = true;
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
      diff = 1.0 - device_mev(c, c->d_Ucur, c->d_Upre, c->N, c->shard_samples);
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
    for (uint32_t b = 0; b < c->blk_start.size(); ++b) ooc_block(c, b, c->d_H);
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
      snp_sqnorm(c, c->d_stage);
    } else if (c->source == PCAONE_SRC_DOSAGE) {
      dosage_sqnorm(c, c->d_stage);
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
      add2(c, c->d_H1, c->d_H2, c->d_H, HN);
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
    const double diff = 1.0 - device_mev(c, c->d_V, c->d_Vpre, c->M, !c->shard_samples);
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
    if (start[i] == stop[i] + 1) continue;  // empty placeholder window (a sharded job keeps every rank on one schedule)
    if (stop[i] >= c->M || start[i] > stop[i]) throw std::runtime_error("set_blocks: block out of range");
    mb = std::max(mb, stop[i] - start[i] + 1);
  }
  if (mb > c->max_block && c->d_blk[0]) throw std::runtime_error("set_blocks: cannot grow blocks after streaming began");
  c->max_block = std::max(c->max_block, mb);
  cache_release(c);  // the cache is laid out per block of the plan
}

}  // namespace pcaone
