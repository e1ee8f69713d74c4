// pcaone_b200 — genotype sources: resident packed shards, out-of-core block streaming from pinned
// host memory / a .bed file, the HBM cache of streamed tiles, allele frequencies and the dense block
// decode (decode.cuh).
//   Data::prepare block plan          reference src/Data.cpp:14-85
//   FileBed::read_all                 reference src/FilePlink.cpp:26-120
//   FileBed::read_block_initial       reference src/FilePlink.cpp:122-218
//   FileBed::read_block_update        reference src/FilePlink.cpp:220-298
#include "ctx.hpp"
#include "decode.cuh"

namespace pcaone {

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

// enqueue the H2D of block b into buffer `buf`; returns the device pointer (pitch layout). `wait`: the
// compute stream waits for the copy right away (else the caller does, when it consumes the block)
const uint8_t* stage_block(pcaone_ctx* c, uint32_t b, int buf, bool wait) {
  const uint64_t s = c->blk_start[b], e = c->blk_stop[b];
  const uint64_t nrows = e - s + 1;
  const size_t bytes = nrows * c->bpr;
  PCA_CUDA(cudaEventSynchronize(c->ev_done[buf]));  // previous user of this buffer finished
  const uint8_t* src;
  uint64_t src_stride = c->bpr;
  if (c->source == PCAONE_SRC_HOST) {
    src_stride = c->h_row_stride ? c->h_row_stride : c->bpr;
    src = c->h_packed + s * src_stride;
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
  if (src_stride == c->bpr)
    PCA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
  else  // this context's byte columns of a wider bed (sample shard)
    PCA_CUDA(cudaMemcpy2DAsync(dst, c->bpr, src, src_stride, c->bpr, nrows, cudaMemcpyHostToDevice, c->copy_stream));
  c->tm.h2d_bytes += bytes;
  if (c->pitch != c->bpr) {
    k_repitch<<<grid_for(nrows * (c->pitch >> 4), 256, c->sms), 256, 0, c->copy_stream>>>(c->d_raw[buf], c->d_blk[buf],
                                                                                         nrows, c->bpr, c->pitch);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches++;
  }
  PCA_CUDA(cudaEventRecord(c->ev_copied[buf], c->copy_stream));
  if (wait) PCA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
  return c->d_blk[buf];
}

// Allele frequencies + missing counts of `nrows` packed rows at P into d_F / d_nmiss [s, s + nrows).
// Sample-sharded jobs sum the three genotype counts over the ranks first (exact integers), so F is
// the same single division as on one GPU.
void allele_freq_rows(pcaone_ctx* c, const uint8_t* P, uint64_t s, uint64_t nrows) {
  const int grid = grid_for(nrows * 32, 256, c->sms);
  if (c->shard_samples && c->cfg.world > 1) {
    if (nrows > c->cnt_rows) {
      if (c->d_cnt) cudaFree(c->d_cnt);
      dmalloc(&c->d_cnt, 3 * nrows);
      c->cnt_rows = nrows;
    }
    k_allele_freq<<<grid, 256, 0, c->stream>>>(P, c->pitch, (uint32_t)c->N, nrows, nullptr, nullptr, c->d_cnt);
    PCA_CHECK_LAUNCH();
    comm_allreduce_u32(c, c->d_cnt, 3 * nrows);
    k_af_from_counts<<<grid_for(nrows, 256, c->sms), 256, 0, c->stream>>>(c->d_cnt, c->N_total, nrows, c->d_F + s,
                                                                          c->d_nmiss + s);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches += 2;
    return;
  }
  k_allele_freq<<<grid, 256, 0, c->stream>>>(P, c->pitch, (uint32_t)c->N, nrows, c->d_F + s, c->d_nmiss + s);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

void block_af_if_needed(pcaone_ctx* c, const uint8_t* P, uint64_t s, uint64_t nrows) {
  if (c->af_done) return;
  allele_freq_rows(c, P, s, nrows);
}

// Packed rows of SNPs [s0, s0 + nb) of a streamed source on the device (buffer 0, pitch layout), outside
// the block plan: decode_block / residuals_block ask for arbitrary ranges. nb <= stage_range_max(c).
// The caller records ev_done[0] on the compute stream when it is done with the rows.
uint64_t stage_range_max(pcaone_ctx* c, uint64_t want) {
  if (c->d_blk[0]) return std::min<uint64_t>(want, c->max_block);  // buffers are sized by the plan once streaming began
  c->max_block = std::max(c->max_block, want);
  return want;
}
const uint8_t* stage_range(pcaone_ctx* c, uint64_t s0, uint64_t nb) {
  alloc_stream_buffers(c);
  std::vector<uint64_t> sv = c->blk_start, ev = c->blk_stop;
  c->blk_start = {s0};
  c->blk_stop = {s0 + nb - 1};
  const uint8_t* P = nullptr;
  try {
    P = stage_block(c, 0, 0);
  } catch (...) {
    c->blk_start = sv;
    c->blk_stop = ev;
    throw;
  }
  c->blk_start = sv;
  c->blk_stop = ev;
  c->staged_blk[0] = -1;
  return P;
}

void snp_sqnorm(pcaone_ctx* c, double* out) {
  k_snp_sqnorm<<<grid_for(c->M * 32, 256, c->sms), 256, 0, c->stream>>>(c->d_packed, c->pitch, (uint32_t)c->N, c->M,
                                                                       c->d_F, c->lut, out);
  PCA_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- HBM cache of streamed tiles
// An out-of-core source is streamed from the host on the first pass; the int8 route re-tiles every
// block into its two operand layouts (PG: rows = SNPs, PH: rows = samples) anyway, and those tiles
// are what later passes read. When they fit next to the working set they stay in HBM (2 x the
// packed bytes: 125 GB for the 500k x 500k bed on a 180 GB B200), so passes 2.. run at tensor speed
// instead of at the host link's. Blocks that do not fit keep streaming. PCAONE_TILE_CACHE=0 turns
// the cache off, PCAONE_TILE_CACHE_MB caps it.
void cache_release(pcaone_ctx* c) {
  if (c->d_cache_pg) cudaFree(c->d_cache_pg);
  if (c->d_cache_ph) cudaFree(c->d_cache_ph);
  c->d_cache_pg = c->d_cache_ph = nullptr;
  c->cache_bytes = 0;
  c->cache_rows = 0;
  c->cache_filled.clear();
  c->cache_mode = -1;
}

void cache_invalidate(pcaone_ctx* c) {
  std::fill(c->cache_filled.begin(), c->cache_filled.end(), 0);
  c->staged_blk[0] = c->staged_blk[1] = -1;
}

bool cache_covers(const pcaone_ctx* c, uint32_t b) {
  return c->cache_mode == 1 && c->blk_stop[b] + 1 > c->blk_start[b] && c->blk_stop[b] < c->cache_rows;
}

void cache_plan(pcaone_ctx* c) {
  if (c->cache_mode >= 0) return;
  c->cache_mode = 0;
  const size_t nb = c->blk_start.size();
  c->cache_filled.assign(nb, 0);
  if (const char* e = getenv("PCAONE_TILE_CACHE"))
    if (atoi(e) == 0) return;
  if (c->slices == 0 || c->cfg.emu) return;  // FP64 kernels (and EMU update passes) read the packed rows
  size_t free_b = 0, total_b = 0;
  PCA_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t reserve = (size_t)3 << 30;
  const size_t dbl = 2 * (tc_pg_bytes(c, c->max_block) + tc_ph_bytes(c, c->max_block));  // uncached blocks' double buffers
  auto need = [&](uint64_t rows) { return tc_pg_bytes(c, rows) + tc_ph_bytes(c, rows) + 512; };
  size_t budget = free_b > reserve ? free_b - reserve : 0;
  if (const char* e = getenv("PCAONE_TILE_CACHE_MB")) budget = std::min<size_t>(budget, (size_t)atoll(e) << 20);
  // the longest prefix of the plan (contiguous blocks from SNP 0) whose tiling fits
  uint64_t rows = 0;
  for (size_t b = 0; b < nb; ++b) {
    if (c->blk_stop[b] + 1 == c->blk_start[b]) continue;  // empty placeholder
    if (c->blk_start[b] != rows) break;
    const uint64_t upto = c->blk_stop[b] + 1;
    const bool last = upto == c->M;
    if (need(upto) + (last ? 0 : dbl) > budget) break;
    rows = upto;
  }
  if (rows == 0) return;
  PCA_CUDA(cudaMalloc((void**)&c->d_cache_pg, tc_pg_bytes(c, rows)));
  PCA_CUDA(cudaMalloc((void**)&c->d_cache_ph, tc_ph_bytes(c, rows)));
  // zero once: bits of SNP slots that no block fills (the tail of the last k-block) stay code 00
  PCA_CUDA(cudaMemsetAsync(c->d_cache_pg, 0, tc_pg_bytes(c, rows), c->stream));
  PCA_CUDA(cudaMemsetAsync(c->d_cache_ph, 0, tc_ph_bytes(c, rows), c->stream));
  c->cache_bytes = tc_pg_bytes(c, rows) + tc_ph_bytes(c, rows);
  c->cache_rows = rows;
  c->cache_mode = 1;
}

}  // namespace pcaone

using namespace pcaone;

extern "C" {

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

int pcaone_set_host_source2(pcaone_ctx* c, const uint8_t* packed, uint64_t nsnps, uint64_t row_stride) {
  CTX_GUARD(c, {
    if (nsnps != c->M) throw std::runtime_error("set_host_source: nsnps does not match the context");
    if (row_stride != 0 && row_stride < c->bpr) throw std::runtime_error("set_host_source: row stride below ceil(N/4)");
    c->h_packed = packed;
    c->h_row_stride = row_stride;
    c->source = PCAONE_SRC_HOST;
    c->af_done = false;
    c->h_nmiss.clear();
    c->nmiss_prefix.clear();
    cache_invalidate(c);  // new data behind the same plan: stream it again
  });
}
int pcaone_set_host_source(pcaone_ctx* c, const uint8_t* packed, uint64_t nsnps) {
  return pcaone_set_host_source2(c, packed, nsnps, 0);
}

int pcaone_set_reader_source(pcaone_ctx* c, pcaone_read_block_fn fn, void* user) {
  CTX_GUARD(c, {
    c->reader = fn;
    c->reader_user = user;
    c->source = PCAONE_SRC_FILE;
    c->af_done = false;
    cache_invalidate(c);
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
    cache_invalidate(c);
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
      allele_freq_rows(c, c->d_packed, 0, c->M);
    } else if (c->source == PCAONE_SRC_DOSAGE) {
      Timed t(c, 6);
      dosage_allele_freq(c);
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
        c->staged_blk[buf] = -1;
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
      dosage_decode(c, start, B, p, out);
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
    LutParams p = c->lut;
    p.standardize = (standardize && c->cfg.scale == -9) ? 1 : 0;
    const int emu = (update && c->cfg.emu) ? 1 : 0;
    if (emu && !c->have_usv) throw std::runtime_error("decode_block(update) without U,S,V");
    // a streamed source is staged through buffer 0 in pieces no larger than the plan's blocks
    const bool streamed = c->source != PCAONE_SRC_RESIDENT;
    const uint64_t piece = streamed ? stage_range_max(c, B) : B;
    ensure_stage(c, c->N * piece);
    for (uint64_t s0 = start; s0 <= stop; s0 += piece) {
      const uint64_t nb = std::min<uint64_t>(piece, stop - s0 + 1);
      const uint8_t* P = streamed ? stage_range(c, s0, nb) : c->d_packed + s0 * c->pitch;
      if (!c->af_done) allele_freq_rows(c, P, s0, nb);
      Timed t(c, 6);
      k_decode_block<<<grid_for(((c->N + 3) / 4) * nb, 256, c->sms), 256, 0, c->stream>>>(
          P, c->pitch, (uint32_t)c->N, (uint32_t)nb, c->d_F + s0, p, emu, c->d_U, c->lp, c->d_S, c->d_V + s0 * c->lp,
          c->lp, c->k, c->d_stage);
      PCA_CHECK_LAUNCH();
      c->tm.kernel_launches++;
      if (streamed) PCA_CUDA(cudaEventRecord(c->ev_done[0], c->stream));
      PCA_CUDA(cudaMemcpyAsync(out + (s0 - start) * c->N, c->d_stage, c->N * nb * sizeof(double), cudaMemcpyDeviceToHost,
                               c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      c->tm.d2h_bytes += c->N * nb * sizeof(double);
    }
  });
}

}  // extern "C"
