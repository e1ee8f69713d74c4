// pcaone_b200 — LD r2 tiles and greedy pruning (ld.cuh) behind pcaone_ld_r2 / pcaone_ld_prune.
#include "ctx.hpp"
#include "ld.cuh"

namespace pcaone {

// ---------------------------------------------------------------- LD r2 (LD.cpp:450-473)
// Banded tile Gram on the FP64 tensor cores (ld.cuh). The SNP axis is walked in chunks of lead
// SNPs (+ a halo of the widest window) sized to the free HBM, so M x N need not fit at once.
// With `keep_out` the r^2 values stay on the device and feed the greedy pruning kernel chunk by
// chunk (ld_prune_big, LD.cpp:240-268); r2_out may then be NULL.
namespace {

// rows [c0, c0 + rows) of the LD operand -> d_Gs [rows][Np] (column-centred doubles, zero padded)
//   DENSE_F64      host doubles N x M column-major (what data->G holds in the reference)
//   PACKED         resident packed shard, centred by F, unscaled (read_block_initial, standardize = false)
//   RESID_F32      host float32 rows of a .residuals file: cast, centre (FileBin::read_all, FileBinary.cpp:21-30)
//   PACKED_RESID   resident packed shard -> the residuals `--ld` would write and `-B` read back, without the
//                  file: decode, (ld_stats 0) -= U S V^T, centre, round through float32, centre
//                  (Data.cpp:242-291 + FileBinary.cpp:21-30)
//   PACKED_PROJECT resident packed shard -> (I - U U^T) G with the U of --USV (LD.cpp:491-496)
void ld_operand(pcaone_ctx* c, const pcaone_ld_source& src, uint64_t c0, uint64_t rows, uint32_t Np, double* d_Gs,
                void* d_raw, const double* d_Uproj, float* f32_out, const uint8_t* Ppacked = nullptr) {
  const uint8_t* Prows = Ppacked ? Ppacked : c->d_packed + c0 * c->pitch;  // packed rows of SNP c0.. (pitch layout)
  const uint64_t N = c->N;
  LutParams lut = c->lut;
  lut.standardize = 0;  // --ld runs centred, unscaled genotypes (Halko.cpp:283-288)
  switch (src.kind) {
    case PCAONE_LD_DENSE_F64: {
      const double* G = reinterpret_cast<const double*>(src.data);
      PCA_CUDA(cudaMemcpyAsync(d_raw, G + c0 * N, rows * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      c->tm.h2d_bytes += rows * N * sizeof(double);
      ld::k_pad_rows<<<grid_for(rows * Np, 256, c->sms), 256, 0, c->stream>>>(reinterpret_cast<double*>(d_raw), rows,
                                                                             (uint32_t)N, Np, d_Gs);
      PCA_CHECK_LAUNCH();
      break;
    }
    case PCAONE_LD_RESID_F32: {
      const float* R = reinterpret_cast<const float*>(src.data);
      PCA_CUDA(cudaMemcpyAsync(d_raw, R + c0 * N, rows * N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      c->tm.h2d_bytes += rows * N * sizeof(float);
      ld::k_pad_rows_f32<<<grid_for(rows * Np, 256, c->sms), 256, 0, c->stream>>>(reinterpret_cast<float*>(d_raw), rows,
                                                                                 (uint32_t)N, Np, d_Gs);
      PCA_CHECK_LAUNCH();
      ld::k_center_rows<<<grid_for(rows * 32, 256, c->sms), 256, 0, c->stream>>>(d_Gs, rows, (uint32_t)N, Np, 0, nullptr);
      PCA_CHECK_LAUNCH();
      break;
    }
    case PCAONE_LD_PACKED:
      ld::k_decode_rows<<<grid_for(rows * (Np >> 2), 256, c->sms), 256, 0, c->stream>>>(
          Prows, c->pitch, (uint32_t)N, Np, rows, c->d_F + c0, lut, d_Gs);
      PCA_CHECK_LAUNCH();
      break;
    case PCAONE_LD_PACKED_RESID: {
      const int nk = src.ld_stats == 0 ? c->k : 0;
      if (nk && !c->have_usv) throw std::runtime_error("ld: ancestry-adjusted residuals need U, S, V (run the PCA first)");
      const size_t smem = ((size_t)ld::kResSamples * (nk + 1) + (size_t)ld::kResSnps * nk) * sizeof(double);
      ensure_smem(c, ld::k_resid_rows, std::max<size_t>(smem, 1));
      const unsigned grid = (unsigned)(((rows + ld::kResSnps - 1) / ld::kResSnps) * ((Np + ld::kResSamples - 1) / ld::kResSamples));
      ld::k_resid_rows<<<grid, 256, smem, c->stream>>>(Prows, c->pitch, (uint32_t)N, Np, rows,
                                                       c->d_F + c0, lut, c->d_U, c->lp, c->d_S, c->d_V + c0 * c->lp, c->lp,
                                                       nk, d_Gs);
      PCA_CHECK_LAUNCH();
      ld::k_center_rows<<<grid_for(rows * 32, 256, c->sms), 256, 0, c->stream>>>(d_Gs, rows, (uint32_t)N, Np, 1, f32_out);
      PCA_CHECK_LAUNCH();
      break;
    }
    case PCAONE_LD_PACKED_PROJECT:
      ld::k_decode_rows<<<grid_for(rows * (Np >> 2), 256, c->sms), 256, 0, c->stream>>>(
          c->d_packed + c0 * c->pitch, c->pitch, (uint32_t)N, Np, rows, c->d_F + c0, lut, d_Gs);
      PCA_CHECK_LAUNCH();
      ld::k_project_out<<<grid_for(rows * 32, 256, c->sms), 256, 0, c->stream>>>(d_Gs, rows, (uint32_t)N, Np, d_Uproj,
                                                                                  (int)src.ncols, (int)src.ncols);
      PCA_CHECK_LAUNCH();
      break;
    default: throw std::runtime_error("ld: unknown operand source");
  }
  c->tm.kernel_launches += 2;
}

}  // namespace

// Data::write_residuals (Data.cpp:242-291) for SNPs [start, stop] of the resident shard: the float32 rows
// the reference writes to <out>.residuals (out: [stop - start + 1][N] floats, SNP-major like the file).
void residuals_block(pcaone_ctx* c, uint64_t start, uint64_t stop, int ld_stats, float* out) {
  const bool streamed = c->source == PCAONE_SRC_HOST || c->source == PCAONE_SRC_FILE;
  if (!(c->source == PCAONE_SRC_RESIDENT || streamed) || !c->af_done)
    throw std::runtime_error("residuals_block: needs a packed genotype source with allele frequencies");
  if (stop < start || stop >= c->M) throw std::runtime_error("residuals_block: range out of bounds");
  const uint64_t N = c->N, B = stop - start + 1;
  const uint32_t Np = (uint32_t)round_up(N, 16);
  uint64_t piece = std::max<uint64_t>(1, std::min<uint64_t>(B, (512ull << 20) / (Np * 8ull)));
  if (streamed) piece = stage_range_max(c, piece);   // streamed blocks pass through the plan's staging buffer
  double* d_Gs = nullptr;
  float* d_f = nullptr;
  try {
    dmalloc(&d_Gs, piece * Np);
    dmalloc(&d_f, piece * N);
    pcaone_ld_source src{};
    src.kind = PCAONE_LD_PACKED_RESID;
    src.ld_stats = ld_stats;
    for (uint64_t s0 = start; s0 <= stop; s0 += piece) {
      const uint64_t nb = std::min<uint64_t>(piece, stop - s0 + 1);
      const uint8_t* P = streamed ? stage_range(c, s0, nb) : nullptr;
      ld_operand(c, src, s0, nb, Np, d_Gs, nullptr, nullptr, d_f, P);
      if (streamed) PCA_CUDA(cudaEventRecord(c->ev_done[0], c->stream));
      PCA_CUDA(cudaMemcpyAsync(out + (s0 - start) * N, d_f, nb * N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
      c->tm.d2h_bytes += nb * N * sizeof(float);
    }
  } catch (...) {
    cudaFree(d_Gs);
    cudaFree(d_f);
    throw;
  }
  cudaFree(d_Gs);
  cudaFree(d_f);
}

void ld_r2(pcaone_ctx* c, const pcaone_ld_source& src, uint64_t nsnps, const int32_t* ws, const int32_t* we, uint64_t nwin,
           double* r2_out, const double* af, double r2_tol, unsigned char* keep_out) {
  const uint64_t N = c->N;
  if (N < 2) throw std::runtime_error("ld_r2: needs at least two samples");
  const bool host_rows = src.kind == PCAONE_LD_DENSE_F64 || src.kind == PCAONE_LD_RESID_F32;
  const double* G = src.kind == PCAONE_LD_DENSE_F64 ? reinterpret_cast<const double*>(src.data) : nullptr;
  if (host_rows && !src.data) throw std::runtime_error("ld_r2: host operand is NULL");
  if (!host_rows) {
    if (c->source != PCAONE_SRC_RESIDENT || !c->af_done)
      throw std::runtime_error("ld_r2: this operand source needs a resident packed shard with allele frequencies");
    if (nsnps != c->M) throw std::runtime_error("ld_r2: nsnps must equal the resident shard size");
  }
  if (src.kind == PCAONE_LD_PACKED_PROJECT && (!src.data || src.ncols == 0 || src.ncols > 64))
    throw std::runtime_error("ld_r2: (I - U U^T) G needs U (nsamples x ncols, ncols <= 64)");
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
  const double per_row = (double)Np * 8 + (host_rows ? (double)N * (G ? 8 : 4) : 0.0) + 16.0;
  const double per_lead = per_row + (double)(maxwe - 1) * 8;
  const uint64_t halo = maxwe;  // rows beyond the last lead of a chunk
  double leads_d = (budget - (double)halo * per_row) / per_lead;
  if (const char* e = getenv("PCAONE_LD_CHUNK")) leads_d = atof(e);  // test hook: force small chunks
  if (leads_d < ld::kTile) throw std::runtime_error("ld_r2: not enough device memory for one tile row of this window width");
  const uint64_t leads = std::min<uint64_t>(round_up(nsnps, ld::kTile), (uint64_t)leads_d / ld::kTile * ld::kTile);
  const uint64_t max_rows = std::min<uint64_t>(nsnps, leads + halo);

  double *d_Gs = nullptr, *d_isd = nullptr, *d_out = nullptr, *d_Uproj = nullptr;
  void* d_raw = nullptr;
  int32_t *d_winof = nullptr, *d_we = nullptr, *d_ws = nullptr;
  unsigned char* d_keep = nullptr;
  double* d_af = nullptr;
  uint64_t* d_offs = nullptr;
  int2* d_tiles = nullptr;
  size_t out_cap = 0, tiles_cap = 0;
  auto cleanup = [&]() {
    cudaFree(d_Gs); cudaFree(d_raw); cudaFree(d_isd); cudaFree(d_out); cudaFree(d_Uproj);
    cudaFree(d_winof); cudaFree(d_we); cudaFree(d_offs); cudaFree(d_tiles);
    cudaFree(d_ws); cudaFree(d_keep); cudaFree(d_af);
  };
  try {
    dmalloc(&d_Gs, max_rows * Np);
    if (host_rows) PCA_CUDA(cudaMalloc(&d_raw, max_rows * N * (G ? sizeof(double) : sizeof(float))));
    if (src.kind == PCAONE_LD_PACKED_PROJECT) {  // U of --USV: host N x ncols column-major -> [N][ncols] row-major
      std::vector<double> ur((size_t)N * src.ncols);
      const double* Uh = reinterpret_cast<const double*>(src.data);
      for (uint64_t i = 0; i < N; ++i)
        for (uint32_t k = 0; k < src.ncols; ++k) ur[i * src.ncols + k] = Uh[(size_t)k * N + i];
      dmalloc(&d_Uproj, ur.size());
      PCA_CUDA(cudaMemcpyAsync(d_Uproj, ur.data(), ur.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      PCA_CUDA(cudaStreamSynchronize(c->stream));
    }
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
    ensure_smem(c, ld::k_ld_tiles, ld::kSmemBytes);
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
      ld_operand(c, src, c0, rows, Np, d_Gs, d_raw, d_Uproj, nullptr);
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

}  // namespace pcaone
