// pcaone_b200 — launchers of the tall-skinny orthonormalisation and the l x l dense stage.
#include "ctx.hpp"
#include "small_dense.cuh"
#include "tall_skinny.cuh"
#include "orth_fused.cuh"

namespace pcaone {

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
  if (sharded_rows) comm_allreduce_f64(c, C, (uint64_t)c->lp * c->lp);
}

template <int RN>
void rightmult_r(pcaone_ctx* c, const double* A, int l1, const double* T, int l2, uint64_t rows, double* Out) {
  const size_t smem = ((size_t)l1 * 16 * RN + (size_t)64 * (l1 + 1)) * sizeof(double);
  ensure_smem(c, k_ts_rightmult<RN>, smem);
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
  PCA_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PCA_CUDA(cudaStreamSynchronize(c->stream));
  // d_status[1]: eigen-route count of k_orth_fused, bit 0x40000000 = an in-kernel peer exchange gave up waiting
  if (c->h_status[1] & 0x40000000)
    throw std::runtime_error("a rank of the job never reached an in-kernel exchange of the Omega update (peer mailbox timeout)");
  return c->h_status[0];
}

void small_matmul(pcaone_ctx* c, const double* A, int tA, const double* B, int tB, int m, int p, int n, double* C) {
  k_small_matmul<<<1, 1024, 0, c->stream>>>(A, tA, B, tB, m, p, n, c->lp, C);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

void jacobi(pcaone_ctx* c, const double* A, int sym, double* sigma, double* V) {
  const size_t smem = 2 * (size_t)c->l * c->l * sizeof(double);
  ensure_smem(c, k_jacobi_svd, smem);
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
  ensure_smem(c, k_chol_inv, smem);
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
  ensure_smem(c, k_orth_fused<R>, smem);
  void* args[] = {(void*)&a};
  PCA_CUDA(cudaLaunchCooperativeKernel((void*)k_orth_fused<R>, dim3(c->sms), dim3(kOrthThreads), args, smem, c->stream));
  c->tm.kernel_launches++;
}

bool orth_fused_ok(const pcaone_ctx* c) { return c->fused_orth && c->l <= kOrthMaxL; }

// One cooperative launch: Q = orth(A) (CholeskyQR2) [+ Householder signs] [+ flipOmg against Q2].
void orth_fused(pcaone_ctx* c, const double* A, uint64_t rows, double* Q, double* Q2, double* Ttot, bool signs,
                bool flip, int phases = 7, unsigned long long* colmax_out = nullptr, double* flipbuf = nullptr) {
  OrthArgs a{};
  a.flipbuf = flipbuf;
  a.one_shot = (Q != nullptr && c->slices > 0 && c->one_shot_q) ? 1 : 0;
  if (flipbuf && phases == 7) {  // single-launch row-sharded update: in-kernel exchanges over peer memory
    if (!c->peer_ready) throw std::runtime_error("orth_fused: peer mailboxes are not set up");
    a.peer_world = c->cfg.world;
    a.peer_rank = c->cfg.rank;
    a.peer_mbox = c->d_peer_mbox;
    a.peer_flag = c->d_peer_flag;
    a.peer_seq = c->peer_seq;
    c->peer_seq += 3;
  }
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
  a.skip_thresh = 1e5;
  if (Q && c->slices > 0 && c->omega_skip2 && !qr2_always) {  // Omega update on the int8 route (see OrthArgs::skip_thresh)
    a.skip2 = (phases == 7 || phases == 2) ? c->d_status + 3 : nullptr;
    a.skip_thresh = 3e-7 / 2.220446049250313e-16;
    a.force_full = c->omega_force_full;
    a.veto = c->d_status + 4;
    a.veto_tol = 3e-7;
  }
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
    auto reduce_W = [&]() { comm_allreduce_f64(c, c->d_W, (uint64_t)c->l * c->lp); };
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

// The same update when H / Omega are sharded by rows over the ranks (sample-sharded jobs): the fused
// kernel in four launches. Exchanged: the two l x l Gram matrices of CholeskyQR2, then ONE grouped
// reduction of {flipOmg column sums, the Householder signs of the rank that owns the top l rows,
// the column maxima of |Omega| for the int8 slicing} — 2 l^2 + 4 l numbers per update instead of N x l.
void update_omega_sharded(pcaone_ctx* c, const double* H, bool flip) {
  if (!orth_fused_ok(c))
    throw std::runtime_error("sample-sharded jobs need the fused orthonormalisation (k + oversamples <= 80)");
  const bool top = c->samp0 == 0;
  if (top && c->N < (uint64_t)c->l) throw std::runtime_error("sample shard of rank 0 is shorter than k + oversamples");
  unsigned long long* cm = (c->slices > 0 && c->d_tcs) ? c->d_tcs : nullptr;
  double* Q2 = flip ? c->d_Omg2 : nullptr;
  if (c->peer_ready) {
    // ONE cooperative launch: the Gram / flip exchanges run inside the kernel over peer memory (NVLink stores into
    // the peers' mailboxes), no NCCL launch and no extra kernel launch per exchange
    orth_fused(c, H, c->N, c->d_Omg, Q2, nullptr, top, flip, 7, cm, c->d_flipbuf);
    c->tm.omega_updates++;
    c->omega_img_valid = false;
    c->omega_colmax_valid = cm != nullptr;
    return;
  }
  orth_fused(c, H, c->N, c->d_Omg, Q2, nullptr, false, false, 1);
  comm_allreduce_f64(c, c->d_W, (uint64_t)c->l * c->lp);
  orth_fused(c, H, c->N, c->d_Omg, Q2, nullptr, false, false, 2);
  comm_allreduce_f64(c, c->d_W, (uint64_t)c->l * c->lp);
  if (cm) PCA_CUDA(cudaMemsetAsync(cm, 0, (size_t)2 * c->lp * sizeof(unsigned long long), c->stream));
  orth_fused(c, H, c->N, c->d_Omg, Q2, nullptr, top, flip, 4, cm, c->d_flipbuf);
  comm_group_begin(c);
  comm_allreduce_f64(c, c->d_flipbuf, (uint64_t)3 * c->l);
  if (cm) comm_allreduce_u64_max(c, cm, (uint64_t)c->l);
  comm_group_end(c);
  orth_fused(c, H, c->N, c->d_Omg, Q2, nullptr, false, flip, 8, nullptr, c->d_flipbuf);
  c->tm.omega_updates++;
  c->omega_img_valid = false;
  c->omega_colmax_valid = cm != nullptr;
}

// Omega = thinQ(H) (+ flipOmg)   Halko.cpp:120-124 / 208-213
void update_omega(pcaone_ctx* c, const double* H, bool flip) {
  Timed t(c, 2);
  c->omega_force_full = (c->omega_update_no++ % 16 == 0) ? 1 : 0;  // every 16th update measures what skipping leaves out
  if (c->shard_samples && c->cfg.world > 1) {
    update_omega_sharded(c, H, flip);
    return;
  }
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
  ensure_smem(c, k_householder_signs, smem);
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

void add2(pcaone_ctx* c, const double* A, const double* B, double* Out, uint64_t n) {
  k_add2<<<grid_for(n, 256, c->sms), 256, 0, c->stream>>>(A, B, Out, n);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches++;
}

// ---------------------------------------------------------------- host <-> device matrices
void colmajor_to_rowmajor(pcaone_ctx* c, const double* src, uint64_t rows, int cols, double* dst) {
  dim3 blk(32, 8);
  k_colmajor_to_rowmajor<<<ceil_div(rows, 32), blk, 0, c->stream>>>(src, rows, cols, dst, c->lp);
  PCA_CHECK_LAUNCH();
}
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
  if (c->shard_samples && c->cfg.world > 1) {
    // every rank holds all of G (the same bits): each factors its 1/world of the rows, the l x l Gram
    // matrices are summed, and all ranks end up with the same T
    const uint64_t base = c->M / c->cfg.world, rem = c->M % c->cfg.world, r = (uint64_t)c->cfg.rank;
    const uint64_t m0 = r * base + std::min(r, rem), m1 = m0 + base + (r < rem ? 1 : 0);
    if (!orth_fused_ok(c))
      throw std::runtime_error("sample-sharded jobs need the fused orthonormalisation (k + oversamples <= 80)");
    orth2(c, c->d_G + m0 * c->lp, m1 - m0, nullptr, c->d_T, true, true);
    c->g_is_q = false;
  } else {
    c->g_is_q = orth2(c, c->d_G, c->M, c->d_G, c->d_T, true, true);
  }
  mark();
  ts_rightmult(c, c->d_H, c->l, c->d_T, c->l, c->N, c->d_Bt);
  // SVD of B^T (N x l): Gram -> Cholesky -> one-sided Jacobi on the triangular factor
  ts_gemm_tn(c, c->d_Bt, c->l, c->d_Bt, c->l, c->N, c->d_W, c->shard_samples);
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
  if (c->cfg.world > 1 && !c->shard_samples) {  // (a sample-sharded job holds all rows of V on every rank)
    double* slots = c->d_part;  // the partials above are consumed
    k_flip_slot_write<<<1, 256, 0, c->stream>>>(c->d_scal + 8, c->d_sign, c->k, c->cfg.rank, c->cfg.world, slots);
    PCA_CHECK_LAUNCH();
    comm_allreduce_f64(c, slots, (uint64_t)c->cfg.world * 2 * c->k);
    k_flip_slot_pick<<<1, 128, 0, c->stream>>>(slots, c->k, c->cfg.world, c->d_sign);
    PCA_CHECK_LAUNCH();
    c->tm.kernel_launches += 2;
  }
  k_scale_cols<<<grid_for(c->M * c->k, 256, c->sms), 256, 0, c->stream>>>(c->d_V, c->lp, c->k, c->M, c->d_sign);
  k_scale_cols<<<grid_for(c->N * c->k, 256, c->sms), 256, 0, c->stream>>>(c->d_U, c->lp, c->k, c->N, c->d_sign);
  PCA_CHECK_LAUNCH();
  c->tm.kernel_launches += 4;
}

}  // namespace pcaone
