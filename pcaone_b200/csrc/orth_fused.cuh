// pcaone_b200 — the whole tall-skinny orthonormalisation in ONE cooperative kernel.
//
//   Omega = thinQ(H); flipOmg                (reference Halko.cpp:120-124, 208-213; RSVD.hpp:80-89)
//   G = Q R twice, R = R2 R1                 (reference Halko.cpp:55-65)
//
// CholeskyQR2 with the Householder sign convention of the reference's thin Q (see
// k_householder_signs) and flipOmg, as phases of a persistent grid separated by grid.sync():
//   P1  partial Gram A^T A per CTA                        P2  distributed reduction -> W
//   P3  CTA 0: T1 = chol(W)^-1 (SVQB eigen route if W is numerically rank deficient)
//   P4  partial Gram of Q1 = A T1 (rows recomputed, never stored)     P5  reduction
//   P6  CTA 0: T2, Ttot = T1 T2, Householder signs from the top l x l block of Q
//   P7  Q = (A T1) T2 o signs written out; partial flipOmg column sums
//   P8  flip decision, Omega *= flip sign, Omega2 = Omega
// Shortcuts taken from the data (each bounded below 1e-11, see the code): T2 from its first-order
// expansion when Q1^T Q1 is within 1e-11 of I; P4/P5 skipped (T2 = I) for factors-only calls whose
// first pass shows cond_F(A)^2 <= 1e5; for Omega updates the sign replay of P6 runs on CTA 0
// WHILE the other CTAs do P7 on the unsigned Q, and P8 applies hsign * flip.
// winSVD calls this up to 63 times per epoch on an N x l matrix that is a few MB: the multi-kernel
// version (2 Gram + 2 reduce + 2 Cholesky + 2 host status reads + 2 right-multiplies + signs + 3
// flip kernels) is launch/latency bound at ~0.3 ms per update; this kernel is one launch and no
// host round trip. Every reduction has a fixed summation order (deterministic).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "small_dense.cuh"

namespace pcaone {
namespace cg = cooperative_groups;

constexpr int kOrthThreads = 256;
// rows per staged tile. 64-row tiles at R = 5 fit since the sign replay's l x l block moved to global scratch, but
// measured SLOWER on the B200 (P1 318 vs 288 us, P4 527 vs 495 us at 500k x 80): 32 rows keep more tiles in flight per SM
__host__ __device__ constexpr int orth_tile_rows(int R) { return R <= 4 ? 64 : 32; }

struct OrthArgs {
  const double* A;   // [rows][lp] input (H or G); may alias Q
  double* Q;         // [rows][lp] output; nullptr = factors only (Ttot with Q = A Ttot): phases P7/P8 are skipped
  double* Q2;        // Omega2 for flipOmg (read, then overwritten with the new Omega) or nullptr
  uint64_t rows;
  int l, lp;
  int want_signs;    // Householder column signs (Omega updates)
  int want_flip;     // flipOmg (needs Q2)
  double* part;      // [gridDim.x][l*lp] partial Gram / flip sums
  double* Wg;        // [l*lp] reduced Gram
  double* T1g;       // [l*lp]
  double* T2g;       // [l*lp]
  double* Ttot;      // [l*lp] out: T1 T2 (Q = A Ttot), or nullptr
  double* hsign;     // [l] out: Householder signs (1.0 when !want_signs)
  double* fsign;     // [l] out: total applied column sign
  double* jscratch;  // [2*l*l + 2*l] global scratch of the eigen fallback
  int* status;       // out: number of factorizations that took the eigen (rank-deficient) route
  unsigned long long* prof;  // optional: globaltimer (ns) at the phase boundaries, CTA 0 (debug aid)
  // Which part of the kernel this launch runs (bit mask, 0 = 7 = everything). Row-SHARDED inputs
  // (multi-GPU G) need the two l x l Gram matrices summed across ranks between the parts, so the
  // host runs three launches with an allreduce of Wg after each of the first two:
  //   1: P1-P2 (Gram of A -> Wg)   2: P3-P5 (T1 from Wg, Gram of A T1 -> Wg)   4: P6-P8 (T2, Ttot, Q)
  int phases;
  // optional: per-column max |Q| as IEEE bit patterns (atomicMax) — what the int8 route's slicing of
  // the new Omega needs, saving its own column-max kernel. Single-launch calls clear
  // colmax_out[0, 2 lp) themselves (the maxima and the slice kernel's column sums behind them):
  // CTA 0 in P1, barriers before the atomics of P7
  unsigned long long* colmax_out;
  // optional (factors-only calls): device flag through which CTA 0 tells the grid
  // that one Cholesky pass is enough (see P3); nullptr = always CholeskyQR2
  int* skip2;
  double skip_diag;  // phase-split launches: this rank's share of the identity (1 on rank 0, else 0)
  // The same decision for Omega updates on the int8 route (Q != nullptr): there Omega is rounded to 8S-1 = 23 bits
  // per column scale right after this kernel, which moves every column by ~3e-7 of its norm, so an orthonormality
  // defect below that cannot be seen: the second pass is skipped while eps * cond_F(H)^2 <= 3e-7 (skip_thresh holds
  // the bound on cond_F^2; 1e5 for the factors-only calls above). Because eps * cond^2 is what the defect IS in
  // practice, not a rigorous bound, every 16th update runs both passes anyway (force_full) and MEASURES
  // |Q1^T Q1 - I|: above veto_tol it sets *veto and no later update skips.
  double skip_thresh;
  int force_full;
  int* veto;
  double veto_tol;
  // Row-sharded Omega update (rows of A, Q, Q2 are this rank's samples): the launch with phases = 4
  // writes the UNSIGNED Q, leaves its flipOmg column sums (both signs) and — on the rank that owns
  // the top l rows (want_signs) — the Householder signs in flipbuf[3 l] = {dsum, ssum, hsign} and
  // stops; the host sums flipbuf over the ranks and a launch with phases = 8 applies hsign * flip.
  double* flipbuf;
  // int8 route: write Q = A (T1 T2) with ONE tile product instead of (A T1) T2. The two forms differ by
  // ~eps * cond(A) in the orthogonality of Q (1e-13 here); the int8 route rounds Omega to 8S-1 = 23 bits
  // right after, so the difference cannot be seen, and a third of the sweep's flops go away.
  int one_shot;
  // Row-sharded update as ONE launch: the three small exchanges (two l x l Gram matrices; flipOmg sums + signs +
  // column maxima) run INSIDE the kernel over peer memory instead of through three NCCL launches between four
  // kernel launches. Every rank owns a mailbox [kPeerSlots][world][l * lp doubles] plus flags
  // [kPeerSlots][world] in its HBM, mapped into every peer (CUDA IPC / peer access). CTA 0 PUSHES its contribution
  // into slot (seq % kPeerSlots), row `rank`, of every rank's mailbox (NVLink stores), fences, raises flag
  // [slot][rank] = seq on every rank, waits until its own flags of that slot all show seq, and sums its local copies
  // in rank order — the same order on every rank, so all ranks hold the same bits. A slot is reused kPeerSlots
  // exchanges later; passing an exchange implies every peer has consumed the previous one.
  int peer_world, peer_rank;
  double* const* peer_mbox;                 // [world] device pointers (this rank's view) to the mailboxes
  unsigned long long* const* peer_flag;     // [world] device pointers to the flag arrays
  unsigned long long peer_seq;              // sequence number of the launch's first exchange (monotonic, > 0)
};

constexpr int kPeerSlots = 4;
constexpr int kPeerMaxWorld = 16;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// In-kernel allreduce of buf[0, n) over the ranks of the job, in two halves:
//   peer_push      any thread that owns element i of this rank's contribution stores it into slot row `rank` of
//                  EVERY rank's mailbox (NVLink stores; the grid-wide Gram reduction pushes from all SMs at once),
//                  followed by __threadfence_system() and a grid / block barrier
//   peer_complete  ONE CTA raises this rank's flag on every rank, waits for the world's flags in its own mailbox
//                  and sums its local copies in rank order (loads batched 4 elements x world at a time).
// Elements [0, nsum) are summed as doubles, [nsum, n) are combined with max on their bit patterns
// (non-negative doubles). A peer that never shows up ends the wait after ~2 s with *status |= 0x40000000.
__device__ __forceinline__ size_t peer_row_off(const OrthArgs& a, unsigned long long seq, int stride) {
  return ((size_t)(seq % kPeerSlots) * a.peer_world + a.peer_rank) * stride;
}
__device__ inline void peer_complete(const OrthArgs& a, double* buf, int n, int nsum, unsigned long long seq, int stride,
                                     int* status) {
  const int tid = threadIdx.x, nt = blockDim.x, W = a.peer_world, me = a.peer_rank;
  const int slot = (int)(seq % kPeerSlots);
  if (tid < W) st_release_sys(a.peer_flag[tid] + slot * W + me, seq);
  if (tid < W) {
    const unsigned long long* f = a.peer_flag[me] + slot * W + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > 4000000000ll) {
        atomicOr(status, 0x40000000);
        break;
      }
    }
  }
  __syncthreads();
  const double* mine = a.peer_mbox[me] + (size_t)slot * W * stride;
  constexpr int EB = 4;
  for (int i0 = tid; i0 < n; i0 += nt * EB) {
    double v[EB][kPeerMaxWorld];
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int i = i0 + nt * e;
#pragma unroll
      for (int src = 0; src < kPeerMaxWorld; ++src)
        v[e][src] = (i < n && src < W) ? __ldcg(mine + (size_t)src * stride + i) : 0.0;
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int i = i0 + nt * e;
      if (i >= n) continue;
      if (i < nsum) {
        double acc = 0.0;
#pragma unroll
        for (int src = 0; src < kPeerMaxWorld; ++src)
          if (src < W) acc += v[e][src];
        buf[i] = acc;
      } else {
        unsigned long long m = 0ull;
#pragma unroll
        for (int src = 0; src < kPeerMaxWorld; ++src) {
          const unsigned long long x = (unsigned long long)__double_as_longlong(v[e][src]);
          if (src < W) m = x > m ? x : m;
        }
        buf[i] = __longlong_as_double((long long)m);
      }
    }
  }
  __threadfence();
  __syncthreads();
}
// one CTA pushes a small buffer itself, then completes
__device__ inline void peer_allreduce_small(const OrthArgs& a, double* buf, int n, int nsum, unsigned long long seq,
                                            int stride, int* status) {
  const int tid = threadIdx.x, nt = blockDim.x, W = a.peer_world;
  const size_t off = peer_row_off(a, seq, stride);
  for (int idx = tid; idx < n * W; idx += nt) {
    const int dst = idx / n, i = idx - dst * n;
    a.peer_mbox[dst][off + i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  peer_complete(a, buf, n, nsum, seq, stride, status);
}

// shared-memory strides: rows of the staged tiles and of the factor matrices are LC + 4 doubles
// (= 4 mod 16), which makes every DMMA fragment load below bank-conflict free
__host__ __device__ inline size_t orth_smem_bytes(int l, int R) {
  const int LD = 16 * R + 4;
  return ((size_t)2 * l * LD + (size_t)2 * orth_tile_rows(R) * LD) * sizeof(double);
}

// W (l x l, row-major ld, global) -> T (l x l row-major ld, global) with (A T) orthonormal:
// T = R^-1 from the Cholesky factor, or V diag(lam^-1/2) when the pivot test fails.
// Executed by ONE CTA; `Ws` is l*l doubles of shared memory.
// The result is left in shared memory Ts ([l][lc], zero padded) AND written to global T.
template <int R>
__device__ inline void orth_factor(const double* __restrict__ W, int l, int ld, double* __restrict__ T,
                                   double* __restrict__ Ws, double* __restrict__ Ts, int lc,
                                   double* __restrict__ jscratch, int* status) {
  // Right-looking Cholesky W = R^T R with the matrix held in REGISTERS: thread (ty, tx) of the
  // 16 x 16 CTA owns the elements (ty + 16 i, tx + 16 j). Per column: the owners of row j publish
  // it to a double-buffered shared row, ONE __syncthreads, then every thread updates its own
  // elements. (The first version kept W in shared memory: 3 barriers and two integer divisions
  // per column made this single-CTA phase 47 us of a 190 us Omega update.)
  // The inverse comes out of the SAME loop: with L = R^T, forward substitution L Y = I in its
  // right-looking form is the same rank-1 update applied to a second register tile B (= I on
  // entry): Y[j][:] = B[j][:] / L[j][j], B[r][:] -= L[r][j] Y[j][:] for r > j, and T = R^-1 = Y^T.
  // (A separate back substitution, one thread per column, was a serial chain of ~l^2/2 dependent
  // shared-memory FMAs: half of this phase.)
  __shared__ double s_row[2][16 * R], s_brow[2][16 * R], s_piv[16 * R];
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4;
  double a[R][R], bi[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      a[i][j] = (r < l && c < l) ? W[r * ld + c] : 0.0;
      bi[i][j] = (r == c && r < l) ? 1.0 : 0.0;
      if (r == c && r < l) s_row[0][r] = a[i][j];
    }
  if (tid == 0) s_fail = 0;
  for (int i = tid; i < l * lc; i += nt) Ts[i] = 0.0;
  __syncthreads();
  // Second pass of CholeskyQR2: W = Q1^T Q1 = I + E with |E| ~ eps * cond(A)^2. For |E_ij| <= 1e-11
  // the factor is its own first-order expansion to below one ulp (the dropped terms are
  // <= (l * 1e-11)^2 < 1e-18): R = I + U, U = strict_upper(E) + diag(E) / 2, T = R^-1 = I - U.
  // That replaces the l-step serial loop below (24 us of a 130 us Omega update at l = 40) by one
  // parallel pass; anything larger (first pass, ill-conditioned inputs, NaN) takes the loop.
  {
    bool big = false;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        if (r < l && c < l) big |= !(fabs(a[i][j] - (r == c ? 1.0 : 0.0)) <= 1e-11);
      }
    if (!__syncthreads_or(big)) {
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int r = ty + 16 * i, c = tx + 16 * j;
          if (r < l && c < l && r <= c) Ts[r * lc + c] = (r == c) ? 1.0 - 0.5 * (a[i][j] - 1.0) : -a[i][j];
        }
      __syncthreads();
      for (int i = tid; i < l * ld; i += nt) {
        const int r = i / ld, c = i - r * ld;
        T[i] = c < l ? Ts[r * lc + c] : 0.0;
      }
      __syncthreads();
      return;
    }
  }
  double maxd = 0.0;
  for (int i = 0; i < l; ++i) maxd = fmax(maxd, s_row[0][i]);
  __syncthreads();
  const double tol = 64.0 * l * 2.220446049250313e-16 * maxd;
  bool fail = false;
  for (int j = 0; j < l; ++j) {
    const int b = j & 1, jt = j >> 4;
    if (ty == (j & 15)) {
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (i == jt) {
#pragma unroll
          for (int jj = 0; jj < R; ++jj) {
            s_row[b][tx + 16 * jj] = a[i][jj];
            s_brow[b][tx + 16 * jj] = bi[i][jj];
          }
        }
    }
    __syncthreads();
    const double d = s_row[b][j];
    if (!(d > tol)) {  // uniform: every thread reads the same pivot
      fail = true;
      break;
    }
    // 1/d is on the serial path of every step: __drcp_rn is the same correctly rounded reciprocal
    // as 1.0 / d without the generic division sequence; the 1/sqrt(d) scaling of column j of T is
    // deferred to one parallel pass after the loop (s_piv)
    const double invd = __drcp_rn(d);
    if (tid <= j) Ts[tid * lc + j] = s_brow[b][tid];  // column j of T = row j of Y, still times sqrt(d_j)
    if (tid == 0) s_piv[j] = d;
    double rr[R], rc[R], bc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) rr[i] = s_row[b][ty + 16 * i] * invd;
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      rc[jj] = s_row[b][tx + 16 * jj];
      bc[jj] = s_brow[b][tx + 16 * jj];
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int jj = 0; jj < R; ++jj) {
        const int r = ty + 16 * i, c = tx + 16 * jj;
        if (r > j && c >= r) a[i][jj] -= rr[i] * rc[jj];
        if (r > j && c <= j) bi[i][jj] -= rr[i] * bc[jj];
      }
  }
  if (fail && tid == 0) s_fail = 1;
  __syncthreads();
  if (!s_fail) {
    if (tid < l) s_piv[tid] = 1.0 / sqrt(s_piv[tid]);
    __syncthreads();
    for (int i = tid; i < l * lc; i += nt) {
      const int c = i % lc;
      if (c < l) Ts[i] *= s_piv[c];
    }
    __syncthreads();
    for (int i = tid; i < l * ld; i += nt) {
      const int r = i / ld, c = i - r * ld;
      T[i] = c < l ? Ts[r * lc + c] : 0.0;
    }
    __syncthreads();
    return;
  }
  for (int i = tid; i < l * lc; i += nt) Ts[i] = 0.0;
  __syncthreads();
  // ---- eigen route (SVQB): one-sided Jacobi on W in global scratch (rare, slow path)
  double* Aj = jscratch;            // column-major l x l
  double* Vj = Aj + (size_t)l * l;  // column-major l x l
  double* nrm = Vj + (size_t)l * l; // l
  __shared__ int s_rot;
  __shared__ int s_ord[kMaxL];
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int i = tid; i < l * l; i += nt) {
    const int r = i / l, c = i % l;
    Aj[c * l + r] = W[r * ld + c];
    Vj[c * l + r] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int n = (l + 1) & ~1;
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int round = 0; round < n - 1; ++round) {
      for (int pi = warp; pi < n / 2; pi += nw) {
        int p, q;
        if (pi == 0) {
          p = n - 1;
          q = round;
        } else {
          p = (round + pi) % (n - 1);
          q = (round - pi + (n - 1)) % (n - 1);
        }
        if (p > q) {
          const int tmp = p;
          p = q;
          q = tmp;
        }
        if (q >= l) continue;
        double* ap = Aj + p * l;
        double* aq = Aj + q * l;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int r = lane; r < l; r += 32) {
          const double x = ap[r], y = aq[r];
          alpha += x * x;
          beta += y * y;
          gamma += x * y;
        }
        alpha = warp_sum(alpha);
        beta = warp_sum(beta);
        gamma = warp_sum(gamma);
        if (fabs(gamma) > 1e-15 * sqrt(alpha * beta) && gamma != 0.0) {
          if (lane == 0) s_rot = 1;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cc = 1.0 / sqrt(1.0 + tt * tt), ss = cc * tt;
          double* vp = Vj + p * l;
          double* vq = Vj + q * l;
          for (int r = lane; r < l; r += 32) {
            const double x = ap[r], y = aq[r];
            ap[r] = cc * x - ss * y;
            aq[r] = ss * x + cc * y;
            const double vx = vp[r], vy = vq[r];
            vp[r] = cc * vx - ss * vy;
            vq[r] = ss * vx + cc * vy;
          }
        }
        __syncwarp();
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (!rot) break;
  }
  for (int j = warp; j < l; j += nw) {
    double s = 0.0;
    for (int r = lane; r < l; r += 32) s += Aj[j * l + r] * Aj[j * l + r];
    s = warp_sum(s);
    if (lane == 0) nrm[j] = sqrt(s);  // eigenvalue lam_j of W
  }
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < l; ++j) s_ord[j] = j;
    for (int a = 1; a < l; ++a) {
      const int key = s_ord[a];
      int b = a - 1;
      while (b >= 0 && nrm[s_ord[b]] < nrm[key]) {
        s_ord[b + 1] = s_ord[b];
        --b;
      }
      s_ord[b + 1] = key;
    }
    atomicAdd(status, 1);
  }
  __syncthreads();
  const double smax = sqrt(nrm[s_ord[0]]);
  for (int i = tid; i < l * l; i += nt) {
    const int r = i / l, c = i % l;
    const double s = sqrt(nrm[s_ord[c]]);
    Ts[r * lc + c] = (s > 1e-7 * smax && s > 0.0) ? Vj[s_ord[c] * l + r] / s : 0.0;
  }
  __syncthreads();
  for (int i = tid; i < l * ld; i += nt) {
    const int r = i / ld, c = i - r * ld;
    T[i] = c < l ? Ts[r * lc + c] : 0.0;
  }
  __syncthreads();
}

// out[e] = sum_p part[p*stride + e], one warp per element, lanes stride over parts, fixed order
// (push_to: optional peer mailboxes — lane d also stores the element into rank d's mailbox at row offset push_off)
__device__ __forceinline__ void orth_reduce_parts(const double* __restrict__ part, int nparts, size_t stride,
                                                  int nelem, double* __restrict__ out, int gwarp, int nwarps, int lane,
                                                  double* const* push_to = nullptr, int push_world = 0,
                                                  size_t push_off = 0) {
  for (int e = gwarp; e < nelem; e += nwarps) {
    double v = 0.0;
    for (int p = lane; p < nparts; p += 32) v += part[(size_t)p * stride + e];
    v = warp_sum(v);
    if (lane == 0) out[e] = v;
    if (push_to && lane < push_world) push_to[lane][push_off + e] = v;
  }
  if (push_to) __threadfence_system();
}

template <int R>
__global__ void __launch_bounds__(kOrthThreads, 1) k_orth_fused(const OrthArgs a) {
  cg::grid_group grid = cg::this_grid();
  constexpr int LC = 16 * R;
  constexpr int TR = orth_tile_rows(R);  // rows per tile
  constexpr int RI = TR / 16;            // rows per thread in the 16 x 16 epilogue layout
  constexpr int NLD = TR * R / 16;       // register-prefetched doubles per thread (>= TR*lp/256)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = LC + 4;                          // smem row stride of tiles and factors
  constexpr int NT = LC / 8;                          // 8-column MMA tiles per row
  constexpr int MT = TR / 8;                          // 8-row MMA tiles per staged tile (8 or 4)
  constexpr int NSPLIT = 8 / MT;                      // warps sharing one row tile (1 or 2)
  constexpr int NPW = NT / NSPLIT;                    // column tiles per warp in a tile x factor product
  constexpr int NTT = NT * (NT + 1) / 2;              // upper-triangle tiles of the Gram
  constexpr int GPW = (NTT + 7) / 8;                  // Gram tiles per warp
  double* T1s = reinterpret_cast<double*>(smem_raw);  // [l][LD]
  double* T2s = T1s + (size_t)a.l * LD;               // [l][LD]
  double* As = T2s + (size_t)a.l * LD;                // [TR][LD]
  double* Qs = As + TR * LD;                          // [TR][LD]
  double* Ws = a.jscratch;                            // [l][l] in GLOBAL scratch (CTA 0 only, the sign replay's top block):
                                                      // keeping it out of shared memory lets R = 5 stage 64-row tiles too
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int fg = lane >> 2, ft = lane & 3;            // DMMA fragment coordinates (common.cuh dmma884)
  const int m0 = (warp % MT) * 8;                     // this warp's row tile in tile x factor products
  const int nbase = (warp / MT) * NPW;                // ... and its first column tile
  const int l = a.l, lp = a.lp;
  const int gwarp = (blockIdx.x * kOrthThreads + tid) >> 5, nwarps = (gridDim.x * kOrthThreads) >> 5;
  const uint64_t rpc = ((a.rows + gridDim.x - 1) / gridDim.x + TR - 1) / TR * TR;
  uint64_t r0 = min(a.rows, (uint64_t)blockIdx.x * rpc), r1 = min(a.rows, r0 + rpc);  // (re-split for P7/P8 below)
  const size_t pstride = (size_t)l * lp;
  double* mypart = a.part + (size_t)blockIdx.x * pstride;
  const int tile_elems = TR * lp;  // a tile is TR contiguous rows of lp doubles

  // register prefetch of the tile starting at row r (flat, coalesced), zero beyond r1
  double pre[NLD];
  auto prefetch = [&](uint64_t r) {
    const double* src = a.A + r * lp;
    const uint64_t lim = r < r1 ? (r1 - r) * (uint64_t)lp : 0;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int e = tid + kOrthThreads * i;
      pre[i] = (e < tile_elems && (uint64_t)e < lim) ? src[e] : 0.0;
    }
  };
  auto commit_tile = [&]() {  // registers -> As[rr][c]
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int e = tid + kOrthThreads * i;
      if (e < tile_elems) {
        const int rr = e / lp, c = e - rr * lp;
        if (c < LC) As[rr * LD + c] = c < l ? pre[i] : 0.0;
      }
    }
  };
  auto zero_pad_cols = [&]() {  // columns lp..LC-1 of As are never written by commit_tile
    if (lp < LC)
      for (int idx = tid; idx < TR * (LC - lp); idx += kOrthThreads) {
        const int rr = idx / (LC - lp), c = lp + idx - rr * (LC - lp);
        As[rr * LD + c] = 0.0;
      }
  };
  auto load_T = [&](const double* Tg, double* Ts) {
    for (int idx = tid; idx < l * LC; idx += kOrthThreads) {
      const int r = idx / LC, c = idx - r * LC;
      Ts[r * LD + c] = c < l ? Tg[r * lp + c] : 0.0;
    }
  };
  // The three tall products run on the FP64 tensor cores (DMMA m8n8k4). The scalar version (each
  // thread an RI x R register tile fed by 7 shared-memory loads per 12 FMAs) reached ~45 % of the FP64
  // rate and made QR(G) on 1M rows 1.2 ms of a 1.85 ms dense stage.
  //
  // tile x factor: acc[n][0..1] = (src[TR][0..l) * Ts[0..l)[LC]) at rows m0 + fg, columns 8 (nbase + n) + 2 ft (+1)
  auto tile_times_T = [&](const double* src, const double* Ts, double (&acc)[NPW][2]) {
#pragma unroll
    for (int n = 0; n < NPW; ++n) acc[n][0] = acc[n][1] = 0.0;
    const double* ap = src + (m0 + fg) * LD + ft;
    const double* bp = Ts + ft * LD + 8 * nbase + fg;
    const int l4 = (l + 3) & ~3;
#pragma unroll 2
    for (int k = 0; k < l4; k += 4) {
      const double av = ap[k];                       // columns >= l of the staged tile are zero
      const bool kin = k + ft < l;                   // rows >= l of Ts do not exist
#pragma unroll
      for (int n = 0; n < NPW; ++n) {
        const double bv = kin ? bp[k * LD + 8 * n] : 0.0;
        dmma884(acc[n][0], acc[n][1], av, bv);
      }
    }
  };
  auto store_tile = [&](double* dst, const double (&acc)[NPW][2]) {
#pragma unroll
    for (int n = 0; n < NPW; ++n)
      *reinterpret_cast<double2*>(dst + (m0 + fg) * LD + 8 * (nbase + n) + 2 * ft) = make_double2(acc[n][0], acc[n][1]);
  };
  // Gram: upper-triangle 8 x 8 tiles (ti <= tj) of src^T src, tile q of this warp = list entry warp + 8 q
  int g_ti[GPW], g_tj[GPW];
#pragma unroll
  for (int q = 0; q < GPW; ++q) {
    int idx = warp + 8 * q, ti = 0;
    if (idx >= NTT) {
      g_ti[q] = g_tj[q] = -1;
    } else {
      while (idx >= NT - ti) {
        idx -= NT - ti;
        ++ti;
      }
      g_ti[q] = ti;
      g_tj[q] = ti + idx;
    }
  }
  auto gram_accumulate = [&](const double* src, double (&acc)[GPW][2]) {
#pragma unroll 4
    for (int rr = 0; rr < TR; rr += 4) {
      const double* row = src + (rr + ft) * LD + fg;  // A[i = fg][k = ft] = src[k][8 ti + fg], B[k = ft][j = fg] = src[k][8 tj + fg]
#pragma unroll
      for (int q = 0; q < GPW; ++q)
        if (g_ti[q] >= 0) dmma884(acc[q][0], acc[q][1], row[8 * g_ti[q]], row[8 * g_tj[q]]);
    }
  };
  auto store_part = [&](double (&acc)[GPW][2]) {
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
      if (g_ti[q] < 0) continue;
      const int r = 8 * g_ti[q] + fg;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 8 * g_tj[q] + 2 * ft + h;
        const double v = (r < l && c < l) ? acc[q][h] : 0.0;
        if (r < l && c < lp) mypart[r * lp + c] = v;
        if (g_ti[q] != g_tj[q] && c < l && r < lp) mypart[c * lp + r] = v;  // mirror
      }
    }
  };
  // pad entries of the partial Gram that no tile writes (rows < l, columns in [LC, lp) never exist: lp <= LC)

  int prof_i = 0;
  auto stamp = [&]() {
    if (a.prof && blockIdx.x == 0 && tid == 0 && prof_i < 62) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.prof[prof_i] = t;
    }
    ++prof_i;
  };
  // P8 body: Qs[0, 2l) holds the flipOmg column sums of the (possibly unsigned) Q; `unsigned_q`: they
  // were taken before the Householder signs hs[] were applied (with hs = -1 they swap roles)
  auto apply_flip = [&](bool unsigned_q, const double* hs_src) {
    for (int c = tid; c < l; c += kOrthThreads) {
      const double hsg = unsigned_q ? __ldcg(hs_src + c) : 1.0;
      const double dsm = hsg < 0.0 ? Qs[l + c] : Qs[c], ssm = hsg < 0.0 ? Qs[c] : Qs[l + c];
      const double f = (dsm > 2 * ssm) ? -1.0 : 1.0;
      As[c] = f * hsg;
      if (blockIdx.x == 0) a.fsign[c] = unsigned_q ? f * hsg : f * __ldcg(a.hsign + c);
    }
    __syncthreads();
    const uint64_t total = (r1 - r0) * (uint64_t)lp;
    constexpr int PB = 8;  // loads of a batch are issued before the first store (independent round trips)
    for (uint64_t i0 = tid; i0 < total; i0 += (uint64_t)kOrthThreads * PB) {
      double v[PB];
#pragma unroll
      for (int b = 0; b < PB; ++b) {
        const uint64_t i = i0 + (uint64_t)kOrthThreads * b;
        v[b] = i < total ? a.Q[r0 * lp + i] : 0.0;
      }
#pragma unroll
      for (int b = 0; b < PB; ++b) {
        const uint64_t i = i0 + (uint64_t)kOrthThreads * b;
        if (i < total) {
          const int c = (int)(i % lp);
          const double w = c < l ? v[b] * As[c] : v[b];
          a.Q[r0 * lp + i] = w;
          if (a.Q2) a.Q2[r0 * lp + i] = w;
        }
      }
    }
  };
  const int ph = a.phases ? a.phases : 7;
  const bool peer = a.peer_world > 1 && ph == 7;  // single-launch row-sharded update (exchanges over peer memory)
  if (ph == 8) {  // row-sharded Omega update, last launch: flipbuf = {dsum, ssum, hsign} summed over the ranks
    for (int c = tid; c < 2 * l; c += kOrthThreads) Qs[c] = a.flipbuf[c];
    __syncthreads();
    apply_flip(true, a.flipbuf + 2 * l);
    return;
  }
  stamp();
  if (a.colmax_out && ph == 7 && blockIdx.x == 0)
    for (int i = tid; i < 2 * lp; i += kOrthThreads) a.colmax_out[i] = 0ull;
  zero_pad_cols();
  // ---------------- P1: partial Gram of A
  if (ph & 1) {
    double acc[GPW][2];
#pragma unroll
    for (int q = 0; q < GPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    prefetch(r0);
    for (uint64_t r = r0; r < r1; r += TR) {
      __syncthreads();
      commit_tile();
      prefetch(r + TR);
      __syncthreads();
      gram_accumulate(As, acc);
    }
    store_part(acc);
    stamp();
    grid.sync();
    stamp();
    // ---------------- P2: W = sum of partials
    orth_reduce_parts(a.part, gridDim.x, pstride, l * lp, a.Wg, gwarp, nwarps, lane, peer ? a.peer_mbox : nullptr,
                      a.peer_world, peer ? peer_row_off(a, a.peer_seq, l * lp) : 0);
    stamp();
    if (ph & 6) grid.sync();
    stamp();
  }
  // ---------------- P3: T1
  if (ph & 2) {
    if (blockIdx.x == 0) {
      if (peer) peer_complete(a, a.Wg, l * lp, l * lp, a.peer_seq, l * lp, a.status);
      orth_factor<R>(a.Wg, l, lp, a.T1g, Ws, T1s, LD, a.jscratch, a.status);
      if (a.skip2) {
        // One Cholesky pass leaves |Q1^T Q1 - I| ~ eps * cond_2(A)^2; cond_F(A)^2 = trace(W) * |T1|_F^2
        // bounds cond_2(A)^2 from above (by up to a factor l^2; ~l for a flat spectrum: 1.4e4 against
        // 35 on the bench matrix) and costs one block reduction. Up to 1e5 the second pass (a full
        // re-read of A and two thirds of the flops) would change the factor by < 1e-11 in the worst
        // case, ~1e-13 typically — five orders under the 1e-6 the eigenvalues are held to: skip it, T2 = I.
        __shared__ double s_red[2][kOrthThreads / 32];
        double tr = 0.0, tf = 0.0;
        for (int i = tid; i < l; i += kOrthThreads) tr += a.Wg[i * lp + i];
        for (int i = tid; i < l * LD; i += kOrthThreads) tf += T1s[i] * T1s[i];
        tr = warp_sum(tr);
        tf = warp_sum(tf);
        if (lane == 0) {
          s_red[0][warp] = tr;
          s_red[1][warp] = tf;
        }
        __syncthreads();
        if (tid == 0) {
          tr = tf = 0.0;
          for (int w = 0; w < kOrthThreads / 32; ++w) {
            tr += s_red[0][w];
            tf += s_red[1][w];
          }
          *a.skip2 = (!a.force_full && !(a.veto && *a.veto) && tr * tf <= a.skip_thresh) ? 1 : 0;  // NaN -> 0
        }
      }
    }
    __threadfence();
    stamp();
    grid.sync();
    stamp();
  }
  const bool skip = a.skip2 != nullptr && (ph & 2) && __ldcg(a.skip2) != 0;  // uniform across the grid
  if (skip && !(ph & 4)) {
    // phase-split launch (rows sharded over ranks; every rank took the same decision from the same
    // allreduced Gram): hand the host's second allreduce matrices that sum to I, so that the last
    // launch finds T2 = I
    if (blockIdx.x == 0)
      for (int i = tid; i < l * lp; i += kOrthThreads) a.Wg[i] = (i / lp == i % lp) ? a.skip_diag : 0.0;
    return;
  }
  // ---------------- P4: partial Gram of Q1 = A T1
  if ((ph & 2) ? blockIdx.x != 0 : (ph & 4) != 0) {  // T1 is in CTA 0's shared memory only if P3 ran in this launch
    load_T(a.T1g, T1s);
    __syncthreads();
  }
  if ((ph & 2) && !skip) {
    double acc[GPW][2];
#pragma unroll
    for (int q = 0; q < GPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    prefetch(r0);
    for (uint64_t r = r0; r < r1; r += TR) {
      __syncthreads();
      commit_tile();
      prefetch(r + TR);
      __syncthreads();
      double q1[NPW][2];
      tile_times_T(As, T1s, q1);
      store_tile(Qs, q1);
      __syncthreads();
      gram_accumulate(Qs, acc);
    }
    store_part(acc);
    stamp();
    grid.sync();
    stamp();
    // ---------------- P5
    orth_reduce_parts(a.part, gridDim.x, pstride, l * lp, a.Wg, gwarp, nwarps, lane, peer ? a.peer_mbox : nullptr,
                      a.peer_world, peer ? peer_row_off(a, a.peer_seq + 1, l * lp) : 0);
    if (ph & 4) grid.sync();
    stamp();
  }
  if (!(ph & 4)) return;  // uniform across the grid
  // ---------------- P6: T2, Ttot, Householder signs (CTA 0)
  // Omega updates (signs AND flipOmg): the sign replay is a serial l-step loop on one CTA (~20 us at
  // l = 40) that only P8 needs — P7 can write Q unsigned and sum |Omega2 -+ Q| for both signs. So the
  // grid barrier moves up to right after the T2 factor, CTA 0 replays the signs (stage 1) WHILE the
  // other CTAs run P7 on all rows, and P8 applies hsign * flip in its one sweep (same bits as
  // before: |o2 - h q| and |o2 + h q| swap roles when h = -1).
  // (not for in-place calls: CTA 0 still reads the top rows of A while the others write Q)
  const bool overlap = a.want_signs && a.want_flip && a.Q != nullptr && a.Q != a.A && gridDim.x > 1;
  for (int stage = 0; stage < 2; ++stage) {
    if (stage == 0 && blockIdx.x == 0) {
      if (skip) {  // single pass: T2 = I
        for (int i = tid; i < l * LD; i += kOrthThreads) T2s[i] = (i / LD == i % LD) ? 1.0 : 0.0;
        for (int i = tid; i < l * lp; i += kOrthThreads) a.T2g[i] = (i / lp == i % lp) ? 1.0 : 0.0;
        __syncthreads();
      } else {
        if (peer) peer_complete(a, a.Wg, l * lp, l * lp, a.peer_seq + 1, l * lp, a.status);
        if (a.veto) {  // a full update on a route that may skip the second pass: measure the defect it removes
          bool big = false;
          for (int i = tid; i < l * l; i += kOrthThreads) {
            const int r = i / l, c = i - r * l;
            big |= !(fabs(a.Wg[r * lp + c] - (r == c ? 1.0 : 0.0)) <= a.veto_tol);
          }
          if (__syncthreads_or(big) && tid == 0) *a.veto = 1;
        }
        orth_factor<R>(a.Wg, l, lp, a.T2g, Ws, T2s, LD, a.jscratch, a.status);
      }
      if (a.one_shot && a.Q) {  // T2s / T2g <- T1 T2 (both upper triangular) before anybody else reads T2
        double* tmp = a.jscratch + (size_t)l * l;  // [l][l] global scratch (second half; the first holds the sign replay's block)
        for (int idx = tid; idx < l * l; idx += kOrthThreads) {
          const int r = idx / l, c = idx - r * l;
          double acc = 0.0;
          if (r <= c)
            for (int k = r; k <= c; ++k) acc += T1s[r * LD + k] * T2s[k * LD + c];
          tmp[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < l * l; idx += kOrthThreads) {
          const int r = idx / l, c = idx - r * l;
          const double v = tmp[idx];
          T2s[r * LD + c] = v;
          a.T2g[r * lp + c] = v;
        }
        __syncthreads();
      }
      stamp();
    }
    if (blockIdx.x == 0 && stage == (overlap ? 1 : 0)) {
      if (a.Ttot) {
        for (int idx = tid; idx < l * lp; idx += kOrthThreads) {
          const int r = idx / lp, c = idx - r * lp;
          double s = 0.0;
          if (c < l)
            for (int k = 0; k < l; ++k) s += T1s[r * LD + k] * T2s[k * LD + c];
          a.Ttot[idx] = s;
        }
      }
      if (a.want_signs) {
        // top l x l block of Q = (A T1) T2 -> Ws (row-major l x l)
        const uint64_t r1_save = r1;
        for (int rb = 0; rb < l; rb += TR) {
          __syncthreads();
          for (int idx = tid; idx < TR * LC; idx += kOrthThreads) {
            const int rr = idx / LC, c = idx - rr * LC;
            As[rr * LD + c] =
                ((uint64_t)(rb + rr) < a.rows && rb + rr < l && c < l) ? a.A[(uint64_t)(rb + rr) * lp + c] : 0.0;
          }
          __syncthreads();
          double q[NPW][2];
          if (a.one_shot) {
            tile_times_T(As, T2s, q);  // T2s holds T1 T2
          } else {
            tile_times_T(As, T1s, q);
            store_tile(Qs, q);
            __syncthreads();
            tile_times_T(Qs, T2s, q);
          }
#pragma unroll
          for (int n = 0; n < NPW; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int rr = rb + m0 + fg, c = 8 * (nbase + n) + 2 * ft + h;
              if (rr < l && c < l) Ws[rr * l + c] = q[n][h];
            }
        }
        (void)r1_save;
        __syncthreads();
        stamp();
        // sign-modified LU replay of the Householder sign decisions (see k_householder_signs), with
        // the l x l block in registers: per step the owners publish row i and column i, one barrier
        __shared__ double s_lr[2][16 * R], s_lc[2][16 * R];
        double w[R][R];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            w[i][j] = (r < l && c < l) ? Ws[r * l + c] : 0.0;
          }
        for (int i0 = 0; i0 < l; ++i0) {
          const int b = i0 & 1, it = i0 >> 4;
          if (ty == (i0 & 15)) {
#pragma unroll
            for (int i = 0; i < R; ++i)
              if (i == it) {
#pragma unroll
                for (int j = 0; j < R; ++j) s_lr[b][tx + 16 * j] = w[i][j];
              }
          }
          if (tx == (i0 & 15)) {
#pragma unroll
            for (int j = 0; j < R; ++j)
              if (j == it) {
#pragma unroll
                for (int i = 0; i < R; ++i) s_lc[b][ty + 16 * i] = w[i][j];
              }
          }
          __syncthreads();
          const double c0 = s_lr[b][i0];
          const double beta = (c0 >= 0.0) ? -1.0 : 1.0;
          if (tid == 0) a.hsign[i0] = beta;
          const double inv = __drcp_rn(c0 - beta);  // == 1.0 / (c0 - beta), correctly rounded
          double cr[R], rc[R];
#pragma unroll
          for (int i = 0; i < R; ++i) cr[i] = s_lc[b][ty + 16 * i] * inv;
#pragma unroll
          for (int j = 0; j < R; ++j) rc[j] = s_lr[b][tx + 16 * j];
#pragma unroll
          for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const int r = ty + 16 * i, c = tx + 16 * j;
              if (r > i0 && c > i0) w[i][j] -= cr[i] * rc[j];
            }
        }
      } else {
        for (int c = tid; c < l; c += kOrthThreads) a.hsign[c] = 1.0;
      }
    }
    if (stage == 0) {
      __threadfence();
      stamp();
      grid.sync();
      stamp();
      if (!a.Q) {  // factors only (uniform across the grid)
        if (blockIdx.x == 0)
          for (int c = tid; c < l; c += kOrthThreads) a.fsign[c] = a.hsign[c];
        return;
      }
      if (overlap) {  // the rows of P7 / P8 go to CTAs 1.., CTA 0 has none
        const uint64_t rpc7 = ((a.rows + gridDim.x - 2) / (gridDim.x - 1) + TR - 1) / TR * TR;
        r0 = blockIdx.x == 0 ? a.rows : min(a.rows, (uint64_t)(blockIdx.x - 1) * rpc7);
        r1 = min(a.rows, r0 + rpc7);
      }
    }
  }  // stage
  if (overlap && blockIdx.x == 0) __threadfence();  // hsign, Ttot: visible to the grid after the next barrier
  // ---------------- P7: Q = (A T1) T2 o hsign ; partial flip sums
  if (blockIdx.x != 0) load_T(a.T2g, T2s);
  // the epilogue keeps the 16 x 16 thread layout (thread = RI rows x R columns: 3 x R running
  // column statistics per thread instead of 3 x 2 NPW in the MMA fragment layout); the second
  // product goes through shared memory once more (As is free after the first product)
  double hs[R];
#pragma unroll
  for (int j = 0; j < R; ++j) hs[j] = (tx + 16 * j < l) ? ((overlap || a.flipbuf) ? 1.0 : a.hsign[tx + 16 * j]) : 0.0;
  double dsum[R], ssum[R], amax[R];
#pragma unroll
  for (int j = 0; j < R; ++j) dsum[j] = ssum[j] = amax[j] = 0.0;
  zero_pad_cols();
  prefetch(r0);
  stamp();
  for (uint64_t r = r0; r < r1; r += TR) {
    __syncthreads();
    commit_tile();
    prefetch(r + TR);
    __syncthreads();
    stamp();
    // Omega2 of this tile first: the loads fly while the two products run (issued one by one
    // between the stores below they were a chain of dependent L2 round trips, ~8 us per tile)
    double o2[RI][R];
    if (a.want_flip) {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const uint64_t row = r + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int c = tx + 16 * j;
          o2[i][j] = (row < r1 && c < l) ? a.Q2[row * lp + c] : 0.0;
        }
      }
    }
    const double* Res = a.one_shot ? Qs : As;  // where the tile of Q ends up
    {
      double q[NPW][2];
      if (a.one_shot) {
        tile_times_T(As, T2s, q);  // T2s holds T1 T2
        store_tile(Qs, q);
        __syncthreads();
        stamp();
        stamp();
      } else {
        tile_times_T(As, T1s, q);
        store_tile(Qs, q);
        __syncthreads();
        stamp();
        tile_times_T(Qs, T2s, q);
        store_tile(As, q);  // every warp finished reading As before the barrier above
        __syncthreads();
        stamp();
      }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      const uint64_t row = r + ty + 16 * i;
      if (row < r1) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int c = tx + 16 * j;
          if (c < lp) {
            const double qv = c < l ? Res[(ty + 16 * i) * LD + c] * hs[j] : 0.0;
            amax[j] = fmax(amax[j], fabs(qv));
            if (a.want_flip && c < l) {
              dsum[j] += fabs(o2[i][j] - qv);
              ssum[j] += fabs(o2[i][j] + qv);
            }
            a.Q[row * lp + c] = qv;
          }
        }
      }
    }
  }
  if (a.colmax_out) {  // the 16 row-lanes of a column -> one atomicMax per column and CTA
    __syncthreads();
#pragma unroll
    for (int j = 0; j < R; ++j) As[ty * LD + tx + 16 * j] = amax[j];
    __syncthreads();
    for (int c = tid; c < l; c += kOrthThreads) {
      double m = 0.0;
      for (int y = 0; y < 16; ++y) m = fmax(m, As[y * LD + c]);
      if (m > 0.0) atomicMax(a.colmax_out + c, (unsigned long long)__double_as_longlong(m));
    }
  }
  if (!a.want_flip && !a.flipbuf) {
    if (blockIdx.x == 0)
      for (int c = tid; c < l; c += kOrthThreads) a.fsign[c] = a.hsign[c];
    return;  // uniform across the grid: no further grid.sync
  }
  __syncthreads();
  // reduce the 16 row-lanes (ty) per column in fixed order: reuse As/Qs as [16][LD]
#pragma unroll
  for (int j = 0; j < R; ++j) {
    As[ty * LD + tx + 16 * j] = dsum[j];
    Qs[ty * LD + tx + 16 * j] = ssum[j];
  }
  __syncthreads();
  for (int c = tid; c < l; c += kOrthThreads) {
    double d = 0.0, s = 0.0;
    for (int y = 0; y < 16; ++y) {
      d += As[y * LD + c];
      s += Qs[y * LD + c];
    }
    mypart[c] = d;
    mypart[l + c] = s;
  }
  stamp();
  grid.sync();
  stamp();
  if (a.flipbuf) {  // row-sharded: the sums (and rank 0's signs) go through the host's allreduce first
    if (blockIdx.x == 0) {
      for (int c = tid; c < 2 * l; c += kOrthThreads) {
        double v = 0.0;
        for (unsigned p0 = 0; p0 < gridDim.x; ++p0) v += a.part[(size_t)p0 * pstride + c];  // CTA order
        a.flipbuf[c] = v;
      }
      for (int c = tid; c < l; c += kOrthThreads) a.flipbuf[2 * l + c] = a.want_signs ? __ldcg(a.hsign + c) : 0.0;
      if (peer) {
        // {dsum, ssum, hsign} summed, the column maxima of |Q| (bit patterns) max'ed, in one exchange
        for (int c = tid; c < l; c += kOrthThreads)
          a.flipbuf[3 * l + c] = a.colmax_out ? __longlong_as_double((long long)__ldcg(a.colmax_out + c)) : 0.0;
        __syncthreads();
        peer_allreduce_small(a, a.flipbuf, 4 * l, 3 * l, a.peer_seq + 2, l * lp, a.status);
        if (a.colmax_out)
          for (int c = tid; c < l; c += kOrthThreads)
            a.colmax_out[c] = (unsigned long long)__double_as_longlong(a.flipbuf[3 * l + c]);
        __threadfence();
      }
    }
    if (!peer) return;
    grid.sync();
    for (int c = tid; c < 2 * l; c += kOrthThreads) Qs[c] = __ldcg(a.flipbuf + c);
    __syncthreads();
    apply_flip(true, a.flipbuf + 2 * l);
    return;
  }
  // ---------------- P8: flip decision (every CTA, same fixed order), apply to own rows
  __syncthreads();
  for (int c = tid; c < 2 * l; c += kOrthThreads) {
    double v = 0.0;
    for (unsigned p0 = 0; p0 < gridDim.x; p0 += 16) {  // 16 independent loads, then the adds in CTA order
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = (p0 + u < gridDim.x) ? a.part[(size_t)(p0 + u) * pstride + c] : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u) v += t[u];
    }
    Qs[c] = v;
  }
  __syncthreads();
  apply_flip(overlap, a.hsign);
  stamp();
  if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[63] = (unsigned long long)prof_i;
}

}  // namespace pcaone
