// pcaone_b200 — error-free tensor-core GEMMs for the power-iteration products (sm_100a).
//
//   G_b = X_b^T * Omega   (reference Halko.cpp:125,150,195-196,243)
//   H  += X_b   * G_b     (reference Halko.cpp:126,151,200-202,246-248)
//
// Both are the same contraction  T[row][c] = sum_k u(row,k) * I[k][c]  of a packed 2-bit operand
// (rows x K, two bits per entry, u = popcount(code) in {0,1,2} = 2 - dosage) with a tall dense
// matrix, run on the 5th-generation tensor cores as EXACT integer arithmetic (Ozaki scheme):
//
//  * the dense operand (Omega, or W = s o G) is rounded ONCE to S signed 8-bit slices per entry
//    against a per-column power-of-two scale (k_tc_slice): X~ = I * 2^(e_c - p), p = 8S - 1,
//    I = sum_s d_s 256^(S-1-s), d_s in [-128,127]. The rounded matrix X~ is what the algorithm
//    then uses everywhere (G~ is written back), so H = X G~ holds to FP64 accuracy and the
//    rounding only perturbs the iterate the way a slightly different Omega would;
//  * the packed codes are expanded in registers to int8 (one shift + one mask per four
//    genotypes) and written straight into TENSOR MEMORY as the A operand (tcgen05.st), never
//    through shared memory; the slices are the B operand, bulk-copied (UBLKCP) into a shared
//    memory ring in the UMMA no-swizzle K-major core-matrix layout;
//  * tcgen05.mma kind::i8 accumulates s32 in TMEM — exact for K < 2^23 — and the epilogue folds
//    the S slice sums into one int64 per (row, column) and adds it to global memory with integer
//    atomics, so split-K partials combine in any order to the same bits;
//  * centring, scaling by sqrt(ploidy)/sqrt(F(1-F)) and the column-sum (rank-1) terms are applied
//    in FP64 by the finish kernels below.
//
// Missing genotypes (code 01) are mean-imputed (value 0 after centring). A range that contains
// them runs every GEMM twice on the same machinery: once with u = 0 at the missing entries and
// once with the 0/1 missing mask as the packed operand, whose product takes the missing entries
// out of the centring term (see the finish kernels). The EMU fill (a different FP64 value per
// missing entry) is added on top of these mean-imputed products by the FP64 correction kernels of
// emu_fix.cuh, which read the same tiled operands.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pcaone {
namespace tc {

constexpr int kRowTile = 128;  // rows per UMMA (M)
constexpr int kKB = 64;        // contraction entries per k-block = one 16-byte packed load per row
constexpr int kStageKB = 2;    // k-blocks per pipeline stage (128 contraction entries = 4 UMMAs per row tile)
constexpr int kNS = 4;         // pipeline stages: A ring in TMEM + B ring in shared memory
constexpr int kAcol0 = 256;    // first TMEM column of the A ring (accumulators use [0,256))
constexpr int kDG = 2;         // decode groups: group g owns the pipeline stages with (stage counter & 1) == g
constexpr int kPF = 4;         // own stages of packed loads in flight per decode thread (cp.async ring; x kDG stages of lookahead)
constexpr int kMaxNP = 256;
constexpr uint32_t kChunkBytes = kRowTile * 16;  // one (row tile, k-block) chunk of the tiled operand

__host__ __device__ constexpr int tc_threads(int RT) { return 128 + kDG * 128 * RT; }
__host__ __device__ constexpr size_t tc_scratch_bytes(int RT) { return (size_t)kDG * 4 * RT * 32 * 17 * 8; }
// packed-operand ring: every decode thread owns kPF slots of kStageKB x 16 bytes
__host__ __device__ constexpr size_t tc_ring_bytes(int RT) { return (size_t)kPF * kStageKB * kDG * 128 * RT * 16; }
__host__ __device__ inline size_t tc_smem_bytes(int RT, int NP) {
  return (size_t)kNS * kStageKB * kKB * NP + tc_scratch_bytes(RT) + tc_ring_bytes(RT);
}

// position inside a k-block at which the decode puts source entry g (see decode64)
__host__ __device__ __forceinline__ int kpos_of(int g) {
  const int x = g >> 4, rem = g & 15;
  return (x << 4) | ((rem & 3) << 2) | (rem >> 2);
}
// byte offset of element (n, kp) inside one k-block image of the B operand
__host__ __device__ __forceinline__ uint32_t bimg_offset(int n, int kp, int NP) {
  return (uint32_t)(((kp >> 4) * (NP >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (kp & 15));
}

// What the A operand holds for a 2-bit code (FilePlink.cpp:39-47; code 01 = missing):
//   kPlain   u = popcount(code)                   {00,01,10,11} -> {0,1,1,2}  (ranges WITHOUT missing calls)
//   kNonMiss u = popcount(code), 0 where missing  {00,01,10,11} -> {0,0,1,2}
//   kMask    1 where missing                      {00,01,10,11} -> {0,1,0,0}
// A range with missing calls runs the GEMM twice (kNonMiss and kMask): the mask product removes the
// missing entries from the centring term, i.e. mean imputation (FilePlink.cpp:194-197, lut[01] = 0).
enum : int { kPlain = 0, kNonMiss = 1, kMask = 2 };

// 64 two-bit codes (one uint4) -> 64 int8 values, 4 per word.
// Word 4x+w, byte b holds source entry 16x + w + 4b  (=> kpos_of).
template <int MODE>
__device__ __forceinline__ void decode64(const uint4& q, uint32_t* o) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const uint32_t lo = w[x] & 0x55555555u, hi = (w[x] >> 1) & 0x55555555u;
    const uint32_t t = MODE == kPlain ? lo + hi : MODE == kNonMiss ? hi + (hi & lo) : lo & ~hi;
    o[4 * x + 0] = t & 0x03030303u;
    o[4 * x + 1] = (t >> 2) & 0x03030303u;
    o[4 * x + 2] = (t >> 4) & 0x03030303u;
    o[4 * x + 3] = (t >> 6) & 0x03030303u;
  }
}

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

struct TcGemmArgs {
  const uint8_t* PA;              // tiled packed operand: chunk(rt, kb) = PA + rt*stride_rt + kb*stride_kb
  uint64_t stride_rt, stride_kb;  // bytes
  const int8_t* Bimg;             // k-block images of this launch, image of kb0 first, 64*NP bytes each
  uint32_t rt0, nrt;              // row tiles of the launch: [rt0, rt0 + nrt)
  uint32_t kb0, nkb;              // k-blocks of the launch: [kb0, kb0 + nkb), nkb EVEN (zero image padding)
  uint32_t kb_valid_last;         // last k-block that exists in PA (loads are clamped to it)
  uint32_t kb_per_split, nsplit;  // split-K decomposition, kb_per_split EVEN
  uint32_t NP, l, lp;             // padded UMMA N (= round_up(S*l,16)), columns, leading dim of R
  long long row_begin, row_end;   // absolute rows that receive output
  long long row_r0;               // absolute row stored at R[0]
  long long* R;                   // [rows][lp] int64 accumulators (zero on entry, added to)
  // optional: zero_n words that the kernels AFTER this launch accumulate into with atomics and that
  // nothing reads during it (the W column maxima / column sums of the range): cleared by CTA 0
  // instead of a cudaMemsetAsync node per range
  unsigned long long* zero_ptr;
  uint32_t zero_n;
};

// Four UMMAs (K = 4 x 32) of one row tile for one stage: D[d_tmem] (+)= A[a_tmem .. +32 cols] * B.
// `dk` = descriptor increment per K step (two 16-byte K chunks = 2*LBO bytes >> 4).
__device__ __forceinline__ void umma_i8_ts_x4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, uint64_t dk) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 d1, d2, d3;\n\t"
      ".reg .b32 a1, a2, a3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.u64 d1, %2, %5;\n\t"
      "add.u64 d2, d1, %5;\n\t"
      "add.u64 d3, d2, %5;\n\t"
      "add.u32 a1, %1, 8;\n\t"
      "add.u32 a2, %1, 16;\n\t"
      "add.u32 a3, %1, 24;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [a1], d1, %3, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [a2], d2, %3, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [a3], d3, %3, pt;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "l"(dk)
      : "memory");
}

// Pipeline (one CTA per SM, persistent over work items = (row-tile group, K split)):
//   warp 0      B producer   : bulk-copies one stage (128 x NP int8) of the slice image into smem
//   warp 1      UMMA issuer  : per stage 4 UMMAs per row tile, then ONE tcgen05.commit frees the stage
//   warps 4..   decode warps : kDG groups of 4 per row tile (one per TMEM lane quadrant); each thread
//                              owns one output row: 2 x 16-byte packed loads -> 128 int8 -> tcgen05.st.
//                              The groups take alternate stages: one stage costs a decode warp a
//                              SERIAL chain of ~900 cycles (barrier wait, tcgen05 fences, st + wait::st,
//                              arrive: ~600 cycles measured with the decode and the UMMAs compiled
//                              out, plus ~300 of decode) against 550 cycles of UMMA work, so a single
//                              group paced the kernel at 57-65 % tensor-pipe utilisation (ncu).
//                              The packed bytes arrive through a per-thread cp.async ring in shared
//                              memory, kPF stages (~3 us of tensor work) ahead of the decode, so the
//                              HBM latency never reaches the decode warps (round-1 ncu: the single
//                              largest stall was the scoreboard wait on a 4-deep register ring)
//   full[s]  : 4*RT decode-warp arrivals + the B producer's expect_tx arrival + the copied bytes
//   empty[s] : the issuer's tcgen05.commit
// After the last stage of an item the decode warps drain the s32 accumulators (epilogue).
template <int S, int RT, int MODE>
__global__ void __launch_bounds__(tc_threads(RT), 1) k_tc_gemm(const TcGemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kNS], empty[kNS], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t kb_bytes = kKB * a.NP;
  const uint32_t stage_bytes = kStageKB * kb_bytes;
  uint8_t* Bs = smem;
  long long* scratch = reinterpret_cast<long long*>(smem + (size_t)kNS * stage_bytes);

  if (warp == 2) tmem_alloc<512>(&tmem_slot);
  if (blockIdx.x == 0 && a.zero_ptr)
    for (uint32_t i = threadIdx.x; i < a.zero_n; i += blockDim.x) a.zero_ptr[i] = 0ull;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNS; ++i) {
      mbar_init(&full[i], 4 * RT + 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, kDG * 4 * RT);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  const uint32_t n_rtp = (a.nrt + RT - 1) / RT;
  const uint32_t n_items = n_rtp * a.nsplit;
  if (warp == 0) {
    // ------------------------------------------------ B producer
    if (lane == 0) {
      uint32_t sit = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t sp = item / n_rtp;
        const uint32_t kb_begin = a.kb0 + sp * a.kb_per_split;
        const uint32_t kb_end = min(a.kb0 + a.nkb, kb_begin + a.kb_per_split);
        const int8_t* src = a.Bimg + (size_t)(kb_begin - a.kb0) * kb_bytes;
        for (uint32_t kb = kb_begin; kb < kb_end; kb += kStageKB, ++sit, src += stage_bytes) {
          const uint32_t s = sit % kNS, ph = (sit / kNS) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], stage_bytes);
          bulk_g2s(Bs + (size_t)s * stage_bytes, src, stage_bytes, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ UMMA issuer
    // The WHOLE warp walks the loop and one elected lane issues: with warp-uniform control flow
    // the descriptor / TMEM-address arithmetic stays on the uniform datapath. (Issued from a
    // single divergent lane the loop was ~90 SASS instructions per stage, most of them R2UR
    // moves feeding UTCIMMA, and that one thread paced the tensor pipe.)
    const uint32_t idesc = idesc_i8(kRowTile, (int)a.NP);
    const uint32_t lbo = (a.NP >> 3) * 128, sbo = 128;
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(Bs), lbo, sbo);
    const uint64_t dk = (uint64_t)((2 * lbo) >> 4);         // one K step (32 entries)
    const uint64_t dstage = (uint64_t)(stage_bytes >> 4);   // one stage
    uint32_t sit = 0, n = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t sp = item / n_rtp, rtp = item - sp * n_rtp;
      const uint32_t kb_begin = a.kb0 + sp * a.kb_per_split;
      const uint32_t kb_end = min(a.kb0 + a.nkb, kb_begin + a.kb_per_split);
      const uint32_t nst = (kb_end - kb_begin + kStageKB - 1) / kStageKB;
      const bool two = RT > 1 && (a.nrt - rtp * RT) > 1;
      mbar_wait(&acc_empty, (n & 1) ^ 1);
      tc_fence_after();
      uint32_t accflag = 0;
#pragma unroll 1
      for (uint32_t st = 0; st < nst; ++st, ++sit) {
        const uint32_t s = sit % kNS, ph = (sit / kNS) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bd = desc0 + s * dstage;
          const uint32_t at = tbase + kAcol0 + s * (RT * 32);
          umma_i8_ts_x4(tbase, at, bd, idesc, accflag, dk);
          if (two) umma_i8_ts_x4(tbase + a.NP, at + 32, bd, idesc, accflag, dk);
          umma_commit(&empty[s]);
        }
        __syncwarp();
        accflag = 1;
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ decode warps (A producer) + epilogue
    const int dw = warp - 4;
    const int g = dw / (4 * RT);  // decode group
    const int t = (dw % (4 * RT)) >> 2, q = warp & 3;
    const uint32_t rloc = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long long* my_scratch = scratch + (size_t)dw * 32 * 17;
    const uint32_t ring_hstride = kDG * 128 * RT * 16;
    const uint32_t ring0 = smem_u32(smem + (size_t)kNS * stage_bytes + tc_scratch_bytes(RT)) + (threadIdx.x - 128) * 16;
    uint32_t sit0 = 0, n = 0;  // sit0: stage counter at the start of the item (all groups count every stage)
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t sp = item / n_rtp, rtp = item - sp * n_rtp;
      const uint32_t kb_begin = a.kb0 + sp * a.kb_per_split;
      const uint32_t kb_end = min(a.kb0 + a.nkb, kb_begin + a.kb_per_split);
      const uint32_t ntile = min((uint32_t)RT, a.nrt - rtp * RT);
      const bool active = (uint32_t)t < ntile;
      const uint32_t rt = a.rt0 + rtp * RT + t;
      const uint32_t nst = (kb_end - kb_begin + kStageKB - 1) / kStageKB;
      const uint32_t first = ((sit0 & 1u) == (uint32_t)g) ? 0u : 1u;  // my first stage of this item
      if (active) {
        const uint8_t* p = a.PA + (uint64_t)rt * a.stride_rt + (uint64_t)rloc * 16;
        // cp.async ring: slot i of this thread = ring0 + (i * kStageKB + h) * ring_hstride.
        // One commit group per own stage, committed even when the stage is past the item's end (an
        // empty group), so that wait_group<kPF-1> always means "my oldest stage has landed".
        auto issue = [&](uint32_t st, uint32_t slot) {
          if (st < nst) {
#pragma unroll
            for (int h = 0; h < kStageKB; ++h)
              cp_async16(ring0 + (slot * kStageKB + h) * ring_hstride,
                         p + (uint64_t)min(kb_begin + kStageKB * st + h, a.kb_valid_last) * a.stride_kb);
          }
          cp_async_commit();
        };
#pragma unroll 1
        for (uint32_t j = 0; j < (uint32_t)kPF; ++j) issue(first + kDG * j, j);
        uint32_t slot = 0;
        bool st_pending = false;
        uint32_t pend_s = 0;
        // the arrival for a stage is deferred until my next stage has been decoded, so the
        // tcgen05.st latency overlaps that decode instead of stalling the warp
        auto publish = [&]() {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[pend_s]);
        };
#pragma unroll 1
        for (uint32_t st = first; st < nst; st += kDG) {
          cp_async_wait<kPF - 1>();
          const uint4 q0 = lds128(ring0 + (slot * kStageKB + 0) * ring_hstride);
          const uint4 q1 = lds128(ring0 + (slot * kStageKB + 1) * ring_hstride);
          uint32_t o[32];
          decode64<MODE>(q0, o);
          decode64<MODE>(q1, o + 16);
          issue(st + kDG * kPF, slot);  // refill the slot just consumed (its bytes are in registers)
          slot = (slot + 1 == (uint32_t)kPF) ? 0 : slot + 1;
          if (st_pending) publish();
          const uint32_t sit = sit0 + st;
          const uint32_t s = sit % kNS, ph = (sit / kNS) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          tc_fence_after();
          tmem_st32(tbase + lane_addr + kAcol0 + s * (RT * 32) + t * 32, o);
          st_pending = true;
          pend_s = s;
        }
        if (st_pending) publish();
        cp_async_wait<0>();  // only empty groups can be pending here
      } else {
        // idle tile of a ragged last group: keep the stage barriers in step
        for (uint32_t st = first; st < nst; st += kDG) {
          const uint32_t sit = sit0 + st;
          const uint32_t s = sit % kNS, ph = (sit / kNS) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
        }
      }
      sit0 += nst;
      // ---- epilogue: s32 slice sums -> one int64 per (row, column) -> integer atomics
      mbar_wait(&acc_full, n & 1);
      tc_fence_after();
      const long long row0 = (long long)rt * kRowTile + q * 32;  // first row of this warp
      for (uint32_t c0 = 16 * g; c0 < a.l; c0 += 16 * kDG) {  // the groups take alternate 16-column blocks
        if (active) {
          uint32_t v[S][16];
#pragma unroll
          for (int s = 0; s < S; ++s) tmem_ld16(tbase + lane_addr + t * a.NP + c0 * S + 16 * s, v[s]);
          tmem_wait_ld();
#pragma unroll
          for (int ci = 0; ci < 16; ++ci) {
            long long acc = 0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int f = ci * S + s;
              acc += (long long)(int32_t)v[f >> 4][f & 15] << (8 * (S - 1 - s));
            }
            my_scratch[lane * 17 + ci] = acc;
          }
          __syncwarp();
          // two rows per step: lanes 0-15 row rr, lanes 16-31 row rr+1; 16 consecutive columns each
          const int ci = lane & 15;
#pragma unroll 4
          for (int rr = 0; rr < 32; rr += 2) {
            const int r = rr + (lane >> 4);
            const long long row = row0 + r;
            const long long val = my_scratch[r * 17 + ci];
            if (row >= a.row_begin && row < a.row_end && c0 + ci < a.l && val != 0)
              atomicAdd(reinterpret_cast<unsigned long long*>(a.R + (row - a.row_r0) * a.lp + c0 + ci),
                        (unsigned long long)val);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tbase);
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------

// Row-tiled copy of the SNP-major packed rows ([rows][pitch]) into a tiling whose row 0 is `row0` rows
// before P's first row: chunk (rt, kb) holds, for each of 128 rows, the 16 bytes of k-block kb;
// dst chunk = dst + (rt * nkb + kb) * 2 KB. Entries beyond `ncols` (sample padding of the last
// byte; the reference ignores those bits, FilePlink.cpp:42) are code 00 (u = 0). Rows of the first /
// last tile that P does not cover are left as they are: the GEMM masks its output rows to the range,
// and an output row depends on its own operand row only.
__global__ void k_tile_rows(const uint8_t* __restrict__ P, uint32_t pitch, uint64_t rows, uint32_t ncols,
                            uint32_t nkb, uint8_t* __restrict__ dst, uint64_t row0) {
  const uint64_t total = (uint64_t)((rows + kRowTile - 1) / kRowTile) * nkb * kRowTile;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    // consecutive threads: consecutive rows of one k-block (16-byte stores side by side in the chunk)
    const uint64_t blk = idx / kRowTile, r_in = idx % kRowTile;
    const uint64_t rgrp = blk / nkb;
    const uint32_t kb = (uint32_t)(blk % nkb);
    const uint64_t lrow = rgrp * kRowTile + r_in;  // row of P
    if (lrow >= rows) continue;
    const uint64_t grow = row0 + lrow;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (kb * 16 < pitch) {
      v = *reinterpret_cast<const uint4*>(P + lrow * pitch + kb * 16);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const long long first = (long long)kb * 64 + x * 16;
        const long long nvalid = (long long)ncols - first;
        if (nvalid <= 0)
          w[x] = 0;
        else if (nvalid < 16)
          w[x] &= (1u << (2 * nvalid)) - 1u;
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(dst + ((grow / kRowTile) * nkb + kb) * kChunkBytes + (grow % kRowTile) * 16) = v;
  }
}

// Transposed tiled copy: rows = samples, contraction = SNPs. chunk (kb, rt) = dst + (kb * nrt + rt)
// * 2 KB; row r = sample rt*128 + r, byte q = SNPs kb*64 + 4q .. +3 (two bits each, same bit
// order as the bed file along its own axis). P holds SNPs [snp0, snp0 + nsnps) of the tiling; block
// (x, rt) builds k-block snp0/64 + x. A k-block that P covers only partly (a streamed block that
// starts or ends inside it) is merged into what is there: the bits of the other SNPs are kept (the
// neighbouring block wrote or will write them; the H-pass image is zero outside its range, so
// whatever they hold never contributes). Samples >= N are code 00.
__global__ void __launch_bounds__(128) k_tile_transpose(const uint8_t* __restrict__ P, uint32_t pitch, uint64_t nsnps,
                                                        uint32_t N, uint32_t nrt, uint8_t* __restrict__ dst,
                                                        uint64_t snp0) {
  __shared__ uint32_t tile[kKB][9];  // 64 SNP rows x 32 bytes (128 samples), padded
  const uint32_t rt = blockIdx.x % nrt;
  const uint64_t kb = snp0 / kKB + blockIdx.x / nrt;
  const int tid = threadIdx.x;
  const long long g_lo = (long long)snp0 - (long long)(kb * kKB);            // first covered SNP slot of this k-block (may be < 0)
  const long long g_hi = (long long)(snp0 + nsnps) - (long long)(kb * kKB);  // one past the last covered slot (may be > 64)
  for (int i = tid; i < kKB * 8; i += 128) {
    const int g = i >> 3, wq = i & 7;
    const uint32_t byte0 = rt * 32 + wq * 4;
    uint32_t w = 0;
    if (g >= g_lo && g < g_hi && byte0 < pitch)
      w = *reinterpret_cast<const uint32_t*>(P + (uint64_t)((long long)(kb * kKB) + g - (long long)snp0) * pitch + byte0);
    tile[g][wq] = w;
  }
  __syncthreads();
  const uint32_t sample = rt * kRowTile + tid;
  uint32_t out[4] = {0, 0, 0, 0};
  if (sample < N) {
    const int wq = tid >> 4, sh = (tid & 15) * 2;
#pragma unroll 16
    for (int g = 0; g < kKB; ++g) {
      const uint32_t code = (tile[g][wq] >> sh) & 3u;
      out[g >> 4] |= code << (2 * (g & 15));
    }
  }
  uint4* d = reinterpret_cast<uint4*>(dst + (kb * nrt + rt) * kChunkBytes + tid * 16);
  if (g_lo > 0 || g_hi < kKB) {
    uint32_t m[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      uint32_t mm = 0;
      for (int b2 = 0; b2 < 16; ++b2) {
        const int g = 16 * x + b2;
        if (g >= g_lo && g < g_hi) mm |= 3u << (2 * b2);
      }
      m[x] = mm;
    }
    const uint4 o = *d;
    *d = make_uint4((o.x & ~m[0]) | out[0], (o.y & ~m[1]) | out[1], (o.z & ~m[2]) | out[2], (o.w & ~m[3]) | out[3]);
  } else {
    *d = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// column abs-max of X[r0..r1)[0..l) -> colmax bits (atomicMax on the IEEE pattern, which is
// order-preserving for non-negative doubles). One warp per row, lanes across columns.
__global__ void __launch_bounds__(256) k_tc_colmax(const double* __restrict__ X, int lp, int l, uint64_t r0, uint64_t r1,
                                                    unsigned long long* __restrict__ colmax) {
  __shared__ unsigned long long smax[kMaxNP];
  for (int c = threadIdx.x; c < l; c += blockDim.x) smax[c] = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t gw = (uint64_t)blockIdx.x * 8 + wib, nw = (uint64_t)gridDim.x * 8;
  double m[4] = {0.0, 0.0, 0.0, 0.0};
  for (uint64_t r = r0 + gw; r < r1; r += nw) {
    const double* row = X + r * lp;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      if (c < l) m[q] = fmax(m[q], fabs(row[c]));
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = lane + 32 * q;
    if (c < l && m[q] > 0.0) atomicMax(&smax[c], (unsigned long long)__double_as_longlong(m[q]));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < l; c += blockDim.x)
    if (smax[c]) atomicMax(&colmax[c], smax[c]);
}

// power-of-two column scale: X~ = I * 2^(e - p) with |I| < 127.5 * 256^(S-1)
__device__ __forceinline__ int tc_exponent(unsigned long long maxbits) {
  const double m = __longlong_as_double((long long)maxbits);
  if (!(m > 0.0)) return 0;
  return ilogb(m * (128.0 / 126.0)) + 1;
}

struct TcSliceArgs {
  double* X;                          // [rows][lp], rows indexed by the contraction index
  int lp, l, S, NP;
  uint64_t r0, r1;                    // contraction entries [r0, r1) are live, the rest of the k-blocks is zero
  uint32_t kb0;                       // first k-block of the image (= r0 / 64)
  const unsigned long long* colmax;   // per-column abs max bits
  const double* F;                    // per-row allele frequency (indexed like X rows) or nullptr
  LutParams lut;
  int writeback;                      // X[row][c] <- X~[row][c] / s_row  (G~ for the QR stage)
  int dmode;                          // slice D = (f_row - 1) * s_row * X instead (mask operand of the H pass); image only
  int8_t* Bimg;                       // out: [nkb][64*NP]
  long long* Csum;                    // out (atomic): sum_rows I[row][c]
  double* Fpart;                      // out: [gridDim.x][lp] partial sums of f_row * X~[row][c], or nullptr
  double* Fw;                         // out: [lp] sum of Fpart over the blocks in a fixed order (last block done), or nullptr
  unsigned int* done;                 // block counter of the Fw reduction: zero on entry, left zero
};

// Threads per block of the slice / finish kernels: a multiple of lp, so that with the flat index
// e = tid + B*i every thread stays on ONE column (c = tid % lp) and walks rows g0, g0 + B/lp, ...
__host__ __device__ inline int tc_flat_threads(int lp) { return lp * (256 / lp); }

// One block per k-block (64 rows x l columns). X rows are G (H pass: F != nullptr, the kernel
// forms W = s_row * G itself) or Omega (G pass). The 64 rows are one contiguous run of 64*lp
// doubles: loads and the write-back are flat and coalesced; column sums live in registers.
__global__ void __launch_bounds__(256) k_tc_slice(const TcSliceArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  int8_t* img = reinterpret_cast<int8_t*>(sm);  // 64*NP
  __shared__ double s_scale[kKB], s_inv[kKB], s_f[kKB];
  __shared__ double s_pf[256];      // [rows-per-pass][lp] partial sums (rpp * lp <= 256)
  __shared__ long long s_pc[256];
  const uint32_t kb = a.kb0 + blockIdx.x;
  const int tid = threadIdx.x, B = blockDim.x;
  const int p = 8 * a.S - 1;
  const uint64_t row0 = (uint64_t)kb * kKB;
  // The X loads of a batch are issued before their first use, the first batch before anything
  // else: with the write-back to the same array inside the loop the compiler kept every load
  // behind the previous store, a chain of 64 / rpp dependent memory round trips per thread
  // (11 of the 13.5 us of a window-sized launch, ncu source page).
  constexpr int UB = 8;
  const int rpp = B / a.lp;  // rows per pass
  const int g0 = tid / a.lp, c = tid - g0 * a.lp;
  double* Xb = a.X + row0 * a.lp;
  double xv[UB];
  auto load_batch = [&](int gb) {
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int g = gb + u * rpp;
      const uint64_t row = row0 + g;
      xv[u] = (g < kKB && row >= a.r0 && row < a.r1) ? Xb[g * a.lp + c] : 0.0;
    }
  };
  if (c < a.l) load_batch(g0);
  for (int i = tid; i < kKB * a.NP / 16; i += B) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  if (tid < kKB) {
    const uint64_t row = row0 + tid;
    double s = 1.0, f = 0.0;
    if (row >= a.r0 && row < a.r1 && a.F) {
      f = a.F[row];
      s = snp_scale(f, a.lut);
    }
    s_scale[tid] = s;
    s_inv[tid] = 1.0 / s;  // the write-back G~ = W~ / s as one multiplication per element
    s_f[tid] = f;
  }
  double up = 0.0, dn = 0.0;
  if (c < a.l) {
    const int e = tc_exponent(a.colmax[c]);
    up = scalbn(1.0, p - e);
    dn = scalbn(1.0, e - p);
  }
  __syncthreads();
  long long csum = 0;
  double fsum = 0.0;
  if (c < a.l) {
    for (int gb = g0; gb < kKB; gb += rpp * UB) {
      if (gb != g0) load_batch(gb);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int g = gb + u * rpp;
        const uint64_t row = row0 + g;
        if (!(g < kKB && row >= a.r0 && row < a.r1)) continue;
        double x = xv[u] * s_scale[g];
        if (a.dmode) x *= s_f[g] - 1.0;
        // round to nearest even through the 1.5 * 2^52 constant: |I| < 127.5 * 256^(S-1) < 2^31, so
        // the integer is the low word of the sum and the rounded double is (sum - constant) — the
        // same I as llrint() without the F2I / I2F sequences (this loop is issue bound on the
        // merged ranges: ~1300 instructions per thread, 167 us per half-shard launch)
        const double tm = x * up + 6755399441055744.0;
        const int I = __double2loint(tm);
        const double xt = (tm - 6755399441055744.0) * dn;
        if (a.writeback) Xb[g * a.lp + c] = xt * s_inv[g];
        csum += I;
        fsum += s_f[g] * xt;
        if (I != 0) {
          const int kp = kpos_of(g);
          int rem = I;
          for (int s = a.S - 1; s > 0; --s) {
            const int d = ((rem + 128) & 255) - 128;
            rem = (rem - d) >> 8;
            img[bimg_offset(c * a.S + s, kp, a.NP)] = (int8_t)d;
          }
          img[bimg_offset(c * a.S, kp, a.NP)] = (int8_t)rem;
        }
      }
    }
  }
  s_pc[g0 * a.lp + c] = csum;
  s_pf[g0 * a.lp + c] = fsum;
  __syncthreads();
  for (int i = tid; i < kKB * a.NP / 16; i += B)
    reinterpret_cast<uint4*>(a.Bimg + (size_t)blockIdx.x * kKB * a.NP)[i] = reinterpret_cast<const uint4*>(img)[i];
  if (g0 == 0 && c < a.l && !a.dmode) {
    long long tc_ = 0;
    double tf = 0.0;
    for (int q = 0; q < rpp; ++q) {  // fixed order
      tc_ += s_pc[q * a.lp + c];
      tf += s_pf[q * a.lp + c];
    }
    if (tc_) atomicAdd(reinterpret_cast<unsigned long long*>(a.Csum + c), (unsigned long long)tc_);
    if (a.Fpart) a.Fpart[(size_t)blockIdx.x * a.lp + c] = tf;
  }
  // Fw[c] = sum of the per-block partials, by whichever block finishes last, always in the same
  // order (was a separate 5-block launch, ~6 us of pure latency per window).
  if (a.Fpart && a.Fw) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      // thread (g0, c) sums parts g0, g0 + rpp, ... of column c (loads in batches, adds in order),
      // then the rpp partial sums of a column are folded in order
      double v = 0.0;
      if (c < a.l) {
        constexpr int UF = 16;
        for (uint32_t b0 = g0; b0 < gridDim.x; b0 += rpp * UF) {
          double t[UF];
#pragma unroll
          for (int u = 0; u < UF; ++u) {
            const uint32_t b = b0 + u * rpp;
            t[u] = b < gridDim.x ? __ldcg(a.Fpart + (size_t)b * a.lp + c) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < UF; ++u) v += t[u];
        }
      }
      s_pf[g0 * a.lp + c] = v;
      __syncthreads();
      if (g0 == 0 && c < a.l) {
        double tf = 0.0;
        for (int q = 0; q < rpp; ++q) tf += s_pf[q * a.lp + c];
        a.Fw[c] = tf;
      }
      if (tid == 0) *a.done = 0;
    }
  }
}

// threads per block of k_tc_finish_g: every thread owns a PAIR of adjacent columns (16-byte
// accesses) and walks rows g0, g0 + rows-per-pass, ...
__host__ __device__ inline int tc_pair_threads(int lp) { return (lp / 2) * (256 / (lp / 2)); }

// G pass finish: G[j][c] = s_j * 2^(e_c-p) * ((1 - f_j) * (C_c - K[j][c]) - T[j][c] / 2) written to Gout rows;
// K = R2 = mask product (sum of the Omega~ integers over the samples whose call is missing at SNP j),
// nullptr for ranges without missing calls. C_c - K is an exact integer below 2^53
// (the slice kernel turns s o G into the rounded G~ afterwards); also the column abs-max of
// W = s o G for that slicing. R is re-zeroed. Blocks walk 64-row groups (contiguous 64*lp values);
// all loads of a group are issued before the first use (the int64 accumulators come back from L2
// with long latency right after the GEMM's atomics), running maxima stay in registers.
__global__ void __launch_bounds__(256)
k_tc_finish_g(long long* __restrict__ R, long long* __restrict__ R2, uint64_t nrows, int l, int lp, int S, const double* __restrict__ F,
              LutParams lut, const long long* __restrict__ Csum, const unsigned long long* __restrict__ colmax_in,
              double* __restrict__ Gout, unsigned long long* __restrict__ colmax_out) {
  constexpr int U = 8;  // row passes per group held in registers (rows-per-pass >= 8 -> 64 rows)
  __shared__ double s_s[kKB], s_omf[kKB];
  const int tid = threadIdx.x, B = blockDim.x;
  const int p = 8 * S - 1;
  const int hp = lp >> 1;
  const int rpp = B / hp;  // rows per pass, >= 8 for lp <= 64, >= 4 for lp <= 128
  const int g0 = tid / hp, c = 2 * (tid - g0 * hp);
  double cs[2] = {0.0, 0.0}, sc[2] = {0.0, 0.0}, m[2] = {0.0, 0.0};
  long long csi[2] = {0, 0};
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (c + h < l) {
      csi[h] = Csum[c + h];
      cs[h] = (double)csi[h];
      sc[h] = scalbn(1.0, tc_exponent(colmax_in[c + h]) - p);
    }
  for (uint64_t row0 = (uint64_t)blockIdx.x * kKB; row0 < nrows; row0 += (uint64_t)gridDim.x * kKB) {
    __syncthreads();
    if (tid < kKB) {
      double f = 0.0, sj = 1.0;
      if (row0 + tid < nrows) {
        f = F[row0 + tid];
        sj = snp_scale(f, lut);
      }
      s_s[tid] = sj;
      s_omf[tid] = 1.0 - f;
    }
    const int nr = (int)min((uint64_t)kKB, nrows - row0);
    longlong2* Rb = reinterpret_cast<longlong2*>(R + row0 * lp);
    longlong2* R2b = R2 ? reinterpret_cast<longlong2*>(R2 + row0 * lp) : nullptr;
    double2* Gb = reinterpret_cast<double2*>(Gout + row0 * lp);
    longlong2 T[U], K[U];
    int gb = g0;
    auto load_batch = [&]() {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = gb + u * rpp;
        T[u] = (g < nr) ? Rb[g * hp + (c >> 1)] : make_longlong2(0, 0);
        K[u] = (R2b && g < nr) ? R2b[g * hp + (c >> 1)] : make_longlong2(0, 0);
      }
    };
    load_batch();     // in flight while the first 64 threads finish the per-row scales
    __syncthreads();  // s_s / s_omf of this group are ready
    while (true) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = gb + u * rpp;
        if (g < nr) {
          Rb[g * hp + (c >> 1)] = make_longlong2(0, 0);
          double c0 = cs[0], c1 = cs[1];
          if (R2b) {
            R2b[g * hp + (c >> 1)] = make_longlong2(0, 0);
            c0 = (double)(csi[0] - K[u].x);
            c1 = (double)(csi[1] - K[u].y);
          }
          const double sj = s_s[g], omf = s_omf[g];
          double2 gv;
          gv.x = (c < l) ? ((omf * c0 - 0.5 * (double)T[u].x) * sc[0]) * sj : 0.0;
          gv.y = (c + 1 < l) ? ((omf * c1 - 0.5 * (double)T[u].y) * sc[1]) * sj : 0.0;
          m[0] = fmax(m[0], fabs(gv.x * sj));
          m[1] = fmax(m[1], fabs(gv.y * sj));
          Gb[g * hp + (c >> 1)] = gv;
        }
      }
      gb += rpp * U;
      if (gb >= nr) break;
      load_batch();
    }
  }
  // the rows-per-pass lanes of a column pair fold in shared memory first: one atomicMax per column
  // and block. (Every thread issuing its own pair was 245 blocks x 240 threads x 2 atomics on three
  // cache lines, serialised in L2: most of the 27 us of a window-sized launch.)
  __shared__ double s_m[256][2];
  s_m[tid][0] = m[0];
  s_m[tid][1] = m[1];
  __syncthreads();
  if (g0 == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double v = 0.0;
      for (int q = 0; q < rpp; ++q) v = fmax(v, s_m[q * hp + (c >> 1)][h]);
      if (c + h < l && v > 0.0) atomicMax(&colmax_out[c + h], (unsigned long long)__double_as_longlong(v));
    }
  }
}

// Fw[c] = sum over the slice kernel's per-block partials (merged ranges: thousands of them), in a
// fixed order: one block per column, thread t sums parts t, t + 256, ... (8 loads in flight), then a
// shared-memory tree. (One warp per column on 5 blocks was 63 us per half-shard launch.)
__global__ void __launch_bounds__(256) k_tc_reduce_fpart(const double* __restrict__ Fpart, uint32_t nparts, int l, int lp,
                                                          double* __restrict__ Fw) {
  __shared__ double s[256];
  const int c = blockIdx.x, t = threadIdx.x;
  if (c >= l) return;
  constexpr int UF = 8;
  double v = 0.0;
  for (uint32_t b0 = t; b0 < nparts; b0 += 256 * UF) {
    double x[UF];
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      const uint32_t b = b0 + 256 * u;
      x[u] = b < nparts ? Fpart[(size_t)b * lp + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UF; ++u) v += x[u];
  }
  s[t] = v;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) s[t] += s[t + w];
    __syncthreads();
  }
  if (t == 0) Fw[c] = s[0];
}

// H pass finish: Hacc[i][c] (+)= 2^(e_c-p) * (Cw_c - T[i][c] / 2) - Fw_c. R is re-zeroed. Optionally
// also Hsum = Hacc + Hother (the H = H1 + H2 of a winSVD Omega update) in the same sweep.
// Ranges with missing calls add + 2^(e_c-p) * K[i][c], K = R2 = mask product with the image of
// D~ = round((f_j - 1) W~_j): removes (1 - f_j) W~_j for the SNPs j missing in sample i.
__global__ void k_tc_finish_h(long long* __restrict__ R, long long* __restrict__ R2, uint64_t nrows, int l, int lp, int S,
                              const long long* __restrict__ Csum, const unsigned long long* __restrict__ colmax,
                              const double* __restrict__ Fw, const double* __restrict__ Fpart, uint32_t nparts,
                              double* __restrict__ Hacc, int accumulate, const double* __restrict__ Hother,
                              double* __restrict__ Hsum) {
  __shared__ double sFw[kMaxNP], sScale[kMaxNP], sC[kMaxNP];
  const int p = 8 * S - 1;
  // Batches of U elements per thread with every load of a batch issued before the first use, and
  // the first batch before the per-column constants are formed: with one element per thread and
  // the loads interleaved with the stores (R, Hacc, Hother came back one after the other) a
  // window-sized launch was three dependent memory round trips per block behind the preamble,
  // 23 us for 19 MB of traffic (ncu source page: 66 % of the samples on the Hacc add).
  constexpr int U = 4;
  const uint64_t total = nrows * (uint64_t)lp;
  const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t first = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t iters = (total + nth * U - 1) / (nth * U);
  long long T[U], K[U];
  double ha[U], ho[U];
  int cc[U];
  auto load_batch = [&](uint64_t it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t idx = first + (it * U + u) * nth;
      const bool in = idx < total;
      cc[u] = in ? (int)(idx % lp) : lp;
      const bool live = in && cc[u] < l;
      T[u] = live ? R[idx] : 0;
      K[u] = (live && R2) ? R2[idx] : 0;
      ha[u] = (in && accumulate) ? Hacc[idx] : 0.0;
      ho[u] = (in && Hsum) ? Hother[idx] : 0.0;
    }
  };
  load_batch(0);
  for (int c = threadIdx.x; c < l; c += blockDim.x) {
    double fw;
    if (Fpart) {
      // Fw[c] = sum of the slice kernel's per-block partials, redone by every block in the same
      // fixed order (few partials: a window) instead of a separate reduction launch
      fw = 0.0;
      for (uint32_t b = 0; b < nparts; ++b) fw += Fpart[(size_t)b * lp + c];
    } else {
      fw = Fw[c];
    }
    sFw[c] = fw;
    sScale[c] = scalbn(1.0, tc_exponent(colmax[c]) - p);
    sC[c] = (double)Csum[c];
  }
  __syncthreads();
  for (uint64_t it = 0; it < iters; ++it) {
    if (it) load_batch(it);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t idx = first + (it * U + u) * nth;
      if (idx >= total) continue;
      const int c = cc[u];
      double out = 0.0;
      if (c < l) {
        R[idx] = 0;
        double h = sScale[c] * (sC[c] - 0.5 * (double)T[u]) - sFw[c];
        if (R2) {
          h += sScale[c] * (double)K[u];
          R2[idx] = 0;
        }
        out = accumulate ? ha[u] + h : h;
        Hacc[idx] = out;
      } else if (!accumulate) {
        Hacc[idx] = 0.0;
      } else {
        out = ha[u];
      }
      if (Hsum) Hsum[idx] = out + ho[u];  // H = H1 + H2 for the Omega update that follows (Halko.cpp:209)
    }
  }
}

}  // namespace tc
}  // namespace pcaone
