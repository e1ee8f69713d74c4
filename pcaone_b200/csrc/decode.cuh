// pcaone_b200 — 2-bit PLINK decode kernels: repitch, allele frequency (bit-exact), dense
// block decode (the reference's read_block_initial / read_block_update output), row gather.
// HBM-bound byte/integer work: coalesced 16-byte loads, popcount reductions, no tensor cores.
#pragma once
#include "common.cuh"

namespace pcaone {

// rows of bpr bytes (as in the .bed file) -> rows of `pitch` bytes (pitch % 16 == 0) so that
// every SNP row starts 16-byte aligned; the pad bytes are zeroed.
__global__ void k_repitch(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t rows,
                          uint32_t bpr, uint32_t pitch) {
  const uint32_t chunks = pitch >> 4;
  const uint64_t total = rows * chunks;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / chunks;
    const uint32_t c = (uint32_t)(idx - r * chunks) << 4;
    const uint8_t* s = src + r * bpr + c;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      if (c + b < bpr) w[b >> 2] |= (uint32_t)s[b] << (8 * (b & 3));
    }
    *reinterpret_cast<uint4*>(dst + r * pitch + c) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// new row j = old row idx[j] (the in-core winSVD column permutation, RSVD.hpp:61-71, applied to
// the packed rows instead of a dense G*P product).
__global__ void k_gather_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                              const uint32_t* __restrict__ idx, uint64_t rows, uint32_t pitch) {
  const uint32_t chunks = pitch >> 4;
  const uint64_t total = rows * chunks;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / chunks;
    const uint32_t c = (uint32_t)(i - r * chunks);
    reinterpret_cast<uint4*>(dst + r * pitch)[c] =
        reinterpret_cast<const uint4*>(src + (uint64_t)idx[r] * pitch)[c];
  }
}
__global__ void k_gather_u32(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                             const uint32_t* __restrict__ idx, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
__global__ void k_gather_f64(const double* __restrict__ src, double* __restrict__ dst,
                             const uint32_t* __restrict__ idx, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}

// Allele frequency, one warp per SNP (FilePlink.cpp:37-60 / :165-187).
// counts: n00 (half-dosage 1), n10 (0.5), n11 (0), n01 (missing); padding samples >= N are
// masked out. F = (n00 + 0.5 n10) / (n00 + n10 + n11): the reference's running FP64 sum holds
// only multiples of 0.5 and is exact, so one IEEE division reproduces it bit for bit.
__global__ void k_allele_freq(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t N, uint64_t nsnps,
                              double* __restrict__ F, uint32_t* __restrict__ nmiss, uint32_t* __restrict__ counts = nullptr) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t nvec = (N + 63) >> 6;  // uint4 = 64 genotypes
  for (uint64_t j = warp; j < nsnps; j += nwarps) {
    const uint4* row = reinterpret_cast<const uint4*>(P + j * pitch);
    uint32_t c01 = 0, c10 = 0, c11 = 0;
    for (uint32_t v = lane; v < nvec; v += 32) {
      const uint4 q = __ldg(row + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int32_t first = (int32_t)(v * 64 + i * 16);
        int32_t nvalid = (int32_t)N - first;
        uint32_t m = 0x55555555u;
        if (nvalid <= 0)
          m = 0u;
        else if (nvalid < 16)
          m &= (1u << (2 * nvalid)) - 1u;
        const uint32_t lo = w[i] & m, hi = (w[i] >> 1) & m;
        c01 += __popc(lo & ~hi);
        c10 += __popc(hi & ~lo);
        c11 += __popc(lo & hi);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c01 += __shfl_xor_sync(0xffffffffu, c01, o);
      c10 += __shfl_xor_sync(0xffffffffu, c10, o);
      c11 += __shfl_xor_sync(0xffffffffu, c11, o);
    }
    if (lane == 0) {
      if (counts) {  // sample-sharded job: the three counts are summed over the ranks first (k_af_from_counts)
        counts[3 * j + 0] = c01;
        counts[3 * j + 1] = c10;
        counts[3 * j + 2] = c11;
        continue;
      }
      const uint32_t c = N - c01;
      const uint32_t c00 = c - c10 - c11;
      double f = 0.0;
      if (c > 0) f = __ddiv_rn(__dadd_rn((double)c00, __dmul_rn(0.5, (double)c10)), (double)c);
      F[j] = f;
      if (nmiss) nmiss[j] = c01;
    }
  }
}

// F and the missing count from (c01, c10, c11) summed over the sample shards: the same single
// IEEE division of exact integer counts as above, so F does not depend on how the samples were split.
__global__ void k_af_from_counts(const uint32_t* __restrict__ counts, uint64_t N_total, uint64_t nsnps,
                                 double* __restrict__ F, uint32_t* __restrict__ nmiss) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nsnps; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c01 = counts[3 * j], c10 = counts[3 * j + 1], c11 = counts[3 * j + 2];
    const uint32_t c = (uint32_t)N_total - c01;
    const uint32_t c00 = c - c10 - c11;
    double f = 0.0;
    if (c > 0) f = __ddiv_rn(__dadd_rn((double)c00, __dmul_rn(0.5, (double)c10)), (double)c);
    F[j] = f;
    if (nmiss) nmiss[j] = c01;
  }
}

// Squared norm of every decoded SNP column, sum_i x_ij^2 (Selection.cpp:21,31 `G.col(j).squaredNorm()`),
// from the code counts: n00 v0^2 + n10 v2^2 + n11 v3^2 (missing entries are 0). One warp per SNP.
__global__ void k_snp_sqnorm(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t N, uint64_t nsnps,
                             const double* __restrict__ F, LutParams lp, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t nvec = (N + 63) >> 6;
  for (uint64_t j = warp; j < nsnps; j += nwarps) {
    const uint4* row = reinterpret_cast<const uint4*>(P + j * pitch);
    uint32_t c01 = 0, c10 = 0, c11 = 0;
    for (uint32_t v = lane; v < nvec; v += 32) {
      const uint4 q = __ldg(row + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int32_t nvalid = (int32_t)N - (int32_t)(v * 64 + i * 16);
        uint32_t m = 0x55555555u;
        if (nvalid <= 0)
          m = 0u;
        else if (nvalid < 16)
          m &= (1u << (2 * nvalid)) - 1u;
        const uint32_t lo = w[i] & m, hi = (w[i] >> 1) & m;
        c01 += __popc(lo & ~hi);
        c10 += __popc(hi & ~lo);
        c11 += __popc(lo & hi);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c01 += __shfl_xor_sync(0xffffffffu, c01, o);
      c10 += __shfl_xor_sync(0xffffffffu, c10, o);
      c11 += __shfl_xor_sync(0xffffffffu, c11, o);
    }
    if (lane == 0) {
      const SnpLut t = make_lut(F[j], lp);
      const double c00 = (double)(N - c01 - c10 - c11);
      out[j] = c00 * t.v[0] * t.v[0] + (double)c10 * t.v[2] * t.v[2] + (double)c11 * t.v[3] * t.v[3];
    }
  }
}

__global__ void k_lookup_scale(const double* __restrict__ F, uint64_t nsnps, LutParams p,
                               double* __restrict__ lut4, double* __restrict__ scale) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nsnps; j += (uint64_t)gridDim.x * blockDim.x) {
    const double f = F[j];
    if (lut4) {  // centered_geno_lookup (4 x M, column-major): unscaled, FilePlink.cpp:193-197
      LutParams q = p;
      q.standardize = 0;
      SnpLut t = make_lut(f, q);
      lut4[4 * j + 0] = t.v[0];
      lut4[4 * j + 1] = t.v[1];
      lut4[4 * j + 2] = t.v[2];
      lut4[4 * j + 3] = t.v[3];
    }
    if (scale) scale[j] = snp_scale(f, p);
  }
}

// EMU fill (FilePlink.cpp:252-259, Data.cpp:334-348): sum_k U(i,k) S(k) V(j,k), k ascending,
// clamped to [-F, 1-F]. US = U*diag(S) is NOT pre-multiplied: the reference forms
// (U(i,k)*S(k))*V(k,j) term by term and so do we.
__device__ __forceinline__ double emu_fill(const double* __restrict__ Urow, const double* __restrict__ S,
                                           const double* __restrict__ Vrow, int k, double F) {
  double acc = 0.0;
  for (int kk = 0; kk < k; ++kk) acc += (Urow[kk] * S[kk]) * Vrow[kk];
  return fmin(fmax(acc, -F), 1.0 - F);
}

// Dense N x B block, column-major (out[j*N + i]) exactly as read_block_initial /
// read_block_update leave data->G (FilePlink.cpp:139-162, 220-298). One thread = 4 samples.
// U: [N][ldu] row-major, V: [M][ldv] row-major (device layouts), S: k.
__global__ void k_decode_block(const uint8_t* __restrict__ P, uint32_t pitch, uint32_t N, uint32_t B,
                               const double* __restrict__ F, LutParams p, int emu,
                               const double* __restrict__ U, int ldu, const double* __restrict__ S,
                               const double* __restrict__ V, int ldv, int k, double* __restrict__ out) {
  const uint32_t nq = (N + 3) >> 2;
  const uint64_t total = (uint64_t)nq * B;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = (uint32_t)(idx / nq);
    const uint32_t q = (uint32_t)(idx - (uint64_t)j * nq);
    const double f = F[j];
    const SnpLut t = make_lut(f, p);
    const double s = snp_scale(f, p);
    uint32_t byte = P[(uint64_t)j * pitch + q];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t i = 4 * q + r;
      if (i < N) {
        const uint32_t code = (byte >> (2 * r)) & 3u;
        double x = t.v[code];
        if (emu && code == 1u) x = __dmul_rn(emu_fill(U + (uint64_t)i * ldu, S, V + (uint64_t)j * ldv, k, f), s);
        out[(uint64_t)j * N + i] = x;
      }
    }
  }
}

}  // namespace pcaone
