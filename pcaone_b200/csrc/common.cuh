// pcaone_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#define PCA_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      throw std::runtime_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) +     \
                               " (" __FILE__ ":" + std::to_string(__LINE__) + ")");             \
  } while (0)

#define PCA_CHECK_LAUNCH() PCA_CUDA(cudaGetLastError())

namespace pcaone {

constexpr double kVarTol = 1e-9;  // VAR_TOL, reference src/Data.hpp:7
constexpr int kMaxL = 112;     // k + oversamples: 2*l*l doubles of Jacobi state must fit 227 KB of shared memory (small_dense.cuh)
constexpr int kOrthMaxL = 80;  // fused orthonormalisation: two l x 16R factor matrices must fit shared memory (orth_fused.cuh)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Per-SNP decode table: value of a genotype code after mean-imputation, centring and
// (optional) scaling. Each entry is ONE rounded IEEE op on top of F exactly as the reference
// computes it (src/FilePlink.cpp:143-147, 193-197; src/Data.cpp:356-360), with FMA contraction
// ruled out by the explicit _rn intrinsics, so decode is bit-exact.
struct SnpLut {
  double v[4];
};

struct LutParams {
  double sqrt_ploidy;  // sqrt((double)ploidy) rounded on the host like the reference
  int standardize;     // standardize && scale == -9
  int mask;            // 1: decode the missing-call indicator (code 01 -> 1, everything else 0) instead of genotypes
};

__device__ __forceinline__ double snp_scale(double F, const LutParams& p) {
  double s = 1.0;
  if (p.standardize) {
    double sd = __dsqrt_rn(__dmul_rn(F, __dsub_rn(1.0, F)));
    if (sd > kVarTol) s = __ddiv_rn(p.sqrt_ploidy, sd);
  }
  return s;
}

__device__ __forceinline__ SnpLut make_lut(double F, const LutParams& p) {
  const double s = snp_scale(F, p);
  SnpLut t;
  if (p.mask) {  // the C matrix of the reference (missing calls), for the per-sample normal equations of --project 2
    t.v[0] = 0.0;
    t.v[1] = 1.0;
    t.v[2] = 0.0;
    t.v[3] = 0.0;
    return t;
  }
  t.v[0] = __dmul_rn(__dsub_rn(1.0, F), s);  // code 00: BED2GENO 1.0
  t.v[1] = 0.0;                              // code 01: missing -> mean-imputed 0
  t.v[2] = __dmul_rn(__dsub_rn(0.5, F), s);  // code 10: 0.5
  t.v[3] = __dmul_rn(__dsub_rn(0.0, F), s);  // code 11: 0.0
  return t;
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64): with g = lane>>2, t = lane&3
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// leading dimension (in doubles) of a [rows][lp] operand tile in shared memory such that the
// DMMA B-fragment load (row = lane&3, col = lane>>2) is bank-conflict free: ld = 4 (mod 16).
__host__ __device__ constexpr int smem_ld(int lp) { return (lp + 15) / 16 * 16 + 4; }

}  // namespace pcaone
