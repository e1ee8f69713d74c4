// Hardware probe for the tensor-core path (run under gpurun; not part of the product):
//  1. checks the operand conventions tc_gemm.cuh relies on — A (int8) written to TMEM with
//     tcgen05.st.32x32b, B (int8) in shared memory in the no-swizzle K-major core-matrix layout,
//     kind::i8 UMMA with s32 accumulators read back with tcgen05.ld — against a CPU product;
//  2. measures the sustained kind::i8 UMMA rate (TS mode) for N = 64/128/256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/probe_umma tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pcaone_b200/csrc/tc_ptx.cuh"

using namespace pcaone::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

// A: [128][64] int8 row-major (global). Bimg: N*64 bytes already in the smem image layout.
// out: [128][N] int32.
__global__ void __launch_bounds__(128, 1)
k_probe(const int8_t* __restrict__ A, const int8_t* __restrict__ Bimg, int N, uint32_t lbo, uint32_t sbo,
        int32_t* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_base_slot);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < N * 64 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_slot;
  // A row of this thread -> 16 columns
  uint32_t v[16];
  const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)threadIdx.x * 64);
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = arow[c];
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  tmem_st16(tbase + lane_base + 0, v);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = idesc_i8(128, N);
    for (int ks = 0; ks < 2; ++ks) {
      // K step = 32 int8 = two 16-byte chunks: advance the B start address by 2 chunk strides
      const uint32_t kstride = (lbo > sbo) ? lbo : sbo;  // the larger one is the K-chunk stride in both variants
      const uint64_t bd = smem_desc_kmajor_noswizzle(smem_u32(smem) + ks * 2 * kstride, lbo, sbo);
      umma_i8_ts(tbase + 256, tbase + ks * 8, bd, idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + lane_base + 256 + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 16; ++c) out[(size_t)threadIdx.x * N + c0 + c] = (int32_t)r[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// throughput: every CTA issues `iters` UMMAs (M=128, N, K=32), cycling 4 B stages / 4 A stages
__global__ void __launch_bounds__(128, 1) k_rate(int N, int iters, int32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_slot);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 4 * N * 64 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_slot;
  uint32_t v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = 0x01010101u;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int s = 0; s < 4; ++s) tmem_st16(tbase + lane_base + 16 * s, v);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = idesc_i8(128, N);
    const uint32_t lbo = (N / 8) * 128, sbo = 128;
    for (int it = 0; it < iters; ++it) {
      const int st = (it >> 1) & 3, ks = it & 1;
      const uint64_t bd = smem_desc_kmajor_noswizzle(smem_u32(smem) + st * N * 64 + ks * 2 * lbo, lbo, sbo);
      umma_i8_ts(tbase + 256, tbase + st * 16 + ks * 8, bd, idesc, it > 0);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[8];
  tmem_ld8(tbase + lane_base + 256, r);
  tmem_wait_ld();
  if (sink) sink[blockIdx.x * 128 + threadIdx.x] = (int32_t)r[0];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  // ---------------- 1. layout check
  for (int N : {64, 128, 240}) {
    std::vector<int8_t> A(128 * 64), B((size_t)N * 64);
    srand(1234 + N);
    for (auto& a : A) a = (int8_t)(rand() % 7 - 3);
    for (auto& b : B) b = (int8_t)(rand() % 256 - 128);
    std::vector<int32_t> ref((size_t)128 * N);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        int32_t s = 0;
        for (int k = 0; k < 64; ++k) s += (int32_t)A[m * 64 + k] * (int32_t)B[(size_t)n * 64 + k];
        ref[(size_t)m * N + n] = s;
      }
    // image: core matrix (n8, kc) at ((kc * N/8) + n8) * 128, inside (n%8)*16 + k%16
    std::vector<int8_t> img((size_t)N * 64);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 64; ++k)
        img[((size_t)(k / 16) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 16)] = B[(size_t)n * 64 + k];
    int8_t *dA, *dB;
    int32_t* dO;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, img.size()));
    CK(cudaMalloc(&dO, ref.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice));
    for (int variant = 0; variant < 1; ++variant) {
      const uint32_t kst = (N / 8) * 128;
      const uint32_t lbo = variant == 0 ? kst : 128, sbo = variant == 0 ? 128 : kst;
      CK(cudaMemset(dO, 0xff, ref.size() * 4));
      CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      k_probe<<<1, 128, N * 64, 0>>>(dA, dB, N, lbo, sbo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("layout N=%d variant %d: kernel error %s\n", N, variant, cudaGetErrorString(e));
        return 1;
      }
      std::vector<int32_t> got(ref.size());
      CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (size_t i = 0; i < ref.size(); ++i) bad += got[i] != ref[i];
      printf("layout N=%d variant %d (lbo=%u sbo=%u): %zu / %zu mismatches", N, variant, lbo, sbo, bad, ref.size());
      if (bad) printf("  e.g. got[0..3]= %d %d %d %d ref= %d %d %d %d", got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3]);
      printf("\n");
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dO);
  }

  // ---------------- 2. rate
  int32_t* sink;
  CK(cudaMalloc(&sink, (size_t)prop.multiProcessorCount * 2 * 128 * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int N : {64, 128, 256}) {
    const int iters = 1 << 16;
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * N * 64));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      k_rate<<<prop.multiProcessorCount, 128, 4 * N * 64>>>(N, iters, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * 128 * N * 32 * (double)iters * prop.multiProcessorCount;
      printf("rate N=%d: %.3f ms  %.1f TOP/s  (%.1f cycles/UMMA at 1.965 GHz)\n", N, ms, ops / ms * 1e-9,
             ms * 1e-3 * 1.965e9 / iters);
    }
  }
  // ---------------- 3. sustained rate under the power cap: N = 240 (the bench's 3 slices x l = 80), ~3 s back to back
  {
    const int N = 240, iters = 1 << 18;
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * N * 64));
    double burst = 0.0, sustained = 0.0;
    float total_ms = 0.f;
    int reps = 0;
    while (total_ms < 3000.f && reps < 2000) {
      CK(cudaEventRecord(e0));
      k_rate<<<prop.multiProcessorCount, 128, 4 * N * 64>>>(N, iters, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double tops = 2.0 * 128 * N * 32 * (double)iters * prop.multiProcessorCount / ms * 1e-9;
      if (tops > burst) burst = tops;
      if (total_ms > 1500.f) sustained = sustained == 0.0 ? tops : 0.5 * (sustained + tops);
      total_ms += ms;
      ++reps;
    }
    printf("JSON {\"int8_dense_tops\": %.1f, \"int8_dense_tops_burst\": %.1f, \"umma_n\": %d, \"seconds\": %.2f, "
           "\"source\": \"tools/probe_umma.cu: tcgen05.mma kind::i8 (A in TMEM, B in smem), one issuing thread per SM, 148 SMs, "
           "no operand traffic; sustained = launches after 1.5 s of back-to-back load\"}\n",
           sustained, burst, N, total_ms * 1e-3);
  }
  return 0;
}
