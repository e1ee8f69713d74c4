// Hardware probe (run under gpurun; not part of the product): sustained FP64 tensor rate of
// mma.sync.m8n8k4.f64 (DMMA), the instruction behind the FP64 route, the orthonormalisation and the
// LD tiles. 8 warps per CTA x 4 CTAs per SM, 8 independent accumulator pairs per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/probe_dmma tools/probe_dmma.cu
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_dmma(int iters, double* sink) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  double* sink;
  CK(cudaMalloc(&sink, sizeof(double) * 1024 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = prop.multiProcessorCount * 4, iters = 1 << 16;
  double burst = 0.0, sustained = 0.0;
  float total = 0.f;
  while (total < 3000.f) {
    CK(cudaEventRecord(e0));
    k_dmma<<<grid, 256>>>(iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp instruction
    const double tf = 512.0 * 8 * (double)iters * 8 /*warps*/ * grid / ms * 1e-9;
    if (tf > burst) burst = tf;
    if (total > 1500.f) sustained = sustained == 0.0 ? tf : 0.5 * (sustained + tf);
    total += ms;
  }
  printf("JSON {\"fp64_dmma_tflops\": %.2f, \"fp64_dmma_tflops_burst\": %.2f, \"seconds\": %.2f, "
         "\"source\": \"tools/probe_dmma.cu: mma.sync.m8n8k4.f64 from registers, 32 warps per SM, 8 independent accumulators per warp\"}\n",
         sustained, burst, total * 1e-3);
  return 0;
}
