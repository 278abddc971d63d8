// microbenchmark: per-SM global store throughput by store width, in the access shape of k_cell_bwd_f's E0
// (16 warps per CTA, one CTA per SM, every warp writes 2 KB regions):  nvcc -arch=sm_100a -O3 -o tools/_bin/store_bw tools/store_bw.cu
//   mode 0: 16 x STG.32  per 2 KB (lane = row: one 128-byte line per instruction)      -- the transposed-tile layout
//   mode 1:  8 x STG.64  per 2 KB
//   mode 2:  4 x STG.128 per 2 KB (lane writes 16 bytes: four lines per instruction)   -- an interleaved-4 layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_store(float* out, int iters, int puts) {
  extern __shared__ uint8_t hog[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = out + (size_t)blockIdx.x * iters * puts * 16 * 512;   // per CTA: iters x puts x 16 warps x 2 KB
  float v = (float)threadIdx.x;
  for (int it = 0; it < iters; ++it)
    for (int p = 0; p < puts; ++p) {
      float* reg = base + ((size_t)(it * puts + p) * 16 + warp) * 512;
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) reg[i * 32 + lane] = v + i;
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) reinterpret_cast<float2*>(reg)[i * 32 + lane] = make_float2(v + i, v);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(reg)[i * 32 + lane] = make_float4(v + i, v, v, v);
      }
    }
  if (hog[0] == 123) out[0] = 1.f;
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 64, puts = 8;                        // per CTA 64 x 8 x 32 KB = 16 MB
  const size_t bytes = (size_t)sms * iters * puts * 16 * 2048;
  float* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      auto launch = [&]() {
        const size_t sm = 200 * 1024;
        if (mode == 0) { cudaFuncSetAttribute(k_store<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k_store<0><<<sms, 512, sm>>>(buf, iters, puts); }
        if (mode == 1) { cudaFuncSetAttribute(k_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k_store<1><<<sms, 512, sm>>>(buf, iters, puts); }
        if (mode == 2) { cudaFuncSetAttribute(k_store<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k_store<2><<<sms, 512, sm>>>(buf, iters, puts); }
      };
      cudaEventRecord(a);
      launch();
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    printf("mode %d: %.3f ms, %.0f GB/s chip, %.1f GB/s per SM (%s)\n", mode, best, bytes / best / 1e6, bytes / best / 1e6 / sms,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
