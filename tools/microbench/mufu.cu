// MUFU throughput probe on B200: tanh.approx.f32 vs tanh.approx.f16x2 vs tanh.approx.bf16x2 vs ex2.approx.f32
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  unsigned ua = __float_as_uint(a) & 0x3fff3fffu, ub = ua + 1, uc = ua + 2, ud = ua + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 1) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ud));
    } else if (MODE == 2) {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ud));
    } else if (MODE == 3) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else {
      a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 1.0001f, 0.5f); c = fmaf(c, 1.0001f, 0.5f); d = fmaf(d, 1.0001f, 0.5f);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ua ^ ub ^ uc ^ ud);
}
template <int MODE>
void run(const char* name, int elems_per_op) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 512>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 2, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 2 * 512 * iters * 4;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %.3f ms  %.1f instr-lanes/clk/SM  (%.1f elements/clk/SM) at %d MHz nominal\n", name, ms,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), ops * elems_per_op / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("tanh.approx.f32", 1); run<1>("tanh.approx.f16x2", 2); run<2>("tanh.approx.bf16x2", 2);
  run<3>("ex2.approx.ftz.f32", 1); run<4>("ffma", 1);
  return 0;
}
