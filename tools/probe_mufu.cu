// Developer probe: special-function (XU pipe) throughput on B200 for the activation inside the conv
// operand transform.  Prints ops/clk/SM for tanh.approx.f32, ex2.approx.f32, rcp.approx.f32,
// tanh.approx.bf16x2, ex2.approx.ftz.bf16x2 and an FMA-pipe polynomial.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  unsigned ua = __float_as_uint(a), ub = __float_as_uint(b), uc = ua ^ 0x1234, ud = ub ^ 0x4321;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 2) {
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 3) {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ud));
    } else if (MODE == 4) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ua)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(uc)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ud));
    } else if (MODE == 5) {  // 8 dependent FMAs per value (polynomial), 4 values
#pragma unroll
      for (int j = 0; j < 8; ++j) { a = fmaf(a, 0.999f, 0.001f); b = fmaf(b, 0.999f, 0.001f); c = fmaf(c, 0.999f, 0.001f); d = fmaf(d, 0.999f, 0.001f); }
    } else if (MODE == 6) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ud));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ua ^ ub ^ uc ^ ud);
}

template <int MODE>
void run(const char* name, int per_iter) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = double(148) * 8 * 256 * iters * per_iter;
  printf("%-28s %8.3f ms  %7.2f Gops/s  ~%5.2f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 4);
  run<1>("ex2.approx.ftz.f32", 4);
  run<2>("rcp.approx.ftz.f32", 4);
  run<3>("tanh.approx.bf16x2 (instr)", 4);
  run<6>("tanh.approx.f16x2 (instr)", 4);
  run<4>("ex2.approx.ftz.bf16x2 (instr)", 4);
  run<5>("fma (8 per value)", 32);
  return 0;
}
