// Developer probe: does the packed fp32 arithmetic of sm_100 (fma.rn.f32x2 / add.f32x2) raise the FMA-pipe
// throughput, or only save issue slots?  And do MUFU.TANH and FFMA overlap?  Prints fp32 results/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + 0.1f * i;
  unsigned long long p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  unsigned long long m, c;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(0.999f));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(0.001f));
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 8 chains x 4 scalar FFMA = 32 results
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, 0.001f);
    } else if (MODE == 1) {   // 4 pair-chains x 4 FFMA2 = 32 results
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(c));
    } else if (MODE == 2) {   // add.f32x2
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c));
    } else if (MODE == 3) {   // 4 tanh + 28 FFMA (transform-like mix)
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
      for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 4; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, 0.001f);
    } else if (MODE == 4) {   // 4 tanh + 14 FFMA2 (= 28 results)
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
      for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 2; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(c));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_iter, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 148 * 2048 * 4);
  const int iters = 8192;
  const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
  k<MODE><<<blocks, threads>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = double(blocks) * threads * iters * per_iter;
  printf("%-34s warps/SM %2d  %8.3f ms  ~%6.2f results/clk/SM (at %d MHz nominal)\n", name, warps_per_sm, ms,
         ops / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    if (w == 8) {
      run<0>("FFMA scalar", 32, 8); run<1>("fma.rn.f32x2", 32, 8); run<2>("add.rn.f32x2", 32, 8);
      run<3>("4 tanh + 28 FFMA (results: 32)", 32, 8); run<4>("4 tanh + 14 FFMA2 (results: 32)", 32, 8);
    } else if (w == 16) {
      run<0>("FFMA scalar", 32, 16); run<1>("fma.rn.f32x2", 32, 16); run<2>("add.rn.f32x2", 32, 16);
      run<3>("4 tanh + 28 FFMA (results: 32)", 32, 16); run<4>("4 tanh + 14 FFMA2 (results: 32)", 32, 16);
    } else {
      run<0>("FFMA scalar", 32, 32); run<1>("fma.rn.f32x2", 32, 32); run<2>("add.rn.f32x2", 32, 32);
      run<3>("4 tanh + 28 FFMA (results: 32)", 32, 32); run<4>("4 tanh + 14 FFMA2 (results: 32)", 32, 32);
    }
  }
  return 0;
}
