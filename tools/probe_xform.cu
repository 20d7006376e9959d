// Developer probe: cycles per 16-byte unit of the conv operand transform (silu(a*x+d) on 8 bf16)
// for different instruction mixes, 4 warps per CTA (one per SM sub-partition) like the conv kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../r2dm_b200/csrc/common.cuh"
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

__device__ __forceinline__ float tanh_f32(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) { uint32_t y; asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t fma_bf16x2(uint32_t a, uint32_t b, uint32_t c) { uint32_t y; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(y) : "r"(a), "r"(b), "r"(c)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out, float a0, float d0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  for (int i = threadIdx.x; i < 24960 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3f003e80u;
  __syncthreads();
  float ca[8], cd[8];
  for (int i = 0; i < 8; ++i) { ca[i] = a0 + i * 1e-3f; cd[i] = d0 - i * 1e-3f; }
  uint32_t pa[4], pd[4];
  for (int i = 0; i < 4; ++i) { pa[i] = pack_bf16x2(ca[2 * i], ca[2 * i + 1]); pd[i] = pack_bf16x2(cd[2 * i], cd[2 * i + 1]); }
  const int tip = threadIdx.x & 63, plane = threadIdx.x >> 6;
  const int n_units = 780;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int i0 = tip; i0 < n_units; i0 += 4 * 64) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * 64 < n_units) raw[u] = lds128(sbase + plane * 12480 + (i0 + u * 64) * 16);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (MODE == 0) {          // fp32: unpack, ffma, tanh.f32, ffma, cvt pack
          float v[8];
          Elem<__nv_bfloat16>::unpack(raw[u], v);
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float h = fmaf(v[q], ca[q], cd[q]); v[q] = fmaf(h, tanh_f32(h), h); }
          raw[u] = Elem<__nv_bfloat16>::pack(v);
        } else if (MODE == 1) {   // all bf16x2: hfma2, tanh.bf16x2, hfma2
          uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { const uint32_t h = fma_bf16x2(w[q], pa[q], pd[q]); w[q] = fma_bf16x2(h, tanh_bf16x2(h), h); }
          raw[u] = make_uint4(w[0], w[1], w[2], w[3]);
        } else if (MODE == 2) {   // fp32 affine, pack, tanh.bf16x2, hfma2
          float v[8];
          Elem<__nv_bfloat16>::unpack(raw[u], v);
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t h = pack_bf16x2(fmaf(v[2 * q], ca[2 * q], cd[2 * q]), fmaf(v[2 * q + 1], ca[2 * q + 1], cd[2 * q + 1]));
            w[q] = fma_bf16x2(h, tanh_bf16x2(h), h);
          }
          raw[u] = make_uint4(w[0], w[1], w[2], w[3]);
        } else if (MODE == 3) {   // fp32 affine only (no activation), cvt pack
          float v[8];
          Elem<__nv_bfloat16>::unpack(raw[u], v);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = fmaf(v[q], ca[q], cd[q]);
          raw[u] = Elem<__nv_bfloat16>::pack(v);
        } else if (MODE == 4) {   // copy only
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * 64 < n_units) sts128(sbase + plane * 12480 + (i0 + u * 64) * 16, raw[u]);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int iters = 200;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  k<MODE><<<148, 128, 32 * 1024>>>(2, d, 0.7f, 0.1f);
  k<MODE><<<148, 128, 32 * 1024>>>(iters, d, 0.7f, 0.1f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  // per stage: 2 planes x 780 units over 128 threads = 12.2 units per thread
  printf("%-44s %7.0f cycles per stage (1560 units)  = %.1f cycles/unit/warp\n", name, avg / iters, avg / iters / 12.19);
  cudaFree(d);
}
int main() {
  run<0>("fp32: unpack ffma tanh.f32 ffma cvt");
  run<1>("bf16x2: hfma2 tanh.bf16x2 hfma2");
  run<2>("fp32 affine, cvt, tanh.bf16x2, hfma2");
  run<3>("fp32 affine only + cvt");
  run<4>("copy only");
  return 0;
}
