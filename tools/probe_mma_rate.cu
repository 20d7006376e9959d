// Developer probe: sustained tcgen05.mma issue/execute rate (cycles per MMA) for M=128, SS mode,
// as a function of N and of the shared-memory operand layout (no-swizzle planar vs 128B swizzle).
#include <cstdio>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

// layout: 0 = no swizzle (LBO = plane stride, SBO = 128), 2 = SW128 (SBO = 1024)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int layout, int iters, int same_d, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(128, N, 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // walk through different A offsets like the 9 taps do
        const uint32_t aoff = (layout == 0) ? (k * 16 + (i & 3) * 2080) : (k & 3) * 32 + ((i & 3) * 8 + (k >> 2)) * 128;
        const uint32_t boff = (layout == 0) ? k * 2 * N * 16 : (k & 3) * 32 + (k >> 2) * N * 128;
        const uint64_t ad = layout == 0 ? make_smem_desc(sa + aoff, 12480, 128, 0) : make_smem_desc(sa + aoff, 16, 1024, 2);
        const uint64_t bd = layout == 0 ? make_smem_desc(sb + boff, N * 16, 128, 0) : make_smem_desc(sb + boff, 16, 1024, 2);
        umma_f16(tmem + (same_d ? 0 : (k & 1) * 256), ad, bd, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int layout : {0, 2})
    for (int same_d : {1, 0})
      for (int N : {64, 128, 192, 256}) {
        rate_kernel<<<148, 128, 200 * 1024>>>(N, layout, 10, same_d, d);
        rate_kernel<<<148, 128, 200 * 1024>>>(N, layout, iters, same_d, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= 148;
        printf("layout=%s sameD=%d N=%3d: %.1f cycles/MMA (ideal %d)  -> %.0f%% of tensor peak\n",
               layout ? "sw128" : "planar", same_d, N, avg / (iters * 8.0), N / 2, 100.0 * (N / 2) / (avg / (iters * 8.0)));
      }
  return 0;
}
