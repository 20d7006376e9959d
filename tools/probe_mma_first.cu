// Developer probe: why is the first pipeline stage of MMAs in a freshly launched kernel slow
// (1.7-2.6 us vs 0.7 us)?  Times successive groups of 8 MMAs (+commit+wait) issued from the same loop
// code in a cold kernel, optionally after one dummy MMA issued from a different code address:
// if the dummy removes the first-group penalty it is a tensor-pipe wake-up, otherwise instruction fetch.
#include <cstdio>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

__global__ void __launch_bounds__(128, 1) first_kernel(int dummy, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(128, 128, 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 96 * 1024);
    uint32_t phase = 0;
    long long t[6];
    t[0] = clock64();
    if (dummy) {
      umma_f16(tmem + 384, make_smem_desc(sa, 2080, 128, 0), make_smem_desc(sb, 2048, 128, 0), idesc, 0);
      umma_commit(&bar);
      mbar_wait(&bar, phase); phase ^= 1;
    }
    t[1] = clock64();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t ad = make_smem_desc(sa + k * 16, 12480, 128, 0);
        const uint64_t bd = make_smem_desc(sb + k * 4096, 2048, 128, 0);
        umma_f16(tmem + (k & 1) * 128, ad, bd, idesc, 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase); phase ^= 1;
      t[2 + g] = clock64();
    }
    if (blockIdx.x == 0) for (int i = 0; i < 6; ++i) out[i] = t[i] - t[0];
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 3; ++rep)
    for (int dummy : {0, 1}) {
      first_kernel<<<148, 128, 200 * 1024>>>(dummy, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      long long h[6];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("rep %d dummy=%d: dummy MMA %lld cyc | groups of 8 MMAs (ideal 512 cyc): %lld %lld %lld %lld\n", rep, dummy,
             h[1], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4]);
    }
  return 0;
}
