// Developer probe: sustained rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, SS mode, no-swizzle
// planar operands like the conv kernels) as a function of N.  Ideal: N/2 cycles per MMA (each SM does 128 x N x 16).
#include <cstdio>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

__global__ void __launch_bounds__(128, 1) pair_rate_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc_pair<512>(&tmem_slot);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(256, N, 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t aoff = k * 16 + (i & 3) * 2080;
        const uint32_t boff = k * (N / 2) * 32;                       // each CTA holds N/2 rows x 2 planes
        const uint64_t ad = make_smem_desc(sa + aoff, 12480, 128, 0);
        const uint64_t bd = make_smem_desc(sb + boff, (N / 2) * 16, 128, 0);
        umma_f16_pair_warp(tmem + (k & 1) * 256, ad, bd, idesc, 1);
      }
    }
    umma_commit_pair_warp(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x / 2] = t1 - t0;
  } else if (rank == 1 && threadIdx.x == 0) {
    mbar_wait(&bar, 0);     // the multicast commit arrives here too: keep the peer (and its smem) alive
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem_slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 74 * 8);
  cudaFuncSetAttribute(pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int N : {64, 128, 192, 256}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, pair_rate_kernel, N, rep ? iters : 10, d);
      if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    long long h[74];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= 74;
    printf("cta_group::2 M=256 N=%3d planar: %.1f cycles/MMA (ideal %d) -> %.0f%% of tensor peak\n", N, avg / (iters * 8.0), N / 2,
           100.0 * (N / 2) / (avg / (iters * 8.0)));
  }
  return 0;
}
