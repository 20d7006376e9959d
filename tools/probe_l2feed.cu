// Developer probe: how fast can all 148 SMs pull data from L2 into shared memory with bulk copies?
// The conv kernel re-fetches its weight stage (same bytes for every CTA) plus a distinct activation
// tile per stage; this probe measures delivered GB/s for
//   mode 0: every CTA loads the SAME 36 KB block        (weights-like)
//   mode 1: every CTA loads DISTINCT blocks from a 64 MB buffer (activation-like, L2 resident)
//   mode 2: cluster of 2, each CTA loads half of the same block and multicasts it to both
//   mode 3: cluster of 2, each CTA loads its own half only (what cta_group::2 needs)
//   mode 4: cluster of 4, each CTA loads a quarter and multicasts to all four
#include <cstdio>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

constexpr int kSlots = 4;
constexpr uint32_t kBlock = 36864;

__device__ __forceinline__ void bulk_load_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                             uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128, 1) feed_kernel(const uint8_t* src, size_t src_bytes, int mode, int iters, int csz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full[kSlots];
  const uint32_t rank = csz > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
  if (threadIdx.x == 0) {
    const size_t nblocks = src_bytes / kBlock;
    for (int i = 0; i < iters + kSlots; ++i) {
      const int s = i % kSlots;
      if (i >= kSlots) mbar_wait(&full[s], ((i / kSlots) - 1) & 1);
      // NOTE: with multicast the peer may still be "using" its slot; a real kernel needs an empty
      // barrier across the cluster.  The probe only measures delivery rate, the data is never read.
      if (i < iters) {
        uint8_t* dst = smem + static_cast<size_t>(s) * kBlock;
        if (mode == 0) {
          mbar_expect_tx(&full[s], kBlock);
          bulk_load(dst, src, kBlock, &full[s]);
        } else if (mode == 1) {
          const size_t blk = (static_cast<size_t>(blockIdx.x) * 977 + static_cast<size_t>(i) * 148) % nblocks;
          mbar_expect_tx(&full[s], kBlock);
          bulk_load(dst, src + blk * kBlock, kBlock, &full[s]);
        } else if (mode == 2 || mode == 4) {
          const uint32_t part = kBlock / csz;
          mbar_expect_tx(&full[s], kBlock);
          bulk_load_mc(dst + rank * part, src + rank * part, part, &full[s], static_cast<uint16_t>((1u << csz) - 1));
        } else if (mode == 3) {
          const uint32_t part = kBlock / csz;
          mbar_expect_tx(&full[s], part);
          bulk_load(dst + rank * part, src + rank * part, part, &full[s]);
        }
      }
    }
  }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
}

int main() {
  uint8_t* src;
  const size_t bytes = 64ull << 20;
  cudaMalloc(&src, bytes);
  cudaMemset(src, 1, bytes);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 4000;
  struct { int mode, csz; const char* name; double delivered_frac; } cases[] = {
      {0, 1, "same block, unicast", 1.0},
      {1, 1, "distinct blocks (L2 resident 64 MB)", 1.0},
      {2, 2, "same block, cluster 2 multicast halves", 1.0},
      {3, 2, "same block, each CTA own half only", 0.5},
      {4, 4, "same block, cluster 4 multicast quarters", 1.0},
  };
  for (auto& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = c.csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchKernelEx(&cfg, feed_kernel, (const uint8_t*)src, bytes, c.mode, rep ? iters : 50, c.csz);
      cudaEventRecord(e1);
      if (err != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("%s: error %s\n", c.name, cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) {
        const double delivered = 148.0 * iters * kBlock * c.delivered_frac;
        printf("%-45s %7.3f ms  delivered into smem %7.1f GB/s total, %6.1f GB/s per SM\n", c.name, ms,
               delivered / ms / 1e6, delivered / ms / 1e6 / 148);
      }
    }
  }
  return 0;
}
