// Developer probe: what does one launch of a persistent, full-shared-memory kernel cost when launches are
// chained back to back (CUDA graph of 200 nodes)?  Varies block size, dynamic shared memory, TMEM allocation,
// mbarrier init and programmatic dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"
using namespace r2dm;

__global__ void k(int tmem, int pdl, float* out) {
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bars[32];
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  if (tmem && threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (pdl) { pdl_launch_dependents(); pdl_wait(); }
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
  __syncthreads();
  if (tmem && threadIdx.x < 32) tmem_dealloc<512>(slot);
}

static float run(int threads, int smem, int tmem, int pdl) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaStream_t s; cudaStreamCreate(&s);
  const int n = 200;
  auto launch_all = [&]() {
    for (int i = 0; i < n; ++i) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, k, tmem, pdl, (float*)nullptr);
    }
  };
  launch_all(); cudaStreamSynchronize(s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  launch_all();
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (5 * n);
}

int main() {
  printf("threads smemKB tmem pdl : us per launch (graph of 200 chained launches, 148 CTAs)\n");
  for (int pdl : {0, 1})
    for (int tmem : {0, 1})
      for (int smem : {0, 100 * 1024, 214 * 1024})
        for (int threads : {128, 640})
          printf("%5d %6d %4d %3d : %.2f\n", threads, smem / 1024, tmem, pdl, run(threads, smem, tmem, pdl));
  return 0;
}
