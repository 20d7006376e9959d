// Small kernels of the PointNet feature extractor (metrics/extractor/pointnet.py of the reference; SURVEY
// section 8 f-4).  The three point-wise layers of each trunk (Conv1d k=1 + BatchNorm1d (eval) + ReLU, :10-12,
// :39-41) run as 1x1 convolutions of conv_umma_kernel over the range image (a LiDAR point cloud IS the
// 64 x 1024 image, evaluate.py:114), the last one with the global max pool (x.amax(dim=2), :27,:57) fused into its
// epilogue; what is left for this file is O(B * 1024) work: the dense layers, the 3x3 input transform and fills.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace r2dm {

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, float v, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

// out[b][n] = act(sum_k in[b][k] w[n][k] + bias[n]); one warp per output (nn.Linear + folded BatchNorm1d + ReLU,
// pointnet.py:28-30, 76-78)
__global__ void __launch_bounds__(256) dense_kernel(const float* __restrict__ in, int in_stride,
                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                    float* __restrict__ out, int out_stride, int K, int N, int relu) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* x = in + static_cast<size_t>(b) * in_stride;
  const float* wr = w + static_cast<size_t>(n) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(x[k], wr[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += bias ? bias[n] : 0.f;
    out[static_cast<size_t>(b) * out_stride + n] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

// y[b][j][n] = sum_i x[b][i][n] trans[b][i][j]   (torch.bmm(x^T, trans)^T, pointnet.py:49-52)
__global__ void __launch_bounds__(256) point_transform_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ trans, float* __restrict__ y,
                                                              int N) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* t = trans + b * 9;
  const float* xb = x + static_cast<size_t>(b) * 3 * N;
  const float x0 = xb[n], x1 = xb[N + n], x2 = xb[2 * N + n];
  float* yb = y + static_cast<size_t>(b) * 3 * N;
#pragma unroll
  for (int j = 0; j < 3; ++j) yb[j * N + n] = fmaf(x2, t[6 + j], fmaf(x1, t[3 + j], x0 * t[j]));
}

cudaError_t fill_launch(float* p, float v, size_t n, cudaStream_t s) {
  fill_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, v, n);
  return cudaGetLastError();
}
cudaError_t dense_launch(const float* in, int in_stride, const float* w, const float* bias, float* out, int out_stride,
                         int B, int K, int N, int relu, cudaStream_t s) {
  dense_kernel<<<dim3((N + 7) / 8, B), 256, 0, s>>>(in, in_stride, w, bias, out, out_stride, K, N, relu);
  return cudaGetLastError();
}
cudaError_t point_transform_launch(const float* x, const float* trans, float* y, int B, int N, cudaStream_t s) {
  point_transform_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(x, trans, y, N);
  return cudaGetLastError();
}

}  // namespace r2dm
