// Caller-side consumers of the sampled range images: everything the reference does to a
// generated point cloud after `LiDARUtility.to_xyz` on the way to a picture or a BEV statistic.
//
//   render_splat_kernel + render_resolve_kernel   utils/render.py:32-80   (render_point_clouds: extrinsics, pinhole
//                                                  projection, exp(-3 depth) weights, bilinear splat, normalisation)
//   rasterize_kernel                               utils/render.py:83-142  (bilinear_rasterizer: 4-corner scatter-add)
//   surface_normal_kernel                          utils/render.py:145-234 (estimate_surface_normal, closest / mean)
//   bev_histogram_kernel                           metrics/bev.py:5-24     (point_cloud_to_histogram = torch.histogramdd
//                                                  of the xy coordinates of the points inside the depth range)
//
// All of it is HBM / atomic bound scatter-gather over 65 536 points per image: one thread per point, no staging -
// the accumulators (800 x 800 x 16 B = 10 MB per image) live in L2.  The arithmetic order follows the reference's
// fp32 expression order (explicit _rn intrinsics where an FMA contraction could move a floor / threshold decision).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

namespace r2dm {

namespace {

// utils/render.py:96-121: the four neighbouring pixels of (h, w), their bilinear weights (zero for a neighbour
// outside the image, zero below 1e-3) and flat indices w + W * h (computed in fp32 like the reference).
struct Corners {
  float wt[4];
  int idx[4];
};
__device__ __forceinline__ Corners corners_of(float h, float w, int H, int W) {
  Corners c;
  const float h_t = floorf(h), w_l = floorf(w);
  const float h_b = __fadd_rn(h_t, 1.f), w_r = __fadd_rn(w_l, 1.f);
  const float Hm = static_cast<float>(H - 1), Wm = static_cast<float>(W - 1);
  const float h_ts = fminf(fmaxf(h_t, 0.f), Hm), h_bs = fminf(fmaxf(h_b, 0.f), Hm);
  const float w_ls = fminf(fmaxf(w_l, 0.f), Wm), w_rs = fminf(fmaxf(w_r, 0.f), Wm);
  const float wh_t = __fmul_rn(__fsub_rn(h_b, h), h_t == h_ts ? 1.f : 0.f);
  const float wh_b = __fmul_rn(__fsub_rn(h, h_t), h_b == h_bs ? 1.f : 0.f);
  const float ww_l = __fmul_rn(__fsub_rn(w_r, w), w_l == w_ls ? 1.f : 0.f);
  const float ww_r = __fmul_rn(__fsub_rn(w, w_l), w_r == w_rs ? 1.f : 0.f);
  const float Wf = static_cast<float>(W);
  const float hh[4] = {h_ts, h_ts, h_bs, h_bs}, ww[4] = {w_ls, w_rs, w_ls, w_rs};
  const float a[4] = {wh_t, wh_t, wh_b, wh_b}, b[4] = {ww_l, ww_r, ww_l, ww_r};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = __fmul_rn(a[k], b[k]);
    c.wt[k] = v >= 1e-3f ? v : 0.f;
    c.idx[k] = static_cast<int>(__fadd_rn(ww[k], __fmul_rn(Wf, hh[k])));
  }
  return c;
}

__device__ __forceinline__ void add4(float4* dst, float4 v) {
  atomicAdd(dst, v);   // red.global.add.v4.f32 (sm_90+): the four accumulators of a pixel share one 16-byte unit
}

}  // namespace

// points [B][N][3] (x, y, z); colors [B][N][3] or null (= ones); R [3][3] row-major applied as p @ R, t [3]
// (either may be null); acc [B][size*size] float4 = (sum w c0, sum w c1, sum w c2, sum w), zero on entry.
__global__ void __launch_bounds__(256) render_splat_kernel(const float* __restrict__ points,
                                                           const float* __restrict__ colors,
                                                           const float* __restrict__ R, const float* __restrict__ t,
                                                           float4* __restrict__ acc, int N, int size, float focal) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = points + (static_cast<size_t>(b) * N + i) * 3;
  float x = p[0], y = p[1], z = -p[2];                              // render.py:40-41
  if (R != nullptr) {                                               // :50-52, row vector times matrix
    const float nx = __fadd_rn(__fadd_rn(__fmul_rn(x, R[0]), __fmul_rn(y, R[3])), __fmul_rn(z, R[6]));
    const float ny = __fadd_rn(__fadd_rn(__fmul_rn(x, R[1]), __fmul_rn(y, R[4])), __fmul_rn(z, R[7]));
    const float nz = __fadd_rn(__fadd_rn(__fmul_rn(x, R[2]), __fmul_rn(y, R[5])), __fmul_rn(z, R[8]));
    x = nx; y = ny; z = nz;
  }
  if (t != nullptr) { x = __fadd_rn(x, t[0]); y = __fadd_rn(y, t[1]); z = __fadd_rn(z, t[2]); }   // :53-55
  // pinhole projection (:57-66; kornia 0.7.0 project_points = convert_points_from_homogeneous, then u = x fx + cx)
  const float sc = fabsf(z) > 1e-8f ? __fdiv_rn(1.f, __fadd_rn(z, 1e-8f)) : 1.f;
  float u = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(sc, x), focal), 0.5f), static_cast<float>(size));
  float v = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(sc, y), focal), 0.5f), static_cast<float>(size));
  const float lim = static_cast<float>(size - 1);
  const bool inside = (0.f < u) && (u < lim) && (0.f < v) && (v < lim);                      // :69-70
  u = __fsub_rn(static_cast<float>(size), u);                                                 // :75
  v = __fsub_rn(static_cast<float>(size), v);
  if (!isfinite(u) || !isfinite(v)) return;   // the reference's index arithmetic is undefined here; we drop the point
  const float depth = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));   // :76
  float wgt = __fdiv_rn(1.f, expf(__fmul_rn(3.f, depth)));                                    // :77
  wgt = depth > 1e-8f ? wgt : 0.f;                                                            // :78
  float c0 = 1.f, c1 = 1.f, c2 = 1.f;
  if (colors != nullptr) {
    const float* c = colors + (static_cast<size_t>(b) * N + i) * 3;
    c0 = c[0]; c1 = c[1]; c2 = c[2];
  }
  const float m = inside ? 1.f : 0.f;                                                         // :72 (colours only)
  const float4 val = make_float4(__fmul_rn(wgt, __fmul_rn(c0, m)), __fmul_rn(wgt, __fmul_rn(c1, m)),
                                 __fmul_rn(wgt, __fmul_rn(c2, m)), wgt);
  const Corners cn = corners_of(u, v, size, size);
  float4* a = acc + static_cast<size_t>(b) * size * size;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (cn.wt[k] != 0.f && wgt != 0.f)
      add4(a + cn.idx[k], make_float4(__fmul_rn(val.x, cn.wt[k]), __fmul_rn(val.y, cn.wt[k]),
                                      __fmul_rn(val.z, cn.wt[k]), __fmul_rn(val.w, cn.wt[k])));
  }
}

// out [B][3][size][size] = colour sums / (weight sum + 1e-8)       (render.py:79-80)
__global__ void __launch_bounds__(256) render_resolve_kernel(const float4* __restrict__ acc, float* __restrict__ out,
                                                             int HW) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float4 a = acc[static_cast<size_t>(b) * HW + i];
  const float den = __fadd_rn(a.w, 1e-8f);
  float* o = out + static_cast<size_t>(b) * 3 * HW + i;
  o[0] = __fdiv_rn(a.x, den);
  o[HW] = __fdiv_rn(a.y, den);
  o[2 * HW] = __fdiv_rn(a.z, den);
}

// coords [B][N][2] = (h, w); values [B][N][C]; out [B][C][H][W], zero on entry          (render.py:83-142)
__global__ void __launch_bounds__(256) rasterize_kernel(const float* __restrict__ coords,
                                                        const float* __restrict__ values, float* __restrict__ out,
                                                        int N, int C, int H, int W) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* hw = coords + (static_cast<size_t>(b) * N + i) * 2;
  const float h = hw[0], w = hw[1];
  if (!isfinite(h) || !isfinite(w)) return;
  const Corners cn = corners_of(h, w, H, W);
  const float* val = values + (static_cast<size_t>(b) * N + i) * C;
  const size_t HW = static_cast<size_t>(H) * W;
  float* o = out + static_cast<size_t>(b) * C * HW;
  for (int c = 0; c < C; ++c) {
    const float v = val[c];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (cn.wt[k] != 0.f) atomicAdd(o + c * HW + cn.idx[k], __fmul_rn(v, cn.wt[k]));
  }
}

// points [B][3][H][W] -> unit normals [B][3][H][W]; neighbours at distance d, rows replicated at the
// elevation border, columns wrapped in azimuth (render.py:154-161); mode 0 = "closest", 1 = "mean".
__global__ void __launch_bounds__(256) surface_normal_kernel(const float* __restrict__ pts, float* __restrict__ out,
                                                             int H, int W, int d, int mode) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  const int h = i / W, w = i % W;
  const float* P = pts + static_cast<size_t>(b) * 3 * HW;
  auto at = [&](int hh, int ww, float v[3]) {
    hh = min(max(hh, 0), H - 1);
    ww = ((ww % W) + W) % W;
    const int j = hh * W + ww;
    v[0] = P[j]; v[1] = P[HW + j]; v[2] = P[2 * HW + j];
  };
  // the 8 neighbours in the reference's order (render.py:172-184), as (dh, dw) in units of d
  const int dh[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
  const int dw[8] = {0, 1, 1, 1, 0, -1, -1, -1};
  float a[3];
  at(h, w, a);
  float nb[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v[3];
    at(h + dh[k] * d, w + dw[k] * d, v);
#pragma unroll
    for (int c = 0; c < 3; ++c) nb[k][c] = __fsub_rn(v[c], a[c]);
  }
  auto cross = [&](const float* u, const float* v, float n[3]) {
    n[0] = __fsub_rn(__fmul_rn(u[1], v[2]), __fmul_rn(u[2], v[1]));
    n[1] = __fsub_rn(__fmul_rn(u[2], v[0]), __fmul_rn(u[0], v[2]));
    n[2] = __fsub_rn(__fmul_rn(u[0], v[1]), __fmul_rn(u[1], v[0]));
  };
  float n[3] = {0.f, 0.f, 0.f};
  if (mode == 0) {
    float len[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      len[k] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nb[k][0], nb[k][0]), __fmul_rn(nb[k][1], nb[k][1])),
                               __fmul_rn(nb[k][2], nb[k][2])));
    int best = 0;
    float bestv = __fadd_rn(len[0], len[2]);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float v = __fadd_rn(len[k], len[(k + 2) & 7]);
      if (v < bestv) { bestv = v; best = k; }        // first minimum, like torch.argmin
    }
    // pick without dynamic register indexing
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k == best) cross(nb[k], nb[(k + 2) & 7], n);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float c[3];
      cross(nb[k], nb[(k + 2) & 7], c);
      n[0] = __fadd_rn(n[0], c[0]); n[1] = __fadd_rn(n[1], c[1]); n[2] = __fadd_rn(n[2], c[2]);
    }
    n[0] = __fdiv_rn(n[0], 8.f); n[1] = __fdiv_rn(n[1], 8.f); n[2] = __fdiv_rn(n[2], 8.f);
  }
  const float nn = __fadd_rn(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])),
                                             __fmul_rn(n[2], n[2]))), 1e-8f);
  float* o = out + static_cast<size_t>(b) * 3 * HW + i;
  o[0] = __fdiv_rn(n[0], nn);
  o[HW] = __fdiv_rn(n[1], nn);
  o[2 * HW] = __fdiv_rn(n[2], nn);
}

// points [B][N][3]; edges [bins + 1] (the histogramdd bin edges of one axis, shared by x and y); hist [B][bins][bins]
// counts as fp32, zero on entry.  Bin i holds edges[i] <= v < edges[i+1], the last bin also v == edges[bins]
// (torch.histogramdd); points outside (min_depth, max_depth) or outside the field are dropped (metrics/bev.py:14-22).
__global__ void __launch_bounds__(256) bev_histogram_kernel(const float* __restrict__ points,
                                                            const float* __restrict__ edges,
                                                            unsigned int* __restrict__ counts, int N, int bins,
                                                            float min_depth, float max_depth) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = points + (static_cast<size_t>(b) * N + i) * 3;
  const float x = p[0], y = p[1], z = p[2];
  const float depth = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  if (!(depth > min_depth && depth < max_depth)) return;
  auto bin_of = [&](float v) {
    if (!(v >= edges[0] && v <= edges[bins])) return -1;
    int lo = 0, hi = bins + 1;                 // first edge > v  (upper bound)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (edges[mid] <= v) lo = mid + 1; else hi = mid;
    }
    const int pos = lo - 1;
    return pos == bins ? bins - 1 : pos;
  };
  const int bx = bin_of(x), by = bin_of(y);
  if (bx < 0 || by < 0) return;
  atomicAdd(counts + (static_cast<size_t>(b) * bins + bx) * bins + by, 1u);
}

__global__ void __launch_bounds__(256) counts_to_float_kernel(const unsigned int* __restrict__ counts,
                                                              float* __restrict__ hist, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) hist[i] = static_cast<float>(counts[i]);
}

// ------------------------------------------------------------------------------------ launchers
cudaError_t render_splat_launch(const float* points, const float* colors, const float* R, const float* t,
                                float* acc, float* out, int B, int N, int size, float focal, cudaStream_t s) {
  const size_t HW = static_cast<size_t>(size) * size;
  cudaError_t e = cudaMemsetAsync(acc, 0, static_cast<size_t>(B) * HW * sizeof(float4), s);
  if (e != cudaSuccess) return e;
  render_splat_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(points, colors, R, t, reinterpret_cast<float4*>(acc),
                                                              N, size, focal);
  render_resolve_kernel<<<dim3(static_cast<unsigned>((HW + 255) / 256), B), 256, 0, s>>>(
      reinterpret_cast<const float4*>(acc), out, static_cast<int>(HW));
  return cudaGetLastError();
}

cudaError_t rasterize_launch(const float* coords, const float* values, float* out, int B, int N, int C, int H, int W,
                             cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, static_cast<size_t>(B) * C * H * W * sizeof(float), s);
  if (e != cudaSuccess) return e;
  rasterize_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(coords, values, out, N, C, H, W);
  return cudaGetLastError();
}

cudaError_t surface_normal_launch(const float* points, float* out, int B, int H, int W, int d, int mode,
                                  cudaStream_t s) {
  surface_normal_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, s>>>(points, out, H, W, d, mode);
  return cudaGetLastError();
}

cudaError_t bev_histogram_launch(const float* points, const float* edges, unsigned int* counts, float* hist, int B,
                                 int N, int bins, float min_depth, float max_depth, cudaStream_t s) {
  const size_t n = static_cast<size_t>(B) * bins * bins;
  cudaError_t e = cudaMemsetAsync(counts, 0, n * sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  bev_histogram_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(points, edges, counts, N, bins, min_depth, max_depth);
  counts_to_float_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(counts, hist, n);
  return cudaGetLastError();
}

}  // namespace r2dm
