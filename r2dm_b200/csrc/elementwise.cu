// HBM-bound kernels of the sampling path: layout conversion, GroupNorm/AdaGN(+SiLU) apply,
// closed-form FIR resampling, the conditioning (time-embedding / FiLM) table and the fused
// sampler update.  All activation traffic is 16-byte vectorised along the 1024-wide azimuth axis.
#include <cstring>
#include <curand_kernel.h>
#include "common.cuh"
#include "kernels.h"

namespace r2dm {

// =============================================================================== layout conversion
// fp32 NCHW [B][Csrc][H][W] -> channels [c_off, c_off+Csrc) of a planar-16 tensor (zero elsewhere
// inside the touched planes is NOT written: planes are written whole, missing channels = 0).
template <typename T>
__global__ void pack_nchw_kernel(const float* __restrict__ src, int Csrc, uint4* __restrict__ dst, int planes,
                                 int H, int W, int plane_begin, int nplanes, int c_off) {
  constexpr int CW = Elem<T>::CW;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y % H, pl = plane_begin + blockIdx.y / H;
  const int b = blockIdx.z;
  if (x >= W || pl >= plane_begin + nplanes) return;
  float v[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const int c = pl * CW + i - c_off;
    v[i] = (c >= 0 && c < Csrc) ? src[((static_cast<size_t>(b) * Csrc + c) * H + y) * W + x] : 0.f;
  }
  const uint4 pk = Elem<T>::pack_mma(v);
  const size_t idx = pt_index(b, planes, pl, H, W + 2, y, x + 1);
  dst[idx] = pk;
  if (x == 0) dst[idx + W] = pk;
  if (x == W - 1) dst[idx - W] = pk;
}

cudaError_t pack_nchw(int dtype, const float* src, int B, int Csrc, int H, int W, PT dst, int c_off,
                      cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  const int pb = c_off / cw, pe = (c_off + Csrc + cw - 1) / cw;
  dim3 grid((W + 127) / 128, H * (pe - pb), B);
  if (dtype == kBF16)
    pack_nchw_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(src, Csrc, static_cast<uint4*>(dst.ptr), dst.C / cw, H, W,
                                                         pb, pe - pb, c_off);
  else
    pack_nchw_kernel<float><<<grid, 128, 0, s>>>(src, Csrc, static_cast<uint4*>(dst.ptr), dst.C / cw, H, W, pb,
                                                 pe - pb, c_off);
  return cudaGetLastError();
}

template <typename T>
__global__ void unpack_nchw_kernel(const uint4* __restrict__ src, int planes, float* __restrict__ dst, int Cdst,
                                   int H, int W, int c_off) {
  constexpr int CW = Elem<T>::CW;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y % H, pl = c_off / CW + blockIdx.y / H;
  const int b = blockIdx.z;
  if (x >= W) return;
  float v[CW];
  Elem<T>::unpack(src[pt_index(b, planes, pl, H, W + 2, y, x + 1)], v);
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const int c = pl * CW + i - c_off;
    if (c >= 0 && c < Cdst) dst[((static_cast<size_t>(b) * Cdst + c) * H + y) * W + x] = v[i];
  }
}

cudaError_t unpack_nchw(int dtype, PT src, float* dst, int c_off, int Cdst, cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  const int pb = c_off / cw, pe = (c_off + Cdst + cw - 1) / cw;
  dim3 grid((src.W + 127) / 128, src.H * (pe - pb), src.B);
  if (dtype == kBF16)
    unpack_nchw_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const uint4*>(src.ptr), src.C / cw, dst, Cdst,
                                                           src.H, src.W, c_off);
  else
    unpack_nchw_kernel<float><<<grid, 128, 0, s>>>(static_cast<const uint4*>(src.ptr), src.C / cw, dst, Cdst, src.H,
                                                   src.W, c_off);
  return cudaGetLastError();
}

// network input: channels [x (Cx) | coords encoding (Ce) | zero padding]; efficient_unet.py:278-281
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ xin, int Cx, const float* __restrict__ enc, int Ce,
                                  uint4* __restrict__ dst, int planes, int H, int W, int plane_begin) {
  pdl_launch_dependents();   // PDL: let the next kernel become resident; wait for the previous one's results
  pdl_wait();
  constexpr int CW = Elem<T>::CW;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y % H, pl = plane_begin + blockIdx.y / H;
  const int b = blockIdx.z;
  if (x >= W) return;
  float v[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const int c = pl * CW + i;
    float t = 0.f;
    if (c < Cx) t = xin ? xin[((static_cast<size_t>(b) * Cx + c) * H + y) * W + x] : 0.f;
    else if (c < Cx + Ce) t = enc[(static_cast<size_t>(c - Cx) * H + y) * W + x];
    v[i] = t;
  }
  const uint4 pk = Elem<T>::pack_mma(v);
  const size_t idx = pt_index(b, planes, pl, H, W + 2, y, x + 1);
  dst[idx] = pk;
  if (x == 0) dst[idx + W] = pk;
  if (x == W - 1) dst[idx - W] = pk;
}

cudaError_t pack_input(int dtype, const float* x, int Cx, const float* enc, int Ce, PT dst, int plane_begin,
                       int plane_end, cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  dim3 grid((dst.W + 127) / 128, dst.H * (plane_end - plane_begin), dst.B);
  if (dtype == kBF16)
    return launch_pdl(pack_input_kernel<__nv_bfloat16>, grid, dim3(128), 0, s, x, Cx, enc, Ce, static_cast<uint4*>(dst.ptr),
                      dst.C / cw, dst.H, dst.W, plane_begin);
  return launch_pdl(pack_input_kernel<float>, grid, dim3(128), 0, s, x, Cx, enc, Ce, static_cast<uint4*>(dst.ptr),
                    dst.C / cw, dst.H, dst.W, plane_begin);
}

// =============================================================================== statistics
// One block per (row-block, plane, b): partial (sum, sumsq) of a plane slab; slot layout
// [plane_in_unit][row_block].
constexpr int kStatRows = 4;
template <typename T>
__global__ void tensor_stats_kernel(const uint4* __restrict__ src, int planes, int H, int W, float* __restrict__ stats,
                                    int slots, int planes_per_unit) {
  constexpr int CW = Elem<T>::CW;
  const int rb = blockIdx.x, pl = blockIdx.y, b = blockIdx.z;
  float s1 = 0.f, s2 = 0.f;
  for (int r = 0; r < kStatRows; ++r) {
    const int y = rb * kStatRows + r;
    if (y >= H) break;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      float v[CW];
      Elem<T>::unpack(src[pt_index(b, planes, pl, H, W + 2, y, x + 1)], v);
#pragma unroll
      for (int i = 0; i < CW; ++i) { s1 += v[i]; s2 += v[i] * v[i]; }
    }
  }
  __shared__ float red[2][8];
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    const int unit = pl / planes_per_unit, slot = (pl % planes_per_unit) * gridDim.x + rb;
    stats[((static_cast<size_t>(b) * kNU + unit) * slots + slot) * 2 + threadIdx.x] = t;
  }
}

int tensor_stats_slots(int dtype, const PT& t) {
  const int ppu = t.C / kNU / dtype_cw(dtype);
  return ppu * ((t.H + kStatRows - 1) / kStatRows);
}

cudaError_t tensor_stats_launch(int dtype, PT t, cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  const int ppu = t.C / kNU / cw;
  dim3 grid((t.H + kStatRows - 1) / kStatRows, t.C / cw, t.B);
  if (dtype == kBF16)
    tensor_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const uint4*>(t.ptr), t.C / cw, t.H, t.W,
                                                            t.stats, t.slots, ppu);
  else
    tensor_stats_kernel<float><<<grid, 256, 0, s>>>(static_cast<const uint4*>(t.ptr), t.C / cw, t.H, t.W, t.stats,
                                                    t.slots, ppu);
  return cudaGetLastError();
}

// =============================================================================== GN / AdaGN apply
// nn.GroupNorm (efficient_unet.py:33,72) and ops.AdaGN (ops.py:176-200) + nn.SiLU, reading the
// partial statistics left by the producer kernel.  The source may be a channel concat of two
// tensors (efficient_unet.py:290-292), which this kernel materialises for free.
struct GnParams {
  const uint4* src[2];
  const float* stats[2];
  int C[2], slots[2];
  uint4* dst;
  int B, H, W, Ctot, groups;
  float eps;
  const float* gamma; const float* beta;
  const float* film; int film_stride, film_off;
  const int* step_ptr; int row_batch_stride, rows_per_step;
  int silu;
  int rows_per_block;
};

template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnParams p) {
  constexpr int CW = Elem<T>::CW;
  const int pl = blockIdx.y, b = blockIdx.z;
  __shared__ float coef[2][CW];
  __shared__ double red[2][8];
  const int c0 = pl * CW;
  const int gsize = p.Ctot / p.groups;
  const int g = c0 / gsize;
  // ---- group statistics from the producers' partial sums (fp64 combine)
  {
    double s1 = 0.0, s2 = 0.0;
    const int lo = g * gsize, hi = lo + gsize;
    int off = 0;
    for (int si = 0; si < 2; ++si) {
      if (p.src[si] == nullptr) break;
      const int Cs = p.C[si];
      const int a = max(lo, off), e = min(hi, off + Cs);
      if (a < e) {
        const int unit_ch = Cs / kNU;
        const int u0 = (a - off) / unit_ch, u1 = (e - off) / unit_ch;
        const int n = (u1 - u0) * p.slots[si];
        const float* st = p.stats[si] + (static_cast<size_t>(b) * kNU + u0) * p.slots[si] * 2;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { s1 += st[2 * i]; s2 += st[2 * i + 1]; }
      }
      off += Cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < CW) {
      double t1 = 0.0, t2 = 0.0;
      for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
      const double cnt = static_cast<double>(gsize) * p.H * p.W;
      const double mean = t1 / cnt;
      double var = t2 / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
      const int c = c0 + threadIdx.x;
      float ga, be;
      if (p.film != nullptr) {
        const int row = (p.step_ptr ? *p.step_ptr : 0) * p.rows_per_step + b * p.row_batch_stride;
        const float* f = p.film + static_cast<size_t>(row) * p.film_stride + p.film_off;
        ga = 1.f + f[c]; be = f[p.Ctot + c];
      } else {
        ga = p.gamma[c]; be = p.beta[c];
      }
      const float a = rstd * ga;
      coef[0][threadIdx.x] = a;
      coef[1][threadIdx.x] = be - static_cast<float>(mean) * a;
    }
    __syncthreads();
  }
  float ca[CW], cb[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) { ca[i] = coef[0][i]; cb[i] = coef[1][i]; }
  // ---- source plane
  const int planes0 = p.C[0] / CW;
  const uint4* sp;
  int sp_planes, sp_pl;
  if (pl < planes0) { sp = p.src[0]; sp_planes = planes0; sp_pl = pl; }
  else { sp = p.src[1]; sp_planes = p.C[1] / CW; sp_pl = pl - planes0; }
  const int planes_dst = p.Ctot / CW;
  const int Wp = p.W + 2;
  const int y_begin = blockIdx.x * p.rows_per_block;
  const int n = p.rows_per_block * p.W;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = y_begin + i / p.W, x = i % p.W;
    if (y >= p.H) break;
    float v[CW];
    Elem<T>::unpack(sp[pt_index(b, sp_planes, sp_pl, p.H, Wp, y, x + 1)], v);
#pragma unroll
    for (int k = 0; k < CW; ++k) {
      const float t = fmaf(v[k], ca[k], cb[k]);
      v[k] = p.silu ? silu_f(t) : t;
    }
    const uint4 pk = Elem<T>::pack_mma(v);
    const size_t idx = pt_index(b, planes_dst, pl, p.H, Wp, y, x + 1);
    p.dst[idx] = pk;
    if (x == 0) p.dst[idx + p.W] = pk;
    if (x == p.W - 1) p.dst[idx - p.W] = pk;
  }
}

cudaError_t gn_apply_launch(const GnApply& g, cudaStream_t s) {
  GnParams p;
  p.src[0] = static_cast<const uint4*>(g.src0.ptr); p.src[1] = static_cast<const uint4*>(g.src1.ptr);
  p.stats[0] = g.src0.stats; p.stats[1] = g.src1.stats;
  p.C[0] = g.src0.C; p.C[1] = g.src1.ptr ? g.src1.C : 0;
  p.slots[0] = g.src0.slots; p.slots[1] = g.src1.slots;
  p.dst = static_cast<uint4*>(g.dst.ptr);
  p.B = g.dst.B; p.H = g.dst.H; p.W = g.dst.W; p.Ctot = g.dst.C; p.groups = g.groups; p.eps = g.eps;
  p.gamma = g.gamma; p.beta = g.beta; p.film = g.film; p.film_stride = g.film_stride; p.film_off = g.film_off;
  p.step_ptr = g.step_ptr; p.row_batch_stride = g.row_batch_stride; p.rows_per_step = g.rows_per_step;
  p.silu = g.silu;
  const int cw = dtype_cw(g.dtype);
  // ~2048 pixels per block keeps >= 2 waves at every level
  int rpb = 2048 / p.W; if (rpb < 1) rpb = 1; if (rpb > p.H) rpb = p.H;
  p.rows_per_block = rpb;
  dim3 grid((p.H + rpb - 1) / rpb, p.Ctot / cw, p.B);
  if (g.dtype == kBF16) gn_apply_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  else gn_apply_kernel<float><<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// =============================================================================== FIR resampling
// Both resamplers are bound by load / store instruction issue, not by DRAM (ncu, round 1: 25-38 % of DRAM peak at
// 60 % SM busy with one 16-byte access per tap), so the round-2 versions are organised around the fewest and
// widest accesses: a thread owns a 32-byte-aligned PAIR of 16-byte units of the wider tensor (one 256-bit
// LDG / STG, a warp touches 1 KB contiguous), gets its horizontal neighbour from the next lane with shuffles
// (the last lane of a warp loads it), and walks several rows so that every input row is loaded once per block.
// The arithmetic (order of the FMAs) is unchanged from round 1, so the results are bit-identical.
struct U256 { uint4 a, b; };
__device__ __forceinline__ U256 ldg256(const uint4* p) {   // p must be 32-byte aligned
  U256 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(uint4* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ uint4 shfl_down1(const uint4& v) {
  return make_uint4(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1),
                    __shfl_down_sync(0xffffffffu, v.z, 1), __shfl_down_sync(0xffffffffu, v.w, 1));
}

// ops.Resample(down=2) (models/ops.py:52-146) in closed form:
//   y[i,j] = sum_{a,b<4} w_a w_b x[2i-1+a, 2j-1+b],  w = [1,3,3,1]/8, W circular, H zero padded.
// One block = R output rows x 128 output pixels of one plane: thread j owns the padded input pair (2j, 2j+1)
// (taps b = 0, 1; taps 2, 3 are the next lane's pair), filters each of the 2R+2 input rows horizontally once and
// feeds it to the two output rows it belongs to.  Emits GroupNorm partials (one slot per block).
template <typename T>
__global__ void __launch_bounds__(128) down2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int planes,
                                                    int Hi, int Wi, float* __restrict__ stats, int slots,
                                                    int planes_per_unit, int R) {
  pdl_launch_dependents();   // PDL: let the next kernel become resident; wait for the previous one's results
  pdl_wait();
  constexpr int CW = Elem<T>::CW;
  const int Ho = Hi / 2, Wo = Wi / 2;
  const int xsegs = Wo / 128;
  const int xo = (blockIdx.x % xsegs) * 128 + threadIdx.x, yo0 = (blockIdx.x / xsegs) * R;
  const int pl = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const float w[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  const uint4* col = src + pt_index(b, planes, pl, Hi, Wi + 2, 0, 2 * xo);   // padded units 2xo, 2xo+1 of row 0
  const size_t out0 = pt_index(b, planes, pl, Ho, Wo + 2, 0, xo + 1);
  float cur[CW], prev[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) { cur[i] = 0.f; prev[i] = 0.f; }
  float s1 = 0.f, s2 = 0.f;
  // Input rows are visited in pairs (both loads of a pair are in flight together): pair m = image rows
  // 2 (yo0 + m) - 1 and 2 (yo0 + m) carries taps 0, 1 of output yo0 + m (cur) and taps 2, 3 of output yo0 + m - 1
  // (prev), which is complete afterwards.  Same FMA order as one tap at a time.
  auto hfilter = [&](const U256& own, const U256& nb, float* row) {
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) row[i] = 0.f;
    Elem<T>::unpack(own.a, v);
#pragma unroll
    for (int i = 0; i < CW; ++i) row[i] = fmaf(w[0], v[i], row[i]);
    Elem<T>::unpack(own.b, v);
#pragma unroll
    for (int i = 0; i < CW; ++i) row[i] = fmaf(w[1], v[i], row[i]);
    Elem<T>::unpack(nb.a, v);
#pragma unroll
    for (int i = 0; i < CW; ++i) row[i] = fmaf(w[2], v[i], row[i]);
    Elem<T>::unpack(nb.b, v);
#pragma unroll
    for (int i = 0; i < CW; ++i) row[i] = fmaf(w[3], v[i], row[i]);
  };
  for (int m = 0; m <= R; ++m) {
    const int y0 = 2 * (yo0 + m) - 1, y1 = y0 + 1;
    const bool v0 = y0 >= 0, v1 = y1 < Hi;          // (y0 < Hi and y1 >= 0 always hold)
    const uint4* p0 = col + static_cast<ptrdiff_t>(y0) * (Wi + 2);
    const uint4* p1 = p0 + (Wi + 2);
    U256 o0, o1, n0, n1;
    if (v0) o0 = ldg256(p0);
    if (v1) o1 = ldg256(p1);
    if (lane == 31) {
      if (v0) n0 = ldg256(p0 + 2);
      if (v1) n1 = ldg256(p1 + 2);
    }
    if (v0) {
      const uint4 sa = shfl_down1(o0.a), sb = shfl_down1(o0.b);
      if (lane != 31) { n0.a = sa; n0.b = sb; }
      float row[CW];
      hfilter(o0, n0, row);
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        prev[i] = fmaf(w[2], row[i], prev[i]);
        cur[i] = fmaf(w[0], row[i], cur[i]);
      }
    }
    if (v1) {
      const uint4 sa = shfl_down1(o1.a), sb = shfl_down1(o1.b);
      if (lane != 31) { n1.a = sa; n1.b = sb; }
      float row[CW];
      hfilter(o1, n1, row);
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        prev[i] = fmaf(w[3], row[i], prev[i]);
        cur[i] = fmaf(w[1], row[i], cur[i]);
      }
    }
    if (m > 0) {
      const uint4 pk = Elem<T>::pack(prev);
      const size_t idx = out0 + static_cast<size_t>(yo0 + m - 1) * (Wo + 2);
      dst[idx] = pk;
      if (xo == 0) dst[idx + Wo] = pk;
      if (xo == Wo - 1) dst[idx - Wo] = pk;
#pragma unroll
      for (int i = 0; i < CW; ++i) { s1 += prev[i]; s2 += prev[i] * prev[i]; }
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) { prev[i] = cur[i]; cur[i] = 0.f; }
  }
  if (stats != nullptr) {
    __shared__ float red[2][4];
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      const float t = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3];
      const int unit = pl / planes_per_unit, slot = (pl % planes_per_unit) * gridDim.x + blockIdx.x;
      stats[((static_cast<size_t>(b) * kNU + unit) * slots + slot) * 2 + threadIdx.x] = t;
    }
  }
}

static int resample_rows_per_block(int rows) { return rows % 4 == 0 ? 4 : (rows % 2 == 0 ? 2 : 1); }

int down2_stat_slots(int dtype, const PT& dst) {
  const int ppu = dst.C / kNU / dtype_cw(dtype);
  return ppu * (dst.H / resample_rows_per_block(dst.H)) * (dst.W / 128);
}

cudaError_t down2_launch(int dtype, PT src, PT dst, cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  if (dst.W % 128 != 0) return cudaErrorInvalidValue;
  const int R = resample_rows_per_block(dst.H);
  dim3 grid((dst.H / R) * (dst.W / 128), dst.C / cw, dst.B);
  const int ppu = dst.C / kNU / cw;
  if (dtype == kBF16)
    return launch_pdl(down2_kernel<__nv_bfloat16>, grid, dim3(128), 0, s, static_cast<const uint4*>(src.ptr),
                      static_cast<uint4*>(dst.ptr), dst.C / cw, src.H, src.W, dst.stats, dst.slots, ppu, R);
  return launch_pdl(down2_kernel<float>, grid, dim3(128), 0, s, static_cast<const uint4*>(src.ptr),
                    static_cast<uint4*>(dst.ptr), dst.C / cw, src.H, src.W, dst.stats, dst.slots, ppu, R);
}

// ops.Resample(up=2) in closed form (separable): y[2i] = (x[i-1] + 3x[i])/4, y[2i+1] = (3x[i] + x[i+1])/4.
// Thread j owns the padded input unit j (= pixel j-1; pixel j comes from the next lane) and the 32-byte-aligned
// padded output pair (2j, 2j+1) = pixels 2j-1, 2j, which depend on exactly those two input columns; threads
// j = 0 .. Wi therefore also produce both wrap-halo columns.  A block walks R input rows and stores 2R output rows.
template <typename T>
__global__ void __launch_bounds__(128) up2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int planes,
                                                  int Hi, int Wi, int R) {
  pdl_launch_dependents();   // PDL: let the next kernel become resident; wait for the previous one's results
  pdl_wait();
  constexpr int CW = Elem<T>::CW;
  const int Ho = Hi * 2, Wo = Wi * 2;
  const int xsegs = Wi / blockDim.x;
  const int j = (blockIdx.x % xsegs) * blockDim.x + threadIdx.x, i0 = (blockIdx.x / xsegs) * R;
  const int pl = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const bool last = j == Wi - 1;      // this thread also produces the pair of column j + 1 = Wi (pixels Wo-1, halo)
  const uint4* col = src + pt_index(b, planes, pl, Hi, Wi + 2, 0, j);
  uint4* out = dst + pt_index(b, planes, pl, Ho, Wo + 2, 0, 2 * j);
  // one output pair from (near row A, far row B) x (left column l = pixel j-1, right column r = pixel j), in the FMA
  // order of round 1: near row first, and within a row the nearer column first
  auto emit = [&](uint4* o, const uint4& Alr, const uint4& Arr, const uint4& Blr, const uint4& Brr, bool hasB) {
    float Al[CW], Ar[CW], Bl[CW], Br[CW];
    Elem<T>::unpack(Alr, Al); Elem<T>::unpack(Arr, Ar); Elem<T>::unpack(Blr, Bl); Elem<T>::unpack(Brr, Br);
    float lo[CW], hi[CW];   // pixels 2j-1 (nearer to l) and 2j (nearer to r)
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      float x = fmaf(0.5625f, Al[c], 0.f), y = fmaf(0.5625f, Ar[c], 0.f);
      x = fmaf(0.1875f, Ar[c], x); y = fmaf(0.1875f, Al[c], y);
      if (hasB) {
        x = fmaf(0.1875f, Bl[c], x); y = fmaf(0.1875f, Br[c], y);
        x = fmaf(0.0625f, Br[c], x); y = fmaf(0.0625f, Bl[c], y);
      }
      lo[c] = x; hi[c] = y;
    }
    stg256(o, Elem<T>::pack_mma(lo), Elem<T>::pack_mma(hi));
  };
  // previous / current input row as raw units: left, right, (last thread only) the column after
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  uint4 Pl = zero, Pr = zero, Px = zero, Cl = zero, Cr = zero, Cx = zero;
  bool hasP = false;
  auto load_row = [&](int i, uint4& l, uint4& r, uint4& x) {
    const uint4* p = col + static_cast<size_t>(i) * (Wi + 2);
    l = *p;
    r = shfl_down1(l);
    if (lane == 31) r = p[1];
    if (last) x = p[2];
  };
  if (i0 > 0) { load_row(i0 - 1, Pl, Pr, Px); hasP = true; }
  for (int i = i0; i <= i0 + R; ++i) {
    const bool hasC = i < Hi;
    if (hasC) load_row(i, Cl, Cr, Cx);
    if (i > i0) {            // odd output row 2i-1 of input row i-1: near = previous row, far = current row
      uint4* o = out + static_cast<size_t>(2 * i - 1) * (Wo + 2);
      emit(o, Pl, Pr, Cl, Cr, hasC);
      if (last) emit(o + 2, Pr, Px, Cr, Cx, hasC);
    }
    if (i < i0 + R) {        // even output row 2i of input row i: near = current row, far = previous row
      uint4* o = out + static_cast<size_t>(2 * i) * (Wo + 2);
      emit(o, Cl, Cr, Pl, Pr, hasP);
      if (last) emit(o + 2, Cr, Cx, Pr, Px, hasP);
    }
    Pl = Cl; Pr = Cr; Px = Cx;
    hasP = true;
  }
}

cudaError_t up2_launch(int dtype, PT src, PT dst, cudaStream_t s) {
  const int cw = dtype_cw(dtype);
  if (dst.W % 128 != 0) return cudaErrorInvalidValue;
  const int R = resample_rows_per_block(src.H);
  const int threads = src.W % 128 == 0 ? 128 : 64;   // dst.W % 128 == 0 guarantees src.W % 64 == 0
  dim3 grid((src.H / R) * (src.W / threads), dst.C / cw, dst.B);
  if (dtype == kBF16)
    return launch_pdl(up2_kernel<__nv_bfloat16>, grid, dim3(threads), 0, s, static_cast<const uint4*>(src.ptr),
                      static_cast<uint4*>(dst.ptr), dst.C / cw, src.H, src.W, R);
  return launch_pdl(up2_kernel<float>, grid, dim3(threads), 0, s, static_cast<const uint4*>(src.ptr),
                    static_cast<uint4*>(dst.ptr), dst.C / cw, src.H, src.W, R);
}

// =============================================================================== conditioning
// efficient_unet.py:232-237,273-275 + ops.py:14-29,190-198: for every network condition (one per
// sampler step) the sinusoidal embedding, the 2-layer MLP, SiLU and ALL AdaGN projections,
// producing a FiLM table [rows][F] that the GN-apply kernels index by the device step counter.
__global__ void __launch_bounds__(256) cond_temb_kernel(const CondEmbed c) {
  extern __shared__ float sm[];
  float* emb = sm;                 // [base_ch]
  float* h1 = sm + c.base_ch;      // [temb]
  const int row = blockIdx.x;
  const float t = c.cond[row];
  const int half = c.base_ch / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(-logf(10000.f) / static_cast<float>(half - 1) * static_cast<float>(i));
    const float a = t * f;
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c.temb_ch; o += blockDim.x) {
    float acc = c.b1[o];
    const float* w = c.w1 + static_cast<size_t>(o) * c.base_ch;
    for (int k = 0; k < c.base_ch; ++k) acc = fmaf(w[k], emb[k], acc);
    h1[o] = silu_f(acc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c.temb_ch; o += blockDim.x) {
    float acc = c.b2[o];
    const float* w = c.w2 + static_cast<size_t>(o) * c.temb_ch;
    for (int k = 0; k < c.temb_ch; ++k) acc = fmaf(w[k], h1[k], acc);
    c.temb_scratch[static_cast<size_t>(row) * c.temb_ch + o] = silu_f(acc);  // SiLU of AdaGN.proj[0]
  }
}

// film[row][f] = bf[f] + wf[f][:] . silu(temb[row][:]) ; one warp per output, 8 rows per block pass
__global__ void __launch_bounds__(256) cond_film_kernel(const CondEmbed c) {
  extern __shared__ float sm[];  // [rows_here][temb]
  constexpr int RB = 8;
  const int row0 = blockIdx.y * RB;
  const int nrows = min(RB, c.rows - row0);
  for (int i = threadIdx.x; i < nrows * c.temb_ch; i += blockDim.x)
    sm[i] = c.temb_scratch[static_cast<size_t>(row0) * c.temb_ch + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * 8 + warp;
  if (f >= c.F) return;
  const float* w = c.wf + static_cast<size_t>(f) * c.temb_ch;
  float acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  for (int k = lane; k < c.temb_ch; k += 32) {
    const float wk = w[k];
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r < nrows) acc[r] = fmaf(wk, sm[r * c.temb_ch + k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const float t = warp_sum(acc[r]);
    if (lane == 0 && r < nrows) c.film[static_cast<size_t>(row0 + r) * c.F + f] = t + c.bf[f];
  }
}

cudaError_t cond_embed_launch(const CondEmbed& c, cudaStream_t s) {
  cond_temb_kernel<<<c.rows, 256, (c.base_ch + c.temb_ch) * sizeof(float), s>>>(c);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid((c.F + 7) / 8, (c.rows + 7) / 8);
  cond_film_kernel<<<grid, 256, 8 * c.temb_ch * sizeof(float), s>>>(c);
  return cudaGetLastError();
}

// =============================================================================== sampler update
// continuous_time.py:208-229 / discrete_time.py:140-179 folded into per-step scalar coefficients:
//   x0 = clamp(ux x_t + up pred) ;  x_s = kx x_t + k0 x0 + kn noise
// and, for RePaint (continuous_time.py:296-299), x_s = mask (qa known + qs noise2) + (1-mask) x_s.
// ---- in-kernel noise: Philox4x32-10 + Box-Muller through the cuRAND device API, laid out like ATen's
// normal_ kernel so that the values equal torch.randn(generator=<CUDA generator>) bit for bit.
__device__ __forceinline__ unsigned long long philox_base_offset(const PhiloxDraw& ph, int b, int index) {
  const int draw = (ph.ctr0 ? ph.mul0 * *ph.ctr0 : 0) + (ph.ctr1 ? ph.mul1 * *ph.ctr1 : 0) + index;
  return ph.offsets[b] + static_cast<unsigned long long>(ph.offset_per_draw) * static_cast<unsigned long long>(draw);
}
__device__ __forceinline__ float philox_normal_at(unsigned long long seed, unsigned long long offset,
                                                  unsigned threads, size_t l) {
  const unsigned idx = static_cast<unsigned>(l % threads);
  const size_t q = l / threads;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, idx, offset + 4ull * (q >> 2), &st);
  const float4 n = curand_normal4(&st);
  const unsigned c = static_cast<unsigned>(q & 3);
  return c == 0 ? n.x : (c == 1 ? n.y : (c == 2 ? n.z : n.w));
}

__global__ void __launch_bounds__(256) sampler_update_kernel(const SamplerUpdate u) {
  pdl_launch_dependents();   // PDL: let the next kernel become resident; wait for the previous one's results
  pdl_wait();
  const int b = blockIdx.y;
  const int row = (u.step_ptr ? *u.step_ptr : 0) * u.rows_per_step + b * u.row_batch_stride;
  const float* cf = u.coef + static_cast<size_t>(row) * u.coef_cols;
  const float ux = cf[0], up = cf[1], kx = cf[2], k0 = cf[3], kn = cf[4];
  const float qa = u.known ? cf[5] : 0.f, qs = u.known ? cf[6] : 0.f;
  const size_t n4 = u.per_sample / 4;
  const bool gen = u.ph.seeds != nullptr;
  const unsigned long long seed = gen ? u.ph.seeds[b] : 0ull;
  const unsigned long long off1 = gen ? philox_base_offset(u.ph, b, u.draw_noise) : 0ull;
  const unsigned long long off2 = (gen && u.known) ? philox_base_offset(u.ph, b, u.draw_noise2) : 0ull;
  const float4* x = reinterpret_cast<const float4*>(u.x + b * u.per_sample);
  const float4* pr = reinterpret_cast<const float4*>(u.pred + b * u.per_sample);
  const float4* nz = gen ? nullptr : reinterpret_cast<const float4*>(u.noise + b * u.per_sample);
  const float4* kn4 = u.known ? reinterpret_cast<const float4*>(u.known + b * u.per_sample) : nullptr;
  const float4* mk = u.known ? reinterpret_cast<const float4*>(u.mask + b * u.per_sample) : nullptr;
  const float4* n2 = (u.known && !gen) ? reinterpret_cast<const float4*>(u.noise2 + b * u.per_sample) : nullptr;
  float4* xo = reinterpret_cast<float4*>(u.x_out + b * u.per_sample);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 xv = x[i], pv = pr[i];
    float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float ps[4] = {pv.x, pv.y, pv.z, pv.w};
    float ns[4] = {0.f, 0.f, 0.f, 0.f};
    if (!gen) {
      const float4 nv = nz[i];
      ns[0] = nv.x; ns[1] = nv.y; ns[2] = nv.z; ns[3] = nv.w;
    } else if (kn != 0.f) {   // (deterministic DDIM: the reference draws and discards; only the offset advances)
#pragma unroll
      for (int k = 0; k < 4; ++k) ns[k] = philox_normal_at(seed, off1, u.ph.threads, 4 * i + k);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0 = fmaf(ux, xs[k], up * ps[k]);
      if (u.clip > 0.f) x0 = fminf(fmaxf(x0, -u.clip), u.clip);
      xs[k] = fmaf(kx, xs[k], fmaf(k0, x0, kn * ns[k]));
    }
    if (kn4 != nullptr) {
      const float4 kv = kn4[i], mv = mk[i];
      float n2s[4];
      if (!gen) {
        const float4 n2v = n2[i];
        n2s[0] = n2v.x; n2s[1] = n2v.y; n2s[2] = n2v.z; n2s[3] = n2v.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) n2s[k] = philox_normal_at(seed, off2, u.ph.threads, 4 * i + k);
      }
      const float ks[4] = {kv.x, kv.y, kv.z, kv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float known_s = fmaf(qa, ks[k], qs * n2s[k]);
        xs[k] = ms[k] * known_s + (1.f - ms[k]) * xs[k];
      }
    }
    xo[i] = make_float4(xs[0], xs[1], xs[2], xs[3]);
  }
}

cudaError_t sampler_update_launch(const SamplerUpdate& u, cudaStream_t s) {
  if (u.per_sample % 4 != 0) return cudaErrorInvalidValue;
  if (u.ph.seeds != nullptr && (u.ph.offsets == nullptr || u.ph.threads == 0)) return cudaErrorInvalidValue;
  const size_t n4 = u.per_sample / 4;
  int gx = static_cast<int>((n4 + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  dim3 grid(gx, u.B);
  return launch_pdl(sampler_update_kernel, grid, dim3(256), 0, s, u);
}

__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                    const float* __restrict__ ac, float* __restrict__ y,
                                                    size_t per_sample, const int* __restrict__ step_ptr,
                                                    int rows_per_step, int row_batch_stride, const PhiloxDraw ph,
                                                    int draw) {
  pdl_launch_dependents();   // PDL: let the next kernel become resident; wait for the previous one's results
  pdl_wait();
  const int b = blockIdx.y;
  const int row = (step_ptr ? *step_ptr : 0) * rows_per_step + b * row_batch_stride;
  const float a = ac[2 * row], c = ac[2 * row + 1];
  const size_t n4 = per_sample / 4;
  const bool gen = ph.seeds != nullptr;
  const unsigned long long seed = gen ? ph.seeds[b] : 0ull;
  const unsigned long long off = gen ? philox_base_offset(ph, b, draw) : 0ull;
  const float4* xv = reinterpret_cast<const float4*>(x + b * per_sample);
  const float4* nv = gen ? nullptr : reinterpret_cast<const float4*>(noise + b * per_sample);
  float4* yv = reinterpret_cast<float4*>(y + b * per_sample);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 p = xv[i];
    float4 q;
    if (!gen) q = nv[i];
    else q = make_float4(philox_normal_at(seed, off, ph.threads, 4 * i), philox_normal_at(seed, off, ph.threads, 4 * i + 1),
                         philox_normal_at(seed, off, ph.threads, 4 * i + 2), philox_normal_at(seed, off, ph.threads, 4 * i + 3));
    yv[i] = make_float4(fmaf(a, p.x, c * q.x), fmaf(a, p.y, c * q.y), fmaf(a, p.z, c * q.z), fmaf(a, p.w, c * q.w));
  }
}

cudaError_t axpby_launch(const float* x, const float* noise, const float* ac, float* y, int B, size_t per_sample,
                         const int* step_ptr, int rows_per_step, int row_batch_stride, cudaStream_t s,
                         const PhiloxDraw* ph, int draw) {
  if (per_sample % 4 != 0) return cudaErrorInvalidValue;
  int gx = static_cast<int>((per_sample / 4 + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  PhiloxDraw none;
  memset(&none, 0, sizeof(none));
  return launch_pdl(axpby_kernel, dim3(gx, B), dim3(256), 0, s, x, noise, ac, y, per_sample, step_ptr, rows_per_step,
                    row_batch_stride, ph ? *ph : none, draw);
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, const PhiloxDraw ph, int draw,
                                                            size_t per_sample) {
  const int b = blockIdx.y;
  const unsigned long long seed = ph.seeds[b], off = philox_base_offset(ph, b, draw);
  for (size_t l = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; l < per_sample;
       l += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[b * per_sample + l] = philox_normal_at(seed, off, ph.threads, l);
}

cudaError_t philox_normal_launch(float* out, const PhiloxDraw& ph, int draw, int B, size_t per_sample, cudaStream_t s) {
  if (ph.seeds == nullptr || ph.offsets == nullptr || ph.threads == 0) return cudaErrorInvalidValue;
  int gx = static_cast<int>((per_sample + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  philox_normal_kernel<<<dim3(gx, B), 256, 0, s>>>(out, ph, draw, per_sample);
  return cudaGetLastError();
}

__global__ void advance_step_kernel(int* p, int d) {
  pdl_launch_dependents();
  pdl_wait();
  *p += d;
}
cudaError_t advance_step_launch(int* step_ptr, int delta, cudaStream_t s) {
  return launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, s, step_ptr, delta);
}

// =============================================================================== LiDAR epilogue
// sample_and_save.py:52-57 + utils/lidar.py:49-70,99-120: denormalize -> revert_depth -> to_xyz,
// output [B][5][H][W] = depth, x, y, z, reflectance.   depth_format: 0 log, 1 inverse, 2 linear.
__global__ void __launch_bounds__(256) lidar_post_kernel(const float* __restrict__ sample,
                                                         const float* __restrict__ angles, float* __restrict__ out,
                                                         int HW, int fmt, float dmin, float dmax, float log2max) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float d = (sample[(static_cast<size_t>(b) * 2 + 0) * HW + i] + 1.f) * 0.5f;
  const float r = (sample[(static_cast<size_t>(b) * 2 + 1) * HW + i] + 1.f) * 0.5f;
  float metric;
  if (fmt == 0) metric = exp2f(d * log2max) - 1.f;
  else if (fmt == 1) metric = dmin / (d + 1e-8f);
  else metric = d * dmax;
  const float m = (metric > dmin && metric < dmax) ? 1.f : 0.f;
  metric *= m;
  const float m2 = (metric > dmin && metric < dmax) ? 1.f : 0.f;
  const float phi = angles[i], theta = angles[HW + i];
  float sp, cp, st, ct;
  sincosf(phi, &sp, &cp);
  sincosf(theta, &st, &ct);
  float* o = out + static_cast<size_t>(b) * 5 * HW + i;
  o[0] = metric;
  o[HW] = metric * cp * ct * m2;
  o[2 * HW] = metric * cp * st * m2;
  o[3 * HW] = metric * sp * m2;
  o[4 * HW] = r;
}

cudaError_t lidar_postprocess_launch(const float* sample, const float* angles, float* out, int B, int H, int W,
                                     int depth_format, float min_depth, float max_depth, cudaStream_t s) {
  const int HW = H * W;
  lidar_post_kernel<<<dim3((HW + 255) / 256, B), 256, 0, s>>>(sample, angles, out, HW, depth_format, min_depth,
                                                              max_depth, log2f(max_depth + 1.f));
  return cudaGetLastError();
}

}  // namespace r2dm
