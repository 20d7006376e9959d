// Shared definitions: the "planar-16" activation layout and 16-byte pack/unpack helpers.
//
// Activation tensors live in HBM as  [B][C/CW][H][W+2][CW]  of T, CW = 16/sizeof(T) channels per
// 16-byte unit (8 for bf16, 4 for fp32/tf32).  Column xp = x+1 holds pixel x; xp = 0 and xp = W+1
// are the azimuth wrap halo (copies of x = W-1 and x = 0) so that the ring convolution's TMA boxes
// never cross the seam; the elevation border is produced by TMA out-of-bounds zero fill.
// One channel plane of a tile is exactly the tcgen05 no-swizzle K-major "core matrix" layout
// (8 rows x 16 B contiguous), which lets the 9 taps of a 3x3 convolution be 9 start addresses
// into one halo tile (verified on hardware by tools/probe_umma.cu, profiles/r01_probe_umma.log).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "kernels.h"
#include "ptx.cuh"

namespace r2dm {

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int CW = 4;
  static constexpr int kFmt = 2;  // tf32
  __device__ static __forceinline__ void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y);
    v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                      __float_as_uint(v[3]));
  }
  // for tensors that are only ever read as tcgen05 kind::tf32 operands: round to nearest tf32
  // (the tensor core itself truncates the 13 low mantissa bits, which is biased)
  __device__ static __forceinline__ uint4 pack_mma(const float* v) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r[i]) : "f"(v[i]));
    return make_uint4(r[0], r[1], r[2], r[3]);
  }
  // packed-pair views (CW / 2 pairs per 16-byte unit)
  __device__ static __forceinline__ void unpack2x(const uint4& u, f32x2* v) {
    v[0] = pack2(__uint_as_float(u.x), __uint_as_float(u.y));
    v[1] = pack2(__uint_as_float(u.z), __uint_as_float(u.w));
  }
  __device__ static __forceinline__ uint4 pack2x(const f32x2* v) {
    float a, b, c, d;
    unpack2(v[0], a, b); unpack2(v[1], c, d);
    return make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int CW = 8;
  static constexpr int kFmt = 1;  // bf16
  __device__ static __forceinline__ void unpack(const uint4& u, float* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static __forceinline__ uint4 pack_mma(const float* v) { return pack(v); }
  // packed-pair views (CW / 2 pairs per 16-byte unit): word i holds channels (2i, 2i+1)
  __device__ static __forceinline__ void unpack2x(const uint4& u, f32x2* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = pack2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
  __device__ static __forceinline__ uint4 pack2x(const f32x2* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float lo, hi;
      unpack2(v[i], lo, hi);
      __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// index (in 16-byte units) of pixel (y, xp) of plane `pl` of image b
__host__ __device__ __forceinline__ size_t pt_index(int b, int planes, int pl, int H, int Wp, int y,
                                                    int xp) {
  return ((static_cast<size_t>(b) * planes + pl) * H + y) * Wp + xp;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
// for values that are rounded to tf32 (2^-11) right afterwards: approximate division (MUFU.RCP + FMUL, 2 ulp)
// instead of the IEEE division sequence
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace r2dm
