// Self-attention core of SelfAttentionBlock (models/efficient_unet.py:42-53, nn.MultiheadAttention
// semantics): per (image, head)  O = softmax(Q K^T / sqrt(hd)) V  over the L = H*W bottleneck tokens.
//
// Round-1 implementation: exact-fp32 flash-style kernel on the FMA pipe (one query per thread,
// K/V tiles broadcast from shared memory, online softmax in the exp2 domain).  The in/out
// projections run on tensor cores through conv_umma (1x1).  A tcgen05 QK^T/PV version is the next
// step for this kernel (see DESIGN.md "what comes next").
#include "common.cuh"
#include "kernels.h"

namespace r2dm {

template <typename T, int HD>
__global__ void __launch_bounds__(128, 2)
attention_kernel(const uint4* __restrict__ qkv, uint4* __restrict__ out, int E, int H, int W, int heads) {
  constexpr int CW = Elem<T>::CW;
  constexpr int KT = 32;             // keys per shared-memory tile
  constexpr int UPT = HD / CW;       // 16-byte units per token per head
  __shared__ __align__(16) float Ks[KT][HD];
  __shared__ __align__(16) float Vs[KT][HD];
  const int L = H * W, Wp = W + 2;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int planes_in = 3 * E / CW, planes_out = E / CW;
  const int tok = qt * 128 + threadIdx.x;
  const int qy = tok / W, qx = tok % W;
  const float qscale = rsqrtf(static_cast<float>(HD)) * 1.4426950408889634f;
  pdl_launch_dependents();
  pdl_wait();

  float q[HD], acc[HD];
#pragma unroll
  for (int u = 0; u < UPT; ++u) {
    float v[CW];
    Elem<T>::unpack(qkv[pt_index(b, planes_in, (h * HD) / CW + u, H, Wp, qy, qx + 1)], v);
#pragma unroll
    for (int i = 0; i < CW; ++i) { q[u * CW + i] = v[i] * qscale; acc[u * CW + i] = 0.f; }
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    // cooperative tile load: KT tokens x UPT units for K and V
    for (int i = threadIdx.x; i < KT * UPT * 2; i += blockDim.x) {
      const int which = i / (KT * UPT);          // 0 = K, 1 = V
      const int r = i % (KT * UPT);
      const int u = r / KT, j = r % KT;          // consecutive threads -> consecutive tokens
      const int t = k0 + j;
      const int ky = t / W, kx = t % W;
      const int plane = ((which + 1) * E + h * HD) / CW + u;
      float v[CW];
      Elem<T>::unpack(qkv[pt_index(b, planes_in, plane, H, Wp, ky, kx + 1)], v);
      float* dstp = which == 0 ? &Ks[j][u * CW] : &Vs[j][u * CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) dstp[c] = v[c];
    }
    __syncthreads();
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&Ks[j][d]);
        a = fmaf(q[d], kv.x, a); a = fmaf(q[d + 1], kv.y, a);
        a = fmaf(q[d + 2], kv.z, a); a = fmaf(q[d + 3], kv.w, a);
      }
      s[j] = a;
      tmax = fmaxf(tmax, a);
    }
    const float mnew = fmaxf(m, tmax);
    const float corr = exp2f(m - mnew);
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
    m = mnew;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float pj = exp2f(s[j] - m);
      l += pj;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        acc[d] = fmaf(pj, vv.x, acc[d]); acc[d + 1] = fmaf(pj, vv.y, acc[d + 1]);
        acc[d + 2] = fmaf(pj, vv.z, acc[d + 2]); acc[d + 3] = fmaf(pj, vv.w, acc[d + 3]);
      }
    }
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int u = 0; u < UPT; ++u) {
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = acc[u * CW + i] * inv;
    const uint4 pk = Elem<T>::pack_mma(v);
    const size_t idx = pt_index(b, planes_out, (h * HD) / CW + u, H, Wp, qy, qx + 1);
    out[idx] = pk;
    if (qx == 0) out[idx + W] = pk;
    if (qx == W - 1) out[idx - W] = pk;
  }
}

cudaError_t attention_launch(int dtype, PT qkv, PT out, int heads, cudaStream_t s) {
  const int E = out.C, hd = E / heads, L = out.H * out.W;
  if (L % 128 != 0 || qkv.C != 3 * E) return cudaErrorInvalidValue;
  dim3 grid(L / 128, heads, out.B);
  const uint4* in = static_cast<const uint4*>(qkv.ptr);
  uint4* o = static_cast<uint4*>(out.ptr);
  if (dtype == kBF16) {
    if (hd == 64) return launch_pdl(attention_kernel<__nv_bfloat16, 64>, grid, dim3(128), 0, s, in, o, E, out.H, out.W, heads);
    if (hd == 32) return launch_pdl(attention_kernel<__nv_bfloat16, 32>, grid, dim3(128), 0, s, in, o, E, out.H, out.W, heads);
  } else {
    if (hd == 64) return launch_pdl(attention_kernel<float, 64>, grid, dim3(128), 0, s, in, o, E, out.H, out.W, heads);
    if (hd == 32) return launch_pdl(attention_kernel<float, 32>, grid, dim3(128), 0, s, in, o, E, out.H, out.W, heads);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace r2dm
