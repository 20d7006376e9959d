// Self-attention core on tcgen05 tensor cores (bf16 operands, fp32 accumulate / softmax).
//
// SelfAttentionBlock of models/efficient_unet.py:42-53 (nn.MultiheadAttention semantics): per
// (image, head)  O = softmax(Q K^T / sqrt(hd)) V over the L = H*W bottleneck tokens.
//
// One CTA = (image b, head h, 128 consecutive query tokens); 128 threads, thread i owns query row i
// (= TMEM lane i).  Per 128-key tile:
//   S = Q K^T      tcgen05.mma, A = Q tile, B = K tile (both K-major, planar-16 straight from the
//                  packed qkv tensor via TMA), accumulator S[128 x 128] in TMEM
//   softmax        each thread reads its row from TMEM (two passes: max, then exp2), keeps the
//                  running (m, l) and writes P as bf16 into shared memory in the A-operand layout
//   O_tile = P V   tcgen05.mma, A = P (K-major), B = V tile used as an MN-major operand (no
//                  transpose: the planar-16 layout already is the MN-major core-matrix layout)
//   o = o*corr + O_tile   in registers (so no TMEM rescaling pass is needed)
// K/V tiles are double buffered: the TMA for tile j+1 is in flight while tile j is processed, and
// two CTAs per SM overlap each other's MMA and softmax phases.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace r2dm {

struct AttnParams {
  CUtensorMap tmap;   // packed qkv, box = 128 px x (hd/8) planes
  uint4* out;
  int B, E, H, W, heads;
  float scale_log2;   // log2(e) / sqrt(hd)
  int mn_swap;        // developer knob: swap LBO/SBO of the MN-major V descriptor
};

template <int HD>
__global__ void __launch_bounds__(128, 2) attention_umma_kernel(const __grid_constant__ AttnParams p) {
  constexpr int PLANES = HD / 8;                 // 16-byte units per token per head
  constexpr int TILE_BYTES = PLANES * 128 * 16;  // one Q / K / V tile
  constexpr int PLANE_BYTES = 128 * 16;
  constexpr int P_BYTES = 16 * PLANE_BYTES;      // P[128 q x 128 keys] bf16
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;        // [2]
  uint8_t* sV = sK + 2 * TILE_BYTES;    // [2]
  uint8_t* sP = sV + 2 * TILE_BYTES;
  __shared__ uint64_t q_bar, kv_bar[2], mma_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int L = p.H * p.W, nkt = L / 128;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qy = (qt * 128) / p.W, qx0 = (qt * 128) % p.W;
  const int planeQ = (h * HD) / 8, planeK = (p.E + h * HD) / 8, planeV = (2 * p.E + h * HD) / 8;

  if (tid == 0) {
    mbar_init(&q_bar, 1); mbar_init(&kv_bar[0], 1); mbar_init(&kv_bar[1], 1); mbar_init(&mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmap);
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = tmem_slot, tmem_O = tmem_slot + 128;
  // programmatic dependent launch: the set-up above overlapped the tail of the in-projection; nothing
  // produced by it is read before this point, and the out-projection may start its own set-up now
  pdl_launch_dependents();
  pdl_wait();

  auto load_kv = [&](int j, int st) {
    const int ky = (j * 128) / p.W, kx0 = (j * 128) % p.W;
    mbar_expect_tx(&kv_bar[st], 2 * TILE_BYTES);
    tma_load_5d(sK + st * TILE_BYTES, &p.tmap, &kv_bar[st], 2 * (kx0 + 1), 0, ky, planeK, b);
    tma_load_5d(sV + st * TILE_BYTES, &p.tmap, &kv_bar[st], 2 * (kx0 + 1), 0, ky, planeV, b);
  };
  if (tid == 0) {
    mbar_expect_tx(&q_bar, TILE_BYTES);
    tma_load_5d(sQ, &p.tmap, &q_bar, 2 * (qx0 + 1), 0, qy, planeQ, b);
    load_kv(0, 0);
  }

  const uint32_t idesc_s = make_idesc(128, 128, 1);
  const uint32_t idesc_o = make_idesc(128, HD, 1) | (1u << 16);   // B operand MN-major
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  float m = -INFINITY, l = 0.f;
  float o[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) o[i] = 0.f;
  uint32_t mma_phase = 0;

  mbar_wait(&q_bar, 0);
  for (int j = 0; j < nkt; ++j) {
    const int st = j & 1;
    if (tid == 0 && j + 1 < nkt) load_kv(j + 1, st ^ 1);
    mbar_wait(&kv_bar[st], (j >> 1) & 1);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint64_t ad = make_smem_desc(qa + kk * 2 * PLANE_BYTES, PLANE_BYTES, 128, 0);
        const uint64_t bd = make_smem_desc(ka + kk * 2 * PLANE_BYTES, PLANE_BYTES, 128, 0);
        umma_f16(tmem_S, ad, bd, idesc_s, kk > 0);
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // ---- pass 1: row maximum of the scaled scores
    float tmax = -INFINITY;
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float s[32];
      tmem_ld16(tmem_S + lane_base + c0, s);
      tmem_ld16(tmem_S + lane_base + c0 + 16, s + 16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, s[i]);
    }
    const float m_new = fmaxf(m, tmax * p.scale_log2);
    const float corr = exp2f(m - m_new);
    float lsum = 0.f;
    // ---- pass 2: probabilities -> shared memory (A operand layout: [key chunk][query][8 keys])
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float s[32];
      tmem_ld16(tmem_S + lane_base + c0, s);
      tmem_ld16(tmem_S + lane_base + c0 + 16, s + 16);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pv[i] = exp2f(fmaf(s[u * 8 + i], p.scale_log2, -m_new));
          lsum += pv[i];
        }
        *reinterpret_cast<uint4*>(sP + ((c0 / 8 + u) * 128 + tid) * 16) = Elem<__nv_bfloat16>::pack(pv);
      }
    }
    l = l * corr + lsum;
    m = m_new;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {   // 16 keys per MMA
        const uint64_t ad = make_smem_desc(pa + kk * 2 * PLANE_BYTES, PLANE_BYTES, 128, 0);
        // V tile as MN-major B: 8 d contiguous (16 B), keys at 16 B stride inside a plane,
        // key groups of 8 every 128 B, d planes every PLANE_BYTES
        const uint32_t vstart = va + kk * 16 * 16;
        const uint64_t bd = p.mn_swap ? make_smem_desc(vstart, PLANE_BYTES, 128, 0)
                                      : make_smem_desc(vstart, 128, PLANE_BYTES, 0);
        umma_f16(tmem_O, ad, bd, idesc_o, kk > 0);
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 16) {
      float t[16];
      tmem_ld16(tmem_O + lane_base + c0, t);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[c0 + i] = fmaf(o[c0 + i], corr, t[i]);
    }
    tc_fence_before();
  }
  const float inv = 1.f / l;
  const int planes_out = p.E / 8, Wp = p.W + 2;
  const int qx = qx0 + tid;
#pragma unroll
  for (int u = 0; u < PLANES; ++u) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[u * 8 + i] * inv;
    const uint4 pk = Elem<__nv_bfloat16>::pack(v);
    const size_t idx = pt_index(b, planes_out, planeQ + u, p.H, Wp, qy, qx + 1);
    p.out[idx] = pk;
    if (qx == 0) p.out[idx + p.W] = pk;
    if (qx == p.W - 1) p.out[idx - p.W] = pk;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_slot);
}

int attention_make_tmap(CUtensorMap* tm, const PT& qkv, int heads);  // conv_umma.cu helper below

template <int HD>
static cudaError_t launch_attn(const PT& qkv, const PT& out, int heads, const CUtensorMap& tm, cudaStream_t s) {
  constexpr int SMEM = (5 * (HD / 8) * 128 * 16) + 16 * 128 * 16 + 256;
  auto kern = attention_umma_kernel<HD>;
  static unsigned long long configured = 0;   // one bit per device (the attribute is per device)
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
  }
  AttnParams p;
  p.tmap = tm;
  p.out = static_cast<uint4*>(out.ptr);
  p.B = out.B; p.E = out.C; p.H = out.H; p.W = out.W; p.heads = heads;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  static int swap = -1;
  if (swap < 0) { const char* e = getenv("R2DM_ATTN_MNSWAP"); swap = e ? atoi(e) : 0; }
  p.mn_swap = swap;
  dim3 grid(out.H * out.W / 128, heads, out.B);
  return launch_pdl(kern, grid, dim3(128), SMEM, s, p);
}

cudaError_t attention_umma_launch(PT qkv, PT out, int heads, const CUtensorMap& tm, cudaStream_t s) {
  const int E = out.C, hd = E / heads;
  if ((out.H * out.W) % 128 != 0 || out.W % 128 != 0 || qkv.C != 3 * E) return cudaErrorInvalidValue;
  if (hd == 64) return launch_attn<64>(qkv, out, heads, tm, s);
  if (hd == 32) return launch_attn<32>(qkv, out, heads, tm, s);
  return cudaErrorInvalidConfiguration;
}

}  // namespace r2dm
