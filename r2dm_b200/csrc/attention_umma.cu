// Self-attention core on tcgen05 tensor cores (bf16 or tf32 operands, fp32 accumulate / softmax).
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
// The fp32 engine runs the same kernel with kind::tf32 MMAs on the 4-channel planar units (T = float): key tiles
// of 64 (the P tile is 4-byte), P rounded to nearest tf32, Q / K / V read as stored (the tensor core drops the low
// 13 mantissa bits).  kind::tf32 returns zeros for a no-swizzle MN-major B operand (measured on B200; the
// 4-byte MN-major path of the hardware is the 128B / 32B-atom swizzle, which needs 32 contiguous channels per
// token), so the V tile is transposed to K-major in shared memory (4x4 blocks, by all threads, while the
// Q K^T MMA runs).  `r2dm_set_option("attn_exact", 1)` selects the fp32 FMA kernel of attention.cu instead.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace r2dm {

// exp2 as ONE MUFU: exp2f() wraps it in a range fix-up for results below 2^-126 (FSETP + two predicated FMULs per
// value, 3/4 of the softmax's exponential instructions); a probability that small is zero next to the row's 1.0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  CUtensorMap tmap;   // packed qkv, box = 128 px x (hd/CW) planes (Q tile)
  CUtensorMap tmap_kv;  // same tensor, box = KT px (K / V tiles)
  uint4* out;
  int B, E, H, W, heads;
  float scale_log2;   // log2(e) / sqrt(hd)
  int mn_swap;        // developer knob: swap LBO/SBO of the MN-major V descriptor
};

template <typename T, int HD>
struct AttnTraits {
  static constexpr int CW = Elem<T>::CW;
  // (developer A/B: -DR2DM_ATTN_KT_BF16=64 gives 64-key tiles and three CTAs per SM for bf16 - measured slower,
  //  0.106-0.110 vs 0.097 ms for the two launches of a forward)
#ifndef R2DM_ATTN_KT_BF16
#define R2DM_ATTN_KT_BF16 128
#endif
  static constexpr int KT = sizeof(T) == 2 ? R2DM_ATTN_KT_BF16 : 64;   // keys per tile
  static constexpr int KSTEP = 32 / sizeof(T);           // K extent of one MMA (16 bf16 / 8 tf32) = two 16-byte units
  static constexpr int PLANES = HD / CW;                 // 16-byte units per token per head
  static constexpr int PLANE_BYTES = 128 * 16;           // Q / P planes: 128 queries
  static constexpr int KV_PLANE_BYTES = KT * 16;
  static constexpr int Q_BYTES = PLANES * PLANE_BYTES;
  static constexpr int KV_BYTES = PLANES * KV_PLANE_BYTES;
  static constexpr int P_BYTES = (KT / CW) * PLANE_BYTES;   // P[128 q x KT keys]
  // T = float: V^T tile [key chunk of 4][d][4 keys], chunk pitch padded by one unit (bank-conflict-free stores)
  static constexpr int VT_PLANE_BYTES = HD * 16 + 16;
  static constexpr int VT_BYTES = sizeof(T) == 4 ? (KT / 4) * VT_PLANE_BYTES : 0;
  static constexpr int SMEM = Q_BYTES + 4 * KV_BYTES + P_BYTES + VT_BYTES + 256;
  static constexpr int TMEM_COLS = KT + HD <= 128 ? 128 : 256;
};

template <typename T, int HD>
#ifndef R2DM_ATTN_MINB
#define R2DM_ATTN_MINB 2
#endif
__global__ void __launch_bounds__(128, R2DM_ATTN_MINB) attention_umma_kernel(const __grid_constant__ AttnParams p) {
  using Tr = AttnTraits<T, HD>;
  constexpr int CW = Tr::CW, KT = Tr::KT, PLANES = Tr::PLANES;
  constexpr int PLANE_BYTES = Tr::PLANE_BYTES, KV_PLANE_BYTES = Tr::KV_PLANE_BYTES;
  constexpr int TILE_BYTES = Tr::KV_BYTES;       // one K / V tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Tr::Q_BYTES;       // [2]
  uint8_t* sV = sK + 2 * TILE_BYTES;    // [2]
  uint8_t* sP = sV + 2 * TILE_BYTES;
  uint8_t* sVt = sP + Tr::P_BYTES;      // T = float only
  __shared__ uint64_t q_bar, kv_bar[2], mma_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int L = p.H * p.W, nkt = L / KT;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qy = (qt * 128) / p.W, qx0 = (qt * 128) % p.W;
  const int planeQ = (h * HD) / CW, planeK = (p.E + h * HD) / CW, planeV = (2 * p.E + h * HD) / CW;

  if (tid == 0) {
    mbar_init(&q_bar, 1); mbar_init(&kv_bar[0], 1); mbar_init(&kv_bar[1], 1); mbar_init(&mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmap);
    tma_prefetch_desc(&p.tmap_kv);
  }
  if (warp == 0) tmem_alloc<Tr::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = tmem_slot, tmem_O = tmem_slot + KT;
  // programmatic dependent launch: the set-up above overlapped the tail of the in-projection; nothing
  // produced by it is read before this point, and the out-projection may start its own set-up now
  pdl_launch_dependents();
  pdl_wait();

  auto load_kv = [&](int j, int st) {
    const int ky = (j * KT) / p.W, kx0 = (j * KT) % p.W;
    mbar_expect_tx(&kv_bar[st], 2 * TILE_BYTES);
    tma_load_5d(sK + st * TILE_BYTES, &p.tmap_kv, &kv_bar[st], 2 * (kx0 + 1), 0, ky, planeK, b);
    tma_load_5d(sV + st * TILE_BYTES, &p.tmap_kv, &kv_bar[st], 2 * (kx0 + 1), 0, ky, planeV, b);
  };
  if (tid == 0) {
    mbar_expect_tx(&q_bar, Tr::Q_BYTES);
    tma_load_5d(sQ, &p.tmap, &q_bar, 2 * (qx0 + 1), 0, qy, planeQ, b);
    load_kv(0, 0);
  }

  const uint32_t idesc_s = make_idesc(128, KT, Elem<T>::kFmt);
  // P V: bf16 reads the V tile in place as an MN-major B operand; tf32 reads the transposed copy (K-major)
  const uint32_t idesc_o = make_idesc(128, HD, Elem<T>::kFmt) | (sizeof(T) == 2 ? (1u << 16) : 0u);
  auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if (sizeof(T) == 2) umma_f16(d, ad, bd, idesc, acc);
    else umma_tf32(d, ad, bd, idesc, acc);
  };
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
      for (int kk = 0; kk < HD / Tr::KSTEP; ++kk) {
        const uint64_t ad = make_smem_desc(qa + kk * 2 * PLANE_BYTES, PLANE_BYTES, 128, 0);
        const uint64_t bd = make_smem_desc(ka + kk * 2 * KV_PLANE_BYTES, KV_PLANE_BYTES, 128, 0);
        mma(tmem_S, ad, bd, idesc_s, kk > 0);
      }
      umma_commit(&mma_bar);
    }
    if constexpr (sizeof(T) == 4) {
      // V[key][d] -> V^T[key chunk][d][4 keys]: one 4x4 block = 4 units in, 4 units out (the previous tile's
      // P V MMA, the only reader of sVt, completed before the end of the previous iteration)
      const uint8_t* v = sV + st * TILE_BYTES;
      constexpr int KC = KT / 4;
      for (int blk = tid; blk < KC * (HD / 4); blk += 128) {
        const int kc = blk % KC, dp = blk / KC;
        uint4 r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = *reinterpret_cast<const uint4*>(v + (dp * KT + 4 * kc + i) * 16);
        uint8_t* dst = sVt + kc * Tr::VT_PLANE_BYTES + (4 * dp) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(r[0].x, r[1].x, r[2].x, r[3].x);
        *reinterpret_cast<uint4*>(dst + 16) = make_uint4(r[0].y, r[1].y, r[2].y, r[3].y);
        *reinterpret_cast<uint4*>(dst + 32) = make_uint4(r[0].z, r[1].z, r[2].z, r[3].z);
        *reinterpret_cast<uint4*>(dst + 48) = make_uint4(r[0].w, r[1].w, r[2].w, r[3].w);
      }
    }
    mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // ---- pass 1: row maximum of the scaled scores
    float tmax = -INFINITY;
#pragma unroll
    for (int c0 = 0; c0 < KT; c0 += 32) {
      float s[32];
      tmem_ld16(tmem_S + lane_base + c0, s);
      tmem_ld16(tmem_S + lane_base + c0 + 16, s + 16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, s[i]);
    }
    const float m_new = fmaxf(m, tmax * p.scale_log2);
    const float corr = ex2_approx(m - m_new);
    float lsum = 0.f;
    // ---- pass 2: probabilities -> shared memory (A operand layout: [key chunk][query][8 keys])
#pragma unroll
    for (int c0 = 0; c0 < KT; c0 += 32) {
      float s[32];
      tmem_ld16(tmem_S + lane_base + c0, s);
      tmem_ld16(tmem_S + lane_base + c0 + 16, s + 16);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 32 / CW; ++u) {
        float pv[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          pv[i] = ex2_approx(fmaf(s[u * CW + i], p.scale_log2, -m_new));
          lsum += pv[i];
        }
        *reinterpret_cast<uint4*>(sP + ((c0 / CW + u) * 128 + tid) * 16) = Elem<T>::pack_mma(pv);
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
      for (int kk = 0; kk < KT / Tr::KSTEP; ++kk) {   // 16 (bf16) / 8 (tf32) keys per MMA
        const uint64_t ad = make_smem_desc(pa + kk * 2 * PLANE_BYTES, PLANE_BYTES, 128, 0);
        // V tile as MN-major B: CW d contiguous (16 B), keys at 16 B stride inside a plane,
        // key groups of 8 every 128 B, d planes every KV_PLANE_BYTES
        const uint32_t vstart = va + kk * Tr::KSTEP * 16;
        uint64_t bd;
        if (sizeof(T) == 4) bd = make_smem_desc(smem_u32(sVt) + kk * 2 * Tr::VT_PLANE_BYTES, Tr::VT_PLANE_BYTES, 128, 0);
        else bd = p.mn_swap ? make_smem_desc(vstart, KV_PLANE_BYTES, 128, 0) : make_smem_desc(vstart, 128, KV_PLANE_BYTES, 0);
        mma(tmem_O, ad, bd, idesc_o, kk > 0);
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
  const int planes_out = p.E / CW, Wp = p.W + 2;
  const int qx = qx0 + tid;
#pragma unroll
  for (int u = 0; u < PLANES; ++u) {
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = o[u * CW + i] * inv;
    const uint4 pk = Elem<T>::pack_mma(v);   // the only consumer is the out-projection's tensor-core operand path
    const size_t idx = pt_index(b, planes_out, planeQ + u, p.H, Wp, qy, qx + 1);
    p.out[idx] = pk;
    if (qx == 0) p.out[idx + p.W] = pk;
    if (qx == p.W - 1) p.out[idx - p.W] = pk;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<Tr::TMEM_COLS>(tmem_slot);
}

template <typename T, int HD>
static cudaError_t launch_attn(const PT& qkv, const PT& out, int heads, const CUtensorMap& tm_q,
                               const CUtensorMap& tm_kv, cudaStream_t s) {
  using Tr = AttnTraits<T, HD>;
  auto kern = attention_umma_kernel<T, HD>;
  static unsigned long long configured = 0;   // one bit per device (the attribute is per device)
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tr::SMEM);
    if (e != cudaSuccess) return e;
  }
  AttnParams p;
  p.tmap = tm_q;
  p.tmap_kv = tm_kv;
  p.out = static_cast<uint4*>(out.ptr);
  p.B = out.B; p.E = out.C; p.H = out.H; p.W = out.W; p.heads = heads;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  static int swap = -1;
  if (swap < 0) { const char* e = getenv("R2DM_ATTN_MNSWAP"); swap = e ? atoi(e) : 0; }
  p.mn_swap = swap;
  dim3 grid(out.H * out.W / 128, heads, out.B);
  return launch_pdl(kern, grid, dim3(128), Tr::SMEM, s, p);
}

int attention_key_tile(int dtype) { return dtype == kBF16 ? R2DM_ATTN_KT_BF16 : 64; }

cudaError_t attention_umma_launch(int dtype, PT qkv, PT out, int heads, const CUtensorMap& tm_q,
                                  const CUtensorMap& tm_kv, cudaStream_t s) {
  const int E = out.C, hd = E / heads;
  if ((out.H * out.W) % 128 != 0 || out.W % 128 != 0 || qkv.C != 3 * E) return cudaErrorInvalidValue;
  if (dtype == kBF16) {
    if (hd == 64) return launch_attn<__nv_bfloat16, 64>(qkv, out, heads, tm_q, tm_kv, s);
    if (hd == 32) return launch_attn<__nv_bfloat16, 32>(qkv, out, heads, tm_q, tm_kv, s);
  } else {
    if (hd == 64) return launch_attn<float, 64>(qkv, out, heads, tm_q, tm_kv, s);
    if (hd == 32) return launch_attn<float, 32>(qkv, out, heads, tm_q, tm_kv, s);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace r2dm
