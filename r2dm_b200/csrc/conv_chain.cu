// A chain of ring 3x3 convolutions of one resolution level in ONE persistent launch.
//
// The launch program of a level is conv1, conv2, conv1, conv2, ... with identical output shape and tile geometry
// (models/efficient_unet.py:56-110: the ResidualBlocks of a Block).  One launch per layer (conv_umma.cu) costs, per
// layer and SM, a pipeline fill (dependency wait, first TMA round trip, statistics fold, first transform) and a
// drain (last epilogue) in which the tensor pipe idles, plus a last round in which only some CTAs have a tile
// (512 tiles on 148 SMs = 3.46 per CTA).  Here every CTA walks the layers itself:
//   * the batch is split into two image groups A and B; per layer a CTA processes its tiles of group A, then of
//     group B.  A tile of layer l and image b needs ALL of image b of layer l-1 (GroupNorm statistics + halo rows):
//     when the CTA reaches (l, A) the other CTAs finished (l-1, A) while everybody was busy with (l-1, B), so the
//     producer / transform / MMA / epilogue pipeline never drains between layers;
//   * the tile ranges rotate from segment to segment, so that over the whole chain every CTA gets the same number
//     of tiles (6 x 512 = 3072 tiles = 20.8 per CTA) instead of max(3, 4) per layer;
//   * dependencies are per-(layer, image) counters in global memory: the epilogue releases (stores, fence,
//     barrier, red.release) one count per finished tile, the producer / transform / epilogue roles acquire
//     `tiles per image` before they touch image b of layer l-1.  TMA reads data written by other CTAs of the same
//     launch, so both sides add fence.proxy.async (generic <-> async proxy); residual / statistics loads bypass
//     L1 (a recycled buffer may have been read through L1 earlier in the launch).
// Dependencies always point to a lower layer and all CTAs are co-resident (grid <= #SMs, one CTA per SM), so the
// waits cannot deadlock.  Same warp roles, pipeline, MMA issue order and epilogue arithmetic as conv_umma_kernel -
// the results are bit-identical to the one-launch-per-layer path.
#include "conv_common.cuh"

namespace r2dm {

struct ChainLayer {
  CUtensorMap tmap0, tmap1;   // input(s)
  CUtensorMap tmap2, tmap3;   // folded skip projection inputs
  XformParams xf;
  const void* wpacked; const float* bias;
  const void* w2packed; const float* bias2;
  int nk, ksplit, nk2, ksplit2;
  const void* residual; void* out; float* stats;
  int slots; float scale;
};

struct ChainParams {
  ChainLayer layer[kMaxChainLayers];
  int nlayers;
  int B, H, W, cout_pad;
  int xtiles, ytiles, ntiles, tiles_per_image;
  int unit_ch;
  int stages, stage_bytes, coef_ch, coef_bytes;
  int rot;          // rotation stride of the tile ranges from segment to segment
  int split;        // images in group A (groups: [0, split), [split, B)); split == B: one group
  int* done;        // [nlayers][B] finished-tile counters, zero on entry
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint4 ldcg128(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldcg_f2(const float2* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// The sequence of tiles of one CTA: layers ascending, per layer group A then group B, per segment a contiguous,
// rotating share of the group's tiles.  Every role runs its own copy.
struct ChainIter {
  int l = 0, g = -1, t = 0, t_end = 0, ngroups, first_image = 0;
  __device__ __forceinline__ bool next(const ChainParams& p, int& b, int& yt, int& xt, int& nt) {
    while (t >= t_end) {
      if (++g == ngroups) { g = 0; ++l; }
      if (l >= p.nlayers) return false;
      const int images = g == 0 ? p.split : p.B - p.split;
      first_image = g == 0 ? 0 : p.split;
      const int n = images * p.tiles_per_image;
      const int G = static_cast<int>(gridDim.x);
      const int r = (static_cast<int>(blockIdx.x) + (l * ngroups + g) * p.rot) % G;
      t = static_cast<int>(static_cast<long long>(r) * n / G);
      t_end = static_cast<int>(static_cast<long long>(r + 1) * n / G);
    }
    int q = t++;
    nt = q % p.ntiles; q /= p.ntiles;
    xt = q % p.xtiles; q /= p.xtiles;
    yt = q % p.ytiles;
    b = first_image + q / p.ytiles;
    return true;
  }
};

template <typename T, int NT, int HT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_chain_kernel(const __grid_constant__ ChainParams p) {
  using Tr = ConvTraits<T, NT, HT, 9, 1>;
  static_assert(Tr::FUSE, "the chain kernel is built for the tap-fused tile shapes");
  constexpr int CW = Tr::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], xf_bar[kMaxStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float stat_w[2][8][NT / 8][2];
  __shared__ float grp_s[2][kNU];
  __shared__ __align__(16) float bias_s[NT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* coef_a = reinterpret_cast<float*>(smem);
  float* coef_d = coef_a + p.coef_ch;
  uint8_t* smem_ring = smem + p.coef_bytes;
  const int ngroups = p.split < p.B ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&xf_bar[i], 4); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
    fence_mbar_init();
    for (int l = 0; l < p.nlayers; ++l) {
      tma_prefetch_desc(&p.layer[l].tmap0);
      if (p.layer[l].ksplit < p.layer[l].nk) tma_prefetch_desc(&p.layer[l].tmap1);
    }
  }
  if (warp == kAllocWarp) tmem_alloc<Tr::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_launch_dependents();

  // image b of layer l - 1 complete (all its tiles stored and released)?
  auto wait_image = [&](int l, int b) {
    if (l == 0) return;
    const int* c = p.done + (l - 1) * p.B + b;
    while (ld_acquire(c) < p.tiles_per_image) __nanosleep(64);
  };

  if (warp == kProdWarp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      pdl_wait();
      ChainIter itr; itr.ngroups = ngroups;
      uint32_t ph = 0;
      int st = 0, b, yt, xt, nt, cur_l = -1, cur_b = -1;
      while (itr.next(p, b, yt, xt, nt)) {
        const ChainLayer& L = p.layer[itr.l];
        if (itr.l != cur_l || b != cur_b) {
          cur_l = itr.l; cur_b = b;
          wait_image(itr.l, b);
          fence_proxy_async_all();     // generic-proxy writes of the other CTAs -> our async-proxy (TMA) reads
        }
        const int x0 = xt * 128, y0 = yt * HT;
        const uint8_t* wsrc = static_cast<const uint8_t*>(L.wpacked) + static_cast<size_t>(nt) * L.nk * Tr::B_BYTES;
        for (int ks = 0; ks < L.nk; ++ks, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_relaxed(&empty_bar[st], ph ^ 1, 2000);
          uint8_t* sa = smem_ring + static_cast<size_t>(st) * p.stage_bytes;
          mbar_expect_tx(&full_bar[st], Tr::A_BYTES + Tr::B_BYTES);
          const bool second = ks >= L.ksplit;
          const int plane0 = (second ? ks - L.ksplit : ks) * Tr::PLANES;
          tma_load_5d(sa, second ? &L.tmap1 : &L.tmap0, &full_bar[st], 2 * x0, 0, y0 - 1, plane0, b);
          bulk_load(sa + Tr::A_BYTES_AL, wsrc + static_cast<size_t>(ks) * Tr::B_BYTES, Tr::B_BYTES, &full_bar[st]);
        }
        const uint8_t* w2src = static_cast<const uint8_t*>(L.w2packed) + static_cast<size_t>(nt) * L.nk2 * Tr::SK_B_BYTES;
        for (int ks = 0; ks < L.nk2; ++ks, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_relaxed(&empty_bar[st], ph ^ 1, 2000);
          uint8_t* sa = smem_ring + static_cast<size_t>(st) * p.stage_bytes;
          mbar_expect_tx(&full_bar[st], Tr::SK_A_BYTES + Tr::SK_B_BYTES);
          const bool second = ks >= L.ksplit2;
          const int plane0 = (second ? ks - L.ksplit2 : ks) * Tr::SK_PLANES;
          tma_load_5d(sa, second ? &L.tmap3 : &L.tmap2, &full_bar[st], 2 * (x0 + 1), 0, y0, plane0, b);
          bulk_load(sa + Tr::SK_A_BYTES, w2src + static_cast<size_t>(ks) * Tr::SK_B_BYTES, Tr::SK_B_BYTES, &full_bar[st]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (warp-convergent)
    const uint32_t idesc = make_idesc(128, NT, Elem<T>::kFmt);
    const uint32_t idesc2 = make_idesc(128, 2 * NT <= 256 ? 2 * NT : NT, Elem<T>::kFmt);
    const uint32_t idesc3 = make_idesc(128, 3 * NT <= 256 ? 3 * NT : NT, Elem<T>::kFmt);
    (void)idesc2; (void)idesc3;
    const uint32_t a_lo_const = static_cast<uint32_t>(Tr::A_PLANE_BYTES >> 4) << 16;
    const uint32_t b_lo_const = static_cast<uint32_t>(Tr::B_PLANE_BYTES >> 4) << 16;
    constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
    ChainIter itr; itr.ngroups = ngroups;
    uint32_t ph = 0;
    int st = 0, j = 0, b, yt, xt, nt;
    while (itr.next(p, b, yt, xt, nt)) {
      const int nk = p.layer[itr.l].nk, nk2 = p.layer[itr.l].nk2;
      const int buf = j & 1;
      mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t dbase = tmem + buf * Tr::ACC_COLS;
      for (int ks = 0; ks < nk; ++ks, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
        mbar_wait(&xf_bar[st], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes);
        const uint32_t sb = sa + Tr::A_BYTES_AL;
        const uint32_t a_lo0 = a_lo_const | ((sa >> 4) & 0x3FFFu);
        const uint32_t b_lo0 = b_lo_const | ((sb >> 4) & 0x3FFFu);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t a_k = static_cast<uint32_t>((kx * 16) >> 4);
          const uint32_t b_k = static_cast<uint32_t>((kx * Tr::B_TAP_BYTES) >> 4);
          // same issue order as conv_umma_kernel (first-touch rows clear the accumulators)
          constexpr int IR_A = HT - 1 < 2 ? HT - 1 : 2;
          constexpr int IR_B = HT - 1 > IR_A ? HT + 1 : -1;
          const bool first = ks == 0 && kx == 0;
#pragma unroll
          for (int idx = 0; idx < HT + 2; ++idx) {
            int ir = idx;
            if (idx == 0) ir = IR_A;
            else if (IR_B >= 0 && idx == 1) ir = IR_B;
            else {
              int k = idx - (IR_B >= 0 ? 2 : 1);
              ir = 0;
              for (int c = 0; c < HT + 2; ++c) {
                if (c == IR_A || c == IR_B) continue;
                if (k == 0) { ir = c; break; }
                --k;
              }
            }
            const int r_lo = ir - 2 > 0 ? ir - 2 : 0, r_hi = ir < HT - 1 ? ir : HT - 1;
            const int ky_hi = ir - r_lo, nrows = r_hi - r_lo + 1;
            const uint64_t adesc = (static_cast<uint64_t>(kHi) << 32) |
                                   (a_lo0 + a_k + static_cast<uint32_t>((ir * Tr::APITCH * 16) >> 4));
            const uint64_t bdesc = (static_cast<uint64_t>(kHi) << 32) |
                                   (b_lo0 + b_k + static_cast<uint32_t>(((2 - ky_hi) * NT * 16) >> 4));
            const uint32_t idn = nrows == 3 ? idesc3 : (nrows == 2 ? idesc2 : idesc);
            const uint32_t acc = (first && (ir == IR_A || ir == IR_B)) ? 0u : 1u;
            if (Elem<T>::kFmt == 2) umma_tf32_warp(dbase + r_lo * NT, adesc, bdesc, idn, acc);
            else umma_f16_warp(dbase + r_lo * NT, adesc, bdesc, idn, acc);
          }
        }
        umma_commit_warp(&empty_bar[st]);
      }
      for (int ks = 0; ks < nk2; ++ks, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
        mbar_wait(&xf_bar[st], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes);
        const uint32_t sb = sa + Tr::SK_A_BYTES;
        const uint32_t a_lo0 = (static_cast<uint32_t>(Tr::SK_A_PLANE_BYTES >> 4) << 16) | ((sa >> 4) & 0x3FFFu);
        const uint32_t b_lo0 = (static_cast<uint32_t>(Tr::SK_B_PLANE_BYTES >> 4) << 16) | ((sb >> 4) & 0x3FFFu);
#pragma unroll
        for (int kk = 0; kk < Tr::SK_PLANES / 2; ++kk) {
#pragma unroll
          for (int r = 0; r < HT; ++r) {
            const uint32_t a_add = static_cast<uint32_t>((kk * 2 * Tr::SK_A_PLANE_BYTES + r * 128 * 16) >> 4);
            const uint32_t b_add = static_cast<uint32_t>((kk * 2 * Tr::SK_B_PLANE_BYTES) >> 4);
            const uint64_t adesc = (static_cast<uint64_t>(kHi) << 32) | (a_lo0 + a_add);
            const uint64_t bdesc = (static_cast<uint64_t>(kHi) << 32) | (b_lo0 + b_add);
            if (Elem<T>::kFmt == 2) umma_tf32_warp(dbase + r * NT, adesc, bdesc, idesc, 1u);
            else umma_f16_warp(dbase + r * NT, adesc, bdesc, idesc, 1u);
          }
        }
        umma_commit_warp(&empty_bar[st]);
      }
      umma_commit_warp(&acc_full[buf]);
      ++j;
    }
  } else if (warp >= kXfWarp0 && warp < kXfWarp0 + 8) {
    // ------------------------------------------------------------------ operand transform (two groups of 4 warps)
    pdl_wait();
    const int grp = (warp - kXfWarp0) >> 2;
    const int t256 = threadIdx.x - kXfWarp0 * 32;
    const int tt = t256 & 127;
    constexpr int TPP = 128 / Tr::PLANES;
    const int my_plane = tt / TPP, tip = tt % TPP;
    ChainIter itr; itr.ngroups = ngroups;
    uint32_t it = 0, ph = 0;
    int st = 0, b, yt, xt, nt, cur_l = -1, cur_b = -1;
    while (itr.next(p, b, yt, xt, nt)) {
      const ChainLayer& L = p.layer[itr.l];
      const XformParams& X = L.xf;
      if (itr.l != cur_l || b != cur_b) {
        cur_l = itr.l; cur_b = b;
        wait_image(itr.l, b);
        // ---- fold statistics + affine / FiLM into per-channel (a, d) for (layer, image)
        const int Ctot = X.C0 + X.C1;
        const int gsize = Ctot / X.groups;
        const double inv_cnt = 1.0 / (static_cast<double>(gsize) * p.H * p.W);
        const float* fl = nullptr;
        if (X.film != nullptr) {
          const int row = (X.step_ptr ? *X.step_ptr : 0) * X.rows_per_step + b * X.row_batch_stride;
          fl = X.film + static_cast<size_t>(row) * X.film_stride + X.film_off;
        }
        constexpr int CPT = kMaxCin / 256;
        float ga_r[CPT], be_r[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int c = t256 + k * 256;
          ga_r[k] = 0.f; be_r[k] = 0.f;
          if (c < Ctot) {
            ga_r[k] = fl ? 1.f + fl[c] : X.gamma[c];
            be_r[k] = fl ? fl[Ctot + c] : X.beta[c];
          }
        }
        {
          const int g = t256 >> 5;
          double s1 = 0.0, s2 = 0.0;
          const int lo = g * gsize, hi_c = lo + gsize;
          int off = 0;
          for (int si = 0; si < 2; ++si) {
            const int Cs = si == 0 ? X.C0 : X.C1;
            if (Cs == 0) break;
            const float* stp = si == 0 ? X.stats0 : X.stats1;
            const int sl = si == 0 ? X.slots0 : X.slots1;
            const int a = max(lo, off), e = min(hi_c, off + Cs);
            if (a < e) {
              const int unit_ch = Cs / kNU;
              const int u0 = (a - off) / unit_ch, u1 = (e - off) / unit_ch;
              const int n = (u1 - u0) * sl;
              const float2* st2 = reinterpret_cast<const float2*>(stp + (static_cast<size_t>(b) * kNU + u0) * sl * 2);
              for (int i = lane; i < n; i += 128) {
                float2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = i + 32 * u < n ? ldcg_f2(st2 + i + 32 * u) : make_float2(0.f, 0.f);
                s1 += (static_cast<double>(v[0].x) + v[1].x) + (static_cast<double>(v[2].x) + v[3].x);
                s2 += (static_cast<double>(v[0].y) + v[1].y) + (static_cast<double>(v[2].y) + v[3].y);
              }
            }
            off += Cs;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (lane == 0) {
            const double mean = s1 * inv_cnt;
            double var = s2 * inv_cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            grp_s[0][g] = static_cast<float>(mean);
            grp_s[1][g] = rsqrtf(static_cast<float>(var) + X.eps);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float fold = X.silu ? 0.5f : 1.f;   // tanh-form SiLU for both engines (see conv_umma.cu)
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int c = t256 + k * 256;
          if (c < Ctot) {
            const int g = c / gsize;
            const float a = grp_s[1][g] * ga_r[k];
            coef_a[c] = a * fold;
            coef_d[c] = (be_r[k] - grp_s[0][g] * a) * fold;
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      const int y_first = yt * HT - 1;
      const int row_lo = max(0, -y_first), row_hi = min(Tr::AROWS, p.H - y_first);
      const int n_units = (row_hi - row_lo) * Tr::APITCH;
      for (int ks = 0; ks < L.nk; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
        if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
        f32x2 ca2[CW / 2], cd2[CW / 2];
        const int c0 = (ks * Tr::PLANES + my_plane) * CW;
#pragma unroll
        for (int i = 0; i < CW / 4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(coef_a + c0 + 4 * i);
          const float4 d = *reinterpret_cast<const float4*>(coef_d + c0 + 4 * i);
          ca2[2 * i] = pack2(a.x, a.y); ca2[2 * i + 1] = pack2(a.z, a.w);
          cd2[2 * i] = pack2(d.x, d.y); cd2[2 * i + 1] = pack2(d.z, d.w);
        }
        mbar_wait_relaxed(&full_bar[st], ph, 500);
        const uint32_t sbase = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes +
                                        my_plane * Tr::A_PLANE_BYTES) + row_lo * Tr::APITCH * 16;
        auto xform_unit = [&](uint4 raw) {
          f32x2 v2[CW / 2];
          Elem<T>::unpack2x(raw, v2);
#pragma unroll
          for (int k = 0; k < CW / 2; ++k) {
            const f32x2 t2 = fma2(v2[k], ca2[k], cd2[k]);
            if (X.silu) {
              float lo, hi;
              unpack2(t2, lo, hi);
              float tl, th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(tl) : "f"(lo));
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hi));
              v2[k] = fma2(t2, pack2(tl, th), t2);
            } else {
              v2[k] = t2;
            }
          }
          if (sizeof(T) == 2) return Elem<T>::pack2x(v2);
          float v[CW];
#pragma unroll
          for (int k = 0; k < CW / 2; ++k) unpack2(v2[k], v[2 * k], v[2 * k + 1]);
          return Elem<T>::pack_mma(v);
        };
        constexpr int XB = 4;
        int i0 = tip;
        for (; i0 + (XB - 1) * TPP < n_units; i0 += XB * TPP) {
          uint4 raw[XB];
#pragma unroll
          for (int u = 0; u < XB; ++u) raw[u] = lds128(sbase + (i0 + u * TPP) * 16);
#pragma unroll
          for (int u = 0; u < XB; ++u) raw[u] = xform_unit(raw[u]);
#pragma unroll
          for (int u = 0; u < XB; ++u) sts128(sbase + (i0 + u * TPP) * 16, raw[u]);
        }
        for (; i0 < n_units; i0 += TPP) sts128(sbase + i0 * 16, xform_unit(lds128(sbase + i0 * 16)));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[st]);
      }
      for (int ks = 0; ks < L.nk2; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
        if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
        mbar_wait_relaxed(&full_bar[st], ph, 500);
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[st]);
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 8) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - kEpiWarp0;
    const int q = ew & 3;
    const int half = ew >> 2;
    const int m = q * 32 + lane;
    const int Wp = p.W + 2;
    const int planes_out = p.cout_pad / CW;
    const uint32_t plane_stride = static_cast<uint32_t>(p.H) * static_cast<uint32_t>(Wp);
    constexpr int CB = 32;
    constexpr int CCOLS = HT > 1 ? NT : NT / 2;
    constexpr int NCHUNK = CCOLS / CB;
    constexpr int NSUB = CB / 8;
    constexpr int RSTEP = HT > 1 ? 2 : 1;
    const int r_begin = HT > 1 ? half : 0;
    const int c_begin = HT > 1 ? 0 : half * (NT / 2);
    const int ethread = threadIdx.x - kEpiWarp0 * 32;
    pdl_wait();
    ChainIter itr; itr.ngroups = ngroups;
    int j = 0, cur_nt = -1, cur_l = -1, cur_b = -1, b, yt, xt, nt;
    while (itr.next(p, b, yt, xt, nt)) {
      const ChainLayer& L = p.layer[itr.l];
      const uint4* res = static_cast<const uint4*>(L.residual);
      uint4* out = static_cast<uint4*>(L.out);
      const int n0 = nt * NT, x = xt * 128 + m, y0 = yt * HT;
      const int buf = j & 1, par = j & 1;
      const bool seam = (x == 0) || (x == p.W - 1);
      const int seam_off = (x == 0) ? p.W : -p.W;
      const bool new_layer = itr.l != cur_l;
      if (new_layer || b != cur_b) {
        // the residual of image b was produced two layers back by other CTAs of this launch: acquire
        cur_b = b;
        if (res != nullptr) wait_image(itr.l, b);
      }
      if (nt != cur_nt || new_layer) {
        cur_nt = nt; cur_l = itr.l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = ethread; i < NT; i += 256)
          bias_s[i] = (L.bias[n0 + i] + (L.bias2 ? L.bias2[n0 + i] : 0.f)) * L.scale;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (res != nullptr) {
        constexpr int SEGS = HT * (NT / CW);
        for (int l2 = ethread; l2 < SEGS * 16; l2 += 256) {
          const int seg = l2 >> 4, row = seg / (NT / CW), pl = seg % (NT / CW);
          if (y0 + row < p.H) {
            const uint4* a = res + pt_index(b, planes_out, n0 / CW + pl, p.H, Wp, y0 + row, xt * 128 + 1) + (l2 & 15) * 8;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
      const uint32_t tile_idx = static_cast<uint32_t>(pt_index(b, planes_out, n0 / CW, p.H, Wp, y0, x + 1));
      mbar_wait_relaxed(&acc_full[buf], (j >> 1) & 1, 1000);
      tc_fence_after();
      const uint32_t tbase = tmem + buf * Tr::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c0 = c_begin + ch * CB;
        f32x2 s1p[NSUB], s2p[NSUB];
#pragma unroll
        for (int u = 0; u < NSUB; ++u) { s1p[u] = pack2(0.f, 0.f); s2p[u] = pack2(0.f, 0.f); }
        const f32x2 scale2 = pack2(L.scale, L.scale);
#pragma unroll
        for (int r = r_begin; r < HT; r += RSTEP) {
          const int y = y0 + r;
          if (y < p.H) {
            const uint32_t idx0 = tile_idx + static_cast<uint32_t>(r) * static_cast<uint32_t>(Wp) +
                                  static_cast<uint32_t>(c0 / CW) * plane_stride;
            uint4 rr[CB / CW];
            if (res != nullptr) {
#pragma unroll
              for (int u = 0; u < CB / CW; ++u) rr[u] = ldcg128(res + idx0 + u * plane_stride);
            }
            float v[CB];
#pragma unroll
            for (int h16 = 0; h16 < CB / 16; ++h16) tmem_ld16(tbase + r * NT + c0 + h16 * 16, v + h16 * 16);
            tmem_ld_wait();
            const float4* bias4 = reinterpret_cast<const float4*>(bias_s + c0);
            f32x2 v2[CB / 2];
#pragma unroll
            for (int i4 = 0; i4 < CB / 4; ++i4) {
              const float4 bv = bias4[i4];
              v2[2 * i4] = fma2(pack2(v[4 * i4], v[4 * i4 + 1]), scale2, pack2(bv.x, bv.y));
              v2[2 * i4 + 1] = fma2(pack2(v[4 * i4 + 2], v[4 * i4 + 3]), scale2, pack2(bv.z, bv.w));
            }
            constexpr int PPU = CW / 2;
#pragma unroll
            for (int u = 0; u < CB / CW; ++u) {
              const uint32_t idx = idx0 + u * plane_stride;
              f32x2 o2[PPU];
#pragma unroll
              for (int i = 0; i < PPU; ++i) o2[i] = v2[u * PPU + i];
              if (res != nullptr) {
                f32x2 r2[PPU];
                Elem<T>::unpack2x(rr[u], r2);
#pragma unroll
                for (int i = 0; i < PPU; ++i) o2[i] = fma2(r2[i], scale2, o2[i]);
              }
              const int sub = (u * CW) / 8;
#pragma unroll
              for (int i = 0; i < PPU; ++i) {
                s1p[sub] = add2(s1p[sub], o2[i]);
                s2p[sub] = fma2(o2[i], o2[i], s2p[sub]);
              }
              const uint4 pk = Elem<T>::pack2x(o2);
              out[idx] = pk;
              if (seam) out[static_cast<int>(idx) + seam_off] = pk;
            }
          }
        }
        float vals[2 * NSUB];
#pragma unroll
        for (int u = 0; u < NSUB; ++u) {
          float lo, hi;
          unpack2(s1p[u], lo, hi); vals[2 * u] = lo + hi;
          unpack2(s2p[u], lo, hi); vals[2 * u + 1] = lo + hi;
        }
        constexpr int LOGV = 3;   // log2(2 * NSUB), NSUB = 4
#pragma unroll
        for (int rd = 0; rd < LOGV; ++rd) {
          const int nv = (2 * NSUB) >> rd, off = 16 >> rd;
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < nv / 2; ++i) {
            const float send = upper ? vals[i] : vals[i + nv / 2];
            const float keep = upper ? vals[i + nv / 2] : vals[i];
            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
#pragma unroll
        for (int off = 16 >> LOGV; off > 0; off >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
        constexpr int LPV = 32 / (2 * NSUB);
        if ((lane & (LPV - 1)) == 0) {
          const int vi = lane / LPV;
          stat_w[par][ew][(c0 >> 3) + (vi >> 1)][vi & 1] = vals[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      // all stores of this tile are ordered before the barrier at CTA scope; the releasing thread's device-scope
      // fence below is cumulative over them
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int units_here = NT / p.unit_ch;
      if (ethread < units_here * 2) {
        const int u = ethread >> 1, k = ethread & 1;
        const int cpu = p.unit_ch >> 3;
        float tot = 0.f;
        for (int sc = u * cpu; sc < (u + 1) * cpu; ++sc) {
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const bool has = HT > 1 ? true : (w >> 2) == (sc >= NT / 16 ? 1 : 0);
            if (has) tot += stat_w[par][w][sc][k];
          }
        }
        const int unit = n0 / p.unit_ch + u;
        const int slot = yt * p.xtiles + xt;
        L.stats[((static_cast<size_t>(b) * kNU + unit) * L.slots + slot) * 2 + k] = tot;
      }
      // the statistics writers are in warp 0 of the epilogue group (ethread < 32): one more warp-level sync, then
      // lane 0 publishes the tile (fence: device scope + generic -> async proxy, then a release increment)
      if (ew == 0) {
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          fence_proxy_async_all();
          red_release_add(p.done + itr.l * p.B + b, 1);
        }
      }
      ++j;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) tmem_dealloc<Tr::TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------ host side
template <typename T, int NT, int HT>
static cudaError_t launch_chain(const ConvLaunch* const* ls, int n, int* done, int batch_split, cudaStream_t s) {
  using Tr = ConvTraits<T, NT, HT, 9, 1>;
  auto kern = conv_chain_kernel<T, NT, HT>;
  constexpr int kBudget = 224 * 1024;
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBudget);
    if (e != cudaSuccess) return e;
  }
  static ChainParams p;   // large (several KB): keep it off the stack; launches are issued from one thread
  memset(&p, 0, sizeof(p));
  const ConvLaunch& f = *ls[0];
  p.nlayers = n;
  p.B = f.out.B; p.H = f.out.H; p.W = f.out.W; p.cout_pad = f.cout_pad;
  p.xtiles = f.out.W / 128; p.ytiles = (f.out.H + HT - 1) / HT; p.ntiles = f.cout_pad / NT;
  p.tiles_per_image = p.xtiles * p.ytiles * p.ntiles;
  p.unit_ch = f.cout / kNU >= 8 ? f.cout / kNU : 8;
  if ((p.unit_ch & (p.unit_ch - 1)) != 0 || p.unit_ch > NT) return cudaErrorInvalidValue;
  int coef_ch = 0;
  bool any_skip = false;
  for (int i = 0; i < n; ++i) {
    const ConvLaunch& l = *ls[i];
    ChainLayer& L = p.layer[i];
    if (!l.xf.enabled || l.out_nchw || l.relu || l.colmax || l.taps != 9 || l.nt != NT || l.ht != HT ||
        l.out.B != p.B || l.out.H != p.H || l.out.W != p.W || l.cout_pad != p.cout_pad || l.out.stats == nullptr)
      return cudaErrorInvalidConfiguration;
    L.tmap0 = l.tmap0; L.tmap1 = l.tmap1; L.tmap2 = l.tmap2; L.tmap3 = l.tmap3;
    L.wpacked = l.wpacked; L.bias = l.bias; L.residual = l.residual;
    L.out = l.out.ptr; L.stats = l.out.stats; L.slots = l.out.slots; L.scale = l.scale;
    L.nk = l.cin_pad / Tr::KCH;
    L.ksplit = l.in1.ptr ? l.in0.C / Tr::KCH : L.nk;
    if (l.sk0.ptr != nullptr) {
      if (l.w2packed == nullptr || l.cin2_pad % Tr::SK_KCH != 0 || l.sk0.C % Tr::SK_KCH != 0) return cudaErrorInvalidValue;
      L.w2packed = l.w2packed; L.bias2 = l.bias2;
      L.nk2 = l.cin2_pad / Tr::SK_KCH;
      L.ksplit2 = l.sk1.ptr ? l.sk0.C / Tr::SK_KCH : L.nk2;
      any_skip = true;
    }
    XformParams& x = L.xf;
    x.enabled = 1; x.silu = l.xf.silu;
    x.stats0 = l.in0.stats; x.C0 = l.xf.c0_real > 0 ? l.xf.c0_real : l.in0.C; x.slots0 = l.in0.slots;
    x.stats1 = l.in1.ptr ? l.in1.stats : nullptr; x.C1 = l.in1.ptr ? l.in1.C : 0; x.slots1 = l.in1.slots;
    x.gamma = l.xf.gamma; x.beta = l.xf.beta; x.film = l.xf.film;
    x.film_stride = l.xf.film_stride; x.film_off = l.xf.film_off;
    x.step_ptr = l.xf.step_ptr; x.rows_per_step = l.xf.rows_per_step; x.row_batch_stride = l.xf.row_batch_stride;
    x.groups = l.xf.groups; x.eps = l.xf.eps;
    if (x.C0 + x.C1 > kMaxCin || x.stats0 == nullptr || x.C0 + x.C1 != l.cin_pad) return cudaErrorInvalidValue;
    coef_ch = std::max(coef_ch, (x.C0 + x.C1 + 31) / 32 * 32);
  }
  p.coef_ch = coef_ch;
  p.coef_bytes = 2 * coef_ch * static_cast<int>(sizeof(float));
  p.stage_bytes = Tr::A_BYTES_AL + Tr::B_BYTES;
  if (any_skip && Tr::SK_A_BYTES + Tr::SK_B_BYTES > p.stage_bytes) return cudaErrorInvalidConfiguration;
  int stages = (kBudget - 256 - p.coef_bytes) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  { const int cap = get_option("max_stages", kMaxStages); if (stages > cap) stages = cap; }
  if (stages > 2) stages &= ~1;    // one transform group per ring slot (see conv_umma.cu)
  if (stages < 2) return cudaErrorInvalidConfiguration;
  p.stages = stages;
  p.split = batch_split;
  p.done = done;
  const int smem = 256 + p.coef_bytes + stages * p.stage_bytes;
  int grid = conv_num_sms();
  const int min_group = std::min(p.split, p.B - p.split > 0 ? p.B - p.split : p.split) * p.tiles_per_image;
  if (grid > min_group && !get_option("chain_noclamp", 0)) grid = min_group;       // every CTA gets a tile in every segment
  p.rot = get_option("chain_rot", 61);
  cudaError_t e = cudaMemsetAsync(done, 0, static_cast<size_t>(n) * p.B * sizeof(int), s);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // the counters are cleared by a memset node right before this launch: no programmatic overlap with it
  cfg.attrs = attr; cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

bool conv_chain_supported(const ConvLaunch& l) {
  return l.taps == 9 && l.nt == 128 && (l.ht == 1 || l.ht == 2) && l.xf.enabled && !l.out_nchw && l.out.stats != nullptr;
}

cudaError_t conv_chain_launch(const ConvLaunch* const* ls, int n, int* done, int batch_split, cudaStream_t s) {
  if (n < 1 || n > kMaxChainLayers) return cudaErrorInvalidValue;
  const ConvLaunch& f = *ls[0];
  if (!conv_chain_supported(f)) return cudaErrorInvalidConfiguration;
  if (f.dtype == kBF16) {
    if (f.ht == 2) return launch_chain<__nv_bfloat16, 128, 2>(ls, n, done, batch_split, s);
    return launch_chain<__nv_bfloat16, 128, 1>(ls, n, done, batch_split, s);
  }
  if (f.ht == 2) return launch_chain<float, 128, 2>(ls, n, done, batch_split, s);
  return launch_chain<float, 128, 1>(ls, n, done, batch_split, s);
}

}  // namespace r2dm
