// CTA-pair variant of the ring 3x3 convolution (thread-block cluster of two CTAs, tcgen05.mma.cta_group::2).
// See conv_umma.cu for the single-CTA kernel, the layouts and the role structure this one shares.
#include "conv_common.cuh"

namespace r2dm {

// ====================================================================================== CTA pairs
// 3x3 ring convolution for layers with >= 256 output channels (bf16): a thread-block cluster of two CTAs
// computes a tile of 2 rows x 128 px x 256 output channels with tcgen05.mma.cta_group::2 (M = 256: one
// image row per CTA, N = 256).  Each CTA loads, transforms and holds only ITS row of the halo tile and
// only ITS half of the weights (128 of the 256 output channels); the tensor cores of the two SMs read the
// other half from the peer's shared memory.  Against the single-CTA kernel (2 rows x 128 channels per CTA)
// an output row needs 3 instead of 4 transformed input rows per 256 channels, each MMA reads half as many
// B bytes from local shared memory and no taps have to be fused.  Only the leader (rank 0) issues MMAs:
//   full_bar / empty_bar / acc_full : local to each CTA (TMA completion; multicast tcgen05.commit)
//   xf_bar / acc_empty              : the LEADER's, arrived on remotely by the peer's warps
// Two shapes: <HT = 1, NP = 256> (one row x 256 channels per CTA, four ring slots) and <HT = 2, NP = 128>
// (two rows x 128 channels per CTA; a slot is only 35 KB because each CTA holds 64 of the 128 weight rows,
// so the ring is six slots deep).  Both are slower than the single-CTA kernel in round 1.
template <int HT_, int NP_>
struct PairTr {
  static constexpr int HT = HT_;
  static constexpr int CW = 8, PLANES = 2, AROWS = HT + 2, APITCH = 130;
  static constexpr int A_PLANE_BYTES = AROWS * APITCH * 16;           // 6240
  static constexpr int A_BYTES = PLANES * A_PLANE_BYTES;              // 12480
  static constexpr int A_BYTES_AL = (A_BYTES + 127) / 128 * 128;      // 12544
  static constexpr int NP = NP_;                                      // N of the pair MMA
  static constexpr int NC = NP / 2;                                   // output channels held per CTA
  static constexpr int B_PLANE_BYTES = NC * 16;                       // 2048
  static constexpr int B_TAP_BYTES = PLANES * B_PLANE_BYTES;          // 4096
  static constexpr int B_BYTES = 9 * B_TAP_BYTES;                     // 36864
  static constexpr int STAGE_BYTES = A_BYTES_AL + B_BYTES;            // 49408
  static constexpr int ACC_COLS = HT * NP;                            // accumulator columns per CTA and buffer
  static constexpr int STAGES = (214 * 1024 - 256) / STAGE_BYTES > 6 ? 6 : (214 * 1024 - 256) / STAGE_BYTES;
  static_assert(2 * ACC_COLS <= 512 && STAGES >= 3, "pair tile does not fit");
};

template <int HT, int NP>
__global__ void __launch_bounds__(kConvThreads, 1) conv_pair_kernel(const __grid_constant__ ConvParams p) {
  using T = __nv_bfloat16;
  using Tr = PairTr<HT, NP>;
  constexpr int kPairStages = Tr::STAGES;
  constexpr int CW = Tr::CW;
  constexpr int NT = Tr::NP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[kPairStages], empty_bar[kPairStages], xf_bar[kPairStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float stat_w[2][8][NT / 8][2];
  __shared__ float coef_s[2][kMaxCin];
  __shared__ float grp_s[2][kNU];
  __shared__ __align__(16) float bias_s[NT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int t_begin = static_cast<int>(static_cast<long long>(cid) * p.tiles_total / ncl);
  const int t_end = static_cast<int>(static_cast<long long>(cid + 1) * p.tiles_total / ncl);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1);
      mbar_init(&xf_bar[i], p.xf.enabled ? 8 : 2);      // 4 transform warps (or one relay lane) per CTA
    }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 16); mbar_init(&acc_empty[1], 16);   // 8 epilogue warps per CTA
    fence_mbar_init();
    tma_prefetch_desc(&p.tmap0);
    if (p.ksplit < p.nk) tma_prefetch_desc(&p.tmap1);
  }
  if (warp == 2) tmem_alloc_pair<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anybody arrives on them remotely
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (p.trace != nullptr && blockIdx.x == p.trace_block && threadIdx.x == 0 && p.trace_cap >= 8) {
    p.trace[p.trace_cap - 4] = static_cast<unsigned long long>(clock64());
    p.trace[p.trace_cap - 3] = gtime();
  }
  pdl_launch_dependents();

  auto decode = [&](int t, int& b, int& y, int& xt, int& nt) {
    nt = t % p.ntiles; t /= p.ntiles;
    xt = t % p.xtiles; t /= p.xtiles;
    y = HT * (2 * (t % p.ytiles) + static_cast<int>(rank));   // ytiles = H / (2 HT); first row of this CTA's tile
    b = t / p.ytiles;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (both CTAs)
    if (lane == 0) {
      pdl_wait();
      int st = 0; uint32_t ph = 0, pit = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int b, y, xt, nt;
        decode(t, b, y, xt, nt);
        const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) +
                              (static_cast<size_t>(nt) * p.nk * 2 + rank) * Tr::B_BYTES;
        for (int ks = 0; ks < p.nk; ++ks, st = (st + 1 == kPairStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_relaxed(&empty_bar[st], ph ^ 1, 2000);
          uint8_t* sa = smem_ring + static_cast<size_t>(st) * Tr::STAGE_BYTES;
          mbar_expect_tx(&full_bar[st], Tr::A_BYTES + Tr::B_BYTES);
          const bool second = ks >= p.ksplit;
          const int plane0 = (second ? ks - p.ksplit : ks) * Tr::PLANES;
          tma_load_5d(sa, second ? &p.tmap1 : &p.tmap0, &full_bar[st], 2 * xt * 128, 0, y - 1, plane0, b);
          bulk_load(sa + Tr::A_BYTES_AL, wsrc + static_cast<size_t>(ks) * 2 * Tr::B_BYTES, Tr::B_BYTES, &full_bar[st]);
          R2DM_TRACE(0, pit); ++pit;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc(256, NT, Elem<T>::kFmt);
      const uint32_t a_lo_const = static_cast<uint32_t>(Tr::A_PLANE_BYTES >> 4) << 16;
      const uint32_t b_lo_const = static_cast<uint32_t>(Tr::B_PLANE_BYTES >> 4) << 16;
      constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
      int st = 0; uint32_t ph = 0, mit = 0;
      int j = 0;
      for (int t = t_begin; t < t_end; ++t, ++j) {
        const int buf = j & 1;
        mbar_wait_cluster(&acc_empty[buf], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dbase = tmem + buf * Tr::ACC_COLS;
        for (int ks = 0; ks < p.nk; ++ks, st = (st + 1 == kPairStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_cluster(&xf_bar[st], ph);
          tc_fence_after();
          if (lane == 0) R2DM_TRACE(1, 2 * mit);
          const uint32_t sa = smem_u32(smem_ring + static_cast<size_t>(st) * Tr::STAGE_BYTES);
          const uint32_t a_lo0 = a_lo_const | ((sa >> 4) & 0x3FFFu);
          const uint32_t b_lo0 = b_lo_const | (((sa + Tr::A_BYTES_AL) >> 4) & 0x3FFFu);
          if (!(p.debug & 2)) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap % 3;
              const uint64_t bdesc = (static_cast<uint64_t>(kHi) << 32) |
                                     (b_lo0 + static_cast<uint32_t>((tap * Tr::B_TAP_BYTES) >> 4));
#pragma unroll
              for (int r = 0; r < HT; ++r) {
                const uint64_t adesc = (static_cast<uint64_t>(kHi) << 32) |
                                       (a_lo0 + static_cast<uint32_t>((((r + dy) * Tr::APITCH + dx) * 16) >> 4));
                umma_f16_pair_warp(dbase + r * NT, adesc, bdesc, idesc, (ks > 0 || tap > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit_pair_warp(&empty_bar[st]);     // frees this stage in BOTH CTAs
          if (lane == 0) R2DM_TRACE(1, 2 * mit + 1);
          ++mit;
        }
        umma_commit_pair_warp(&acc_full[buf]);       // accumulators of both CTAs are complete
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ relay (plain convolution)
    // without the transform nobody in this CTA touches the landed stage: forward "my stage has landed"
    // to the leader's barrier
    if (!p.xf.enabled && lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t)
        for (int ks = 0; ks < p.nk; ++ks, st = (st + 1 == kPairStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_relaxed(&full_bar[st], ph, 500);
          mbar_arrive_cluster(mapa_u32(&xf_bar[st], 0));
        }
    }
  } else if (warp >= 4 && warp < kEpiWarp0) {
    // ------------------------------------------------------------------ operand transform (both CTAs)
    if (p.xf.enabled) {
      pdl_wait();
      const int grp = (warp - 4) >> 2;
      const int t256 = threadIdx.x - 128;
      const int tt = t256 & 127;
      constexpr int TPP = 128 / Tr::PLANES;
      const int my_plane = tt / TPP, tip = tt % TPP;
      const int Ctot = p.xf.C0 + p.xf.C1;
      const int gsize = Ctot / p.xf.groups;
      uint32_t it = 0, ph = 0;
      int st = 0;
      int cur_b = -1;
      for (int t = t_begin; t < t_end; ++t) {
        int b, y, xt, nt;
        decode(t, b, y, xt, nt);
        if (b != cur_b) {
          cur_b = b;
          const float* fl = nullptr;
          if (p.xf.film != nullptr) {
            const int row = (p.xf.step_ptr ? *p.xf.step_ptr : 0) * p.xf.rows_per_step + b * p.xf.row_batch_stride;
            fl = p.xf.film + static_cast<size_t>(row) * p.xf.film_stride + p.xf.film_off;
          }
          constexpr int CPT = kMaxCin / 256;
          float ga_r[CPT], be_r[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) {
            const int c = t256 + k * 256;
            ga_r[k] = 0.f; be_r[k] = 0.f;
            if (c < Ctot) {
              ga_r[k] = fl ? 1.f + fl[c] : p.xf.gamma[c];
              be_r[k] = fl ? fl[Ctot + c] : p.xf.beta[c];
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (grp == 0) {
            const int g = tt >> 4, l16 = tt & 15;
            double s1 = 0.0, s2 = 0.0;
            const int lo = g * gsize, hi_c = lo + gsize;
            int off = 0;
            for (int si = 0; si < 2; ++si) {
              const int Cs = si == 0 ? p.xf.C0 : p.xf.C1;
              if (Cs == 0) break;
              const float* stp = si == 0 ? p.xf.stats0 : p.xf.stats1;
              const int sl = si == 0 ? p.xf.slots0 : p.xf.slots1;
              const int a = max(lo, off), e = min(hi_c, off + Cs);
              if (a < e) {
                const int unit_ch = Cs / kNU;
                const int u0 = (a - off) / unit_ch, u1 = (e - off) / unit_ch;
                const int n = (u1 - u0) * sl;
                const float2* st2 = reinterpret_cast<const float2*>(stp + (static_cast<size_t>(b) * kNU + u0) * sl * 2);
                for (int i = l16; i < n; i += 16) { const float2 v = st2[i]; s1 += v.x; s2 += v.y; }
              }
              off += Cs;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              s1 += __shfl_xor_sync(0xffffffffu, s1, o);
              s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (l16 == 0) {
              const double cnt = static_cast<double>(gsize) * p.H * p.W;
              const double mean = s1 / cnt;
              double var = s2 / cnt - mean * mean;
              if (var < 0.0) var = 0.0;
              grp_s[0][g] = static_cast<float>(mean);
              grp_s[1][g] = rsqrtf(static_cast<float>(var) + p.xf.eps);
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          const float fold = p.xf.silu ? 0.5f : 1.f;
#pragma unroll
          for (int k = 0; k < CPT; ++k) {
            const int c = t256 + k * 256;
            if (c < Ctot) {
              const int g = c / gsize;
              const float a = grp_s[1][g] * ga_r[k];
              coef_s[0][c] = a * fold;
              coef_s[1][c] = (be_r[k] - grp_s[0][g] * a) * fold;
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        const int y_first = y - 1;
        const int row_lo = max(0, -y_first), row_hi = min(Tr::AROWS, p.H - y_first);
        const int n_units = (row_hi - row_lo) * Tr::APITCH;
        for (int ks = 0; ks < p.nk; ++ks, ++it, st = (st + 1 == kPairStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
          float ca[CW], cd[CW];
          const int c0 = (ks * Tr::PLANES + my_plane) * CW;
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const bool ok = c0 + i < Ctot;
            ca[i] = ok ? coef_s[0][c0 + i] : 0.f;
            cd[i] = ok ? coef_s[1][c0 + i] : 0.f;
          }
          mbar_wait_relaxed(&full_bar[st], ph, 500);
          if (tt == 0) R2DM_TRACE(grp ? 4 : 2, 2 * it);
          const uint32_t sbase = smem_u32(smem_ring + static_cast<size_t>(st) * Tr::STAGE_BYTES +
                                          my_plane * Tr::A_PLANE_BYTES) + row_lo * Tr::APITCH * 16;
          auto xform_unit = [&](uint4 raw) {
            float v[CW];
            Elem<T>::unpack(raw, v);
#pragma unroll
            for (int k = 0; k < CW; ++k) {
              const float tv = fmaf(v[k], ca[k], cd[k]);
              v[k] = p.xf.silu ? silu_from_half(tv) : tv;
            }
            return Elem<T>::pack_mma(v);
          };
          if (c0 < Ctot && p.xf.debug == 0) {
            constexpr int XB = HT == 1 ? 3 : 4;   // 390 (520) units per plane / 64 threads: full batches, then a remainder
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
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(&xf_bar[st], 0));
          if (tt == 0) R2DM_TRACE(grp ? 4 : 2, 2 * it + 1);
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own row)
    const int ew = warp - kEpiWarp0;
    const int q = ew & 3;
    const int half = ew >> 2;                 // HT == 1: which half of the columns; HT == 2: which row
    const int m = q * 32 + lane;
    const int Wp = p.W + 2;
    const int planes_out = p.cout_pad / CW;
    const size_t plane_stride = static_cast<size_t>(p.H) * Wp;
    constexpr int CB = 32, NCHUNK = (HT == 1 ? NT / 2 : NT) / CB, NSUB = CB / 8;
    const int c_begin = HT == 1 ? half * (NT / 2) : 0;
    const int r_mine = HT == 1 ? 0 : half;
    const uint4* res = static_cast<const uint4*>(p.residual);
    uint4* out = static_cast<uint4*>(p.out);
    const int ethread = threadIdx.x - kEpiWarp0 * 32;
    pdl_wait();
    int j = 0, cur_nt = -1;
    for (int t = t_begin; t < t_end; ++t, ++j) {
      int b, y, xt, nt;
      decode(t, b, y, xt, nt);
      const int n0 = nt * NT, x = xt * 128 + m;
      const int buf = j & 1, par = j & 1;
      if (nt != cur_nt) {
        cur_nt = nt;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = ethread; i < NT; i += 256) bias_s[i] = p.bias[n0 + i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait_relaxed(&acc_full[buf], (j >> 1) & 1, 1000);
      tc_fence_after();
      if (ethread == 0) R2DM_TRACE(3, 3 * j);
      const uint32_t tbase = tmem + buf * Tr::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < ((p.debug & 1) ? 0 : NCHUNK); ++ch) {
        const int c0 = c_begin + ch * CB;
        float ssum[NSUB][2];
#pragma unroll
        for (int u = 0; u < NSUB; ++u) { ssum[u][0] = 0.f; ssum[u][1] = 0.f; }
        const size_t idx0 = pt_index(b, planes_out, (n0 + c0) / CW, p.H, Wp, y + r_mine, x + 1);
        uint4 rr[CB / CW];
        if (res != nullptr) {
#pragma unroll
          for (int u = 0; u < CB / CW; ++u) rr[u] = res[idx0 + u * plane_stride];
        }
        float v[CB];
#pragma unroll
        for (int h16 = 0; h16 < CB / 16; ++h16) tmem_ld16(tbase + r_mine * NT + c0 + h16 * 16, v + h16 * 16);
        tmem_ld_wait();
        const float4* bias4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
        for (int i4 = 0; i4 < CB / 4; ++i4) {
          const float4 bv = bias4[i4];
          v[4 * i4] += bv.x; v[4 * i4 + 1] += bv.y; v[4 * i4 + 2] += bv.z; v[4 * i4 + 3] += bv.w;
        }
#pragma unroll
        for (int u = 0; u < CB / CW; ++u) {
          const size_t idx = idx0 + u * plane_stride;
          float o[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) o[i] = v[u * CW + i];
          if (res != nullptr) {
            float rv[CW];
            Elem<T>::unpack(rr[u], rv);
#pragma unroll
            for (int i = 0; i < CW; ++i) o[i] += rv[i];
          }
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < CW; ++i) { o[i] *= p.scale; s1 += o[i]; s2 = fmaf(o[i], o[i], s2); }
          ssum[u][0] += s1;
          ssum[u][1] += s2;
          const uint4 pk = Elem<T>::pack(o);
          out[idx] = pk;
          if (x == 0) out[idx + p.W] = pk;
          if (x == p.W - 1) out[idx - p.W] = pk;
        }
        if (p.stats != nullptr) {
          float vals[2 * NSUB];
#pragma unroll
          for (int u = 0; u < NSUB; ++u) { vals[2 * u] = ssum[u][0]; vals[2 * u + 1] = ssum[u][1]; }
#pragma unroll
          for (int rd = 0; rd < 3; ++rd) {
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
          for (int off = 2; off > 0; off >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
          if ((lane & 3) == 0) {
            const int vi = lane >> 2;
            stat_w[par][ew][(c0 >> 3) + (vi >> 1)][vi & 1] = vals[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[buf], 0));
      if (ethread == 0) R2DM_TRACE(3, 3 * j + 1);
      if (p.stats != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int units_here = NT / p.unit_ch;
        if (ethread < units_here * 2) {
          const int u = ethread >> 1, kk = ethread & 1;
          const int cpu = p.unit_ch >> 3;
          float tot = 0.f;
          for (int sc = u * cpu; sc < (u + 1) * cpu; ++sc) {
#pragma unroll
            for (int w = 0; w < 8; ++w)
              if (HT > 1 || (w >> 2) == (sc >= NT / 16 ? 1 : 0)) tot += stat_w[par][w][sc][kk];
          }
          const int unit = n0 / p.unit_ch + u;
          const int slot = (y / HT) * p.xtiles + xt;
          p.stats[((static_cast<size_t>(b) * kNU + unit) * p.slots + slot) * 2 + kk] = tot;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still read / signal here
  if (p.trace != nullptr && blockIdx.x == p.trace_block && threadIdx.x == 0 && p.trace_cap >= 8) {
    p.trace[p.trace_cap - 2] = static_cast<unsigned long long>(clock64());
    p.trace[p.trace_cap - 1] = gtime();
  }
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

template <int HT, int NP>
static cudaError_t launch_pair(const ConvLaunch& l, cudaStream_t s) {
  using Tr = PairTr<HT, NP>;
  constexpr int kPairStages = Tr::STAGES;
  if (l.dtype != kBF16 || l.taps != 9 || l.ht != HT || l.out_nchw != nullptr || l.out.H % (2 * HT) != 0 || l.cout_pad % NP != 0)
    return cudaErrorInvalidConfiguration;
  auto kern = conv_pair_kernel<HT, NP>;
  const int smem = 256 + kPairStages * Tr::STAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  using T = __nv_bfloat16;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tmap0 = l.tmap0; p.tmap1 = l.tmap1;
  p.wpacked = l.wpacked; p.bias = l.bias; p.residual = l.residual;
  p.out = l.out.ptr; p.stats = l.out.stats;
  p.B = l.out.B; p.H = l.out.H; p.W = l.out.W;
  p.cout = l.cout; p.cout_pad = l.cout_pad;
  p.nk = l.cin_pad / 16;
  p.ksplit = l.in1.ptr ? l.in0.C / 16 : p.nk;
  p.xtiles = l.out.W / 128; p.ytiles = l.out.H / (2 * HT); p.ntiles = l.cout_pad / NP;
  p.tiles_total = p.B * p.ytiles * p.xtiles * p.ntiles;      // tiles of the PAIR (2 HT rows x 128 px x NP channels)
  p.unit_ch = l.cout / kNU >= 8 ? l.cout / kNU : 8;
  p.unit_shift = 0;
  while ((1 << p.unit_shift) < p.unit_ch) ++p.unit_shift;
  if (p.stats != nullptr && ((1 << p.unit_shift) != p.unit_ch || p.unit_ch > NP)) return cudaErrorInvalidValue;
  p.slots = l.out.slots;
  p.scale = l.scale;
  if (l.xf.enabled) {
    XformParams& x = p.xf;
    x.enabled = 1; x.silu = l.xf.silu;
    x.stats0 = l.in0.stats; x.C0 = l.xf.c0_real > 0 ? l.xf.c0_real : l.in0.C; x.slots0 = l.in0.slots;
    x.stats1 = l.in1.ptr ? l.in1.stats : nullptr; x.C1 = l.in1.ptr ? l.in1.C : 0; x.slots1 = l.in1.slots;
    x.gamma = l.xf.gamma; x.beta = l.xf.beta; x.film = l.xf.film;
    x.film_stride = l.xf.film_stride; x.film_off = l.xf.film_off;
    x.step_ptr = l.xf.step_ptr; x.rows_per_step = l.xf.rows_per_step; x.row_batch_stride = l.xf.row_batch_stride;
    x.groups = l.xf.groups; x.eps = l.xf.eps;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("R2DM_XF_DEBUG"); dbg = e ? atoi(e) : 0; }
    x.debug = dbg;
    if (x.C0 + x.C1 > kMaxCin || x.stats0 == nullptr) return cudaErrorInvalidValue;
  }
  p.stages = kPairStages; p.stage_bytes = Tr::STAGE_BYTES;
  p.trace = conv_trace_for_this_launch(); p.trace_cap = conv_trace_cap();
  { const char* e = getenv("R2DM_TRACE_BLOCK"); p.trace_block = e ? static_cast<unsigned>(atoi(e)) : 0u; }
  {
    static int cdbg = -1;
    if (cdbg < 0) { const char* e = getenv("R2DM_CONV_DEBUG"); cdbg = e ? atoi(e) : 0; }
    p.debug = cdbg;
  }
  int grid = conv_num_sms() & ~1;
  if (grid > 2 * p.tiles_total) grid = 2 * p.tiles_total;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("R2DM_PDL"); pdl = e ? atoi(e) : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  (void)sizeof(T);
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t conv_pair_launch(const ConvLaunch& l, cudaStream_t s) {
  if (l.nt == 256) return launch_pair<1, 256>(l, s);
  if (l.pair && l.nt == 128) return launch_pair<2, 128>(l, s);
  return cudaErrorInvalidConfiguration;
}

}  // namespace r2dm
