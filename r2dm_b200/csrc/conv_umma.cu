// Ring 3x3 / 1x1 convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a), with the
// preceding GroupNorm / AdaGN (+SiLU) fused into the operand path.
//
// Replaces, on the sampling path, models/ops.py:149-173 (Conv2d + Pad: circular azimuth padding,
// zero elevation padding, 3x3 cross-correlation + bias), the 1x1 skip convolution of
// models/efficient_unet.py:87-91, the nn.MultiheadAttention in/out projections (:34-38), and the
// nn.GroupNorm / ops.AdaGN + nn.SiLU that feed them (efficient_unet.py:72-81,99-106; ops.py:176-200).
//
// Persistent kernel: one CTA per SM walks a contiguous range of output tiles (HT rows x 128 pixels
// x NT output channels).  Warp roles (640 threads):
//   warp 0      producer.  Per pipeline stage (KCH input channels) one 5-D TMA box load brings the
//               halo tile [(HT+2) x 130 px] of those channels into shared memory (planar-16 layout,
//               common.cuh) and one bulk copy brings the pre-packed weights of all taps for those
//               channels - or, when the whole filter bank fits (Cin = Cout = 64), the weights are
//               loaded once and stay resident.  The producer runs ahead across tile boundaries.
//   warp 1      MMA issuer.  For every output row and tap one tcgen05.mma whose A descriptor points
//               at the *shifted* start address inside the halo tile (so the input is read from L2
//               once per tile, not once per tap).  Accumulators: 2 x (HT x NT) fp32 TMEM columns,
//               double buffered so the epilogue of tile j overlaps the MMAs of tile j+1.
//   warp 2      TMEM allocator.
//   warps 4-11  operand transform (optional; two groups of four warps on alternating stages).  Applies y = silu(a_c x + d_c) in place on the freshly
//               landed stage, where (a_c, d_c) fold the GroupNorm statistics left by the producer
//               kernel, the affine / FiLM parameters and the normalisation; rows outside the image
//               stay zero (the conv's zero padding applies to the *normalised* tensor).
//   warps 12-19 epilogue.  tcgen05.ld -> + bias (+ residual) -> * scale -> GroupNorm partial sums
//               for the consumer -> bf16/fp32 planar-16 store (incl. the wrap halo columns), or fp32
//               NCHW store for the network output.
// Elevation borders come from TMA out-of-bounds zero fill; the azimuth wrap from the halo columns.
#include "conv_common.cuh"

// SiLU in the operand transform: h + h tanh(h), h = t / 2, with tanh.approx (one MUFU, 2^-11 relative) for BOTH engines:
// the result is rounded to bf16 (2^-9) or tf32 (2^-11) right after.  -DR2DM_F32_TANH=0 restores x / (1 + exp(-x)) in the
// fp32 engine (measured: same trajectory errors, forward 2.72 vs 2.62 ms at B=4).
#ifndef R2DM_F32_TANH
#define R2DM_F32_TANH 1
#endif

namespace r2dm {

__device__ __forceinline__ bool get_prefetch(const ConvParams& p) { return p.prefetch_w != 0; }

// silu(t) = h + h tanh(h) with h = t/2 (the 1/2 is folded into the affine coefficients): one MUFU op
// PW: point-wise MLP extras of the 1x1 kernel (ReLU, per-channel maximum over pixels; pointnet.cu) - a separate
// instantiation so that the network's own kernels keep their register allocation
template <typename T, int NT, int HT, int TAPS, int KS, bool NCHW, bool PW>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
  using Tr = ConvTraits<T, NT, HT, TAPS, KS>;
  constexpr int CW = Tr::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], xf_bar[kMaxStages];
  __shared__ uint64_t acc_full[2], acc_empty[2], wres_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float stat_w[2][8][NT / 8 > 0 ? NT / 8 : 1][2];   // [tile parity][epilogue warp][8-channel sub-chunk][sum, sumsq]
  __shared__ float grp_s[2][kNU];
  __shared__ __align__(16) float bias_s[NT];
  // PW: running per-channel maximum of the image this CTA currently works on (all N tiles), flushed to global
  // memory with one atomic per channel when the image changes - not one per tile
  __shared__ float pw_max[PW ? kMaxPwChannels : 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.ktime != nullptr && threadIdx.x == 0) atomicMin(p.ktime, gtime());
  const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.tiles_total / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.tiles_total / gridDim.x);
  const uint32_t wres_bytes = p.wres ? static_cast<uint32_t>(p.nk) * Tr::B_BYTES : 0u;
  // dynamic shared memory: [transform coefficients a_c | d_c (2 x coef_ch floats)] [resident weights] [ring]
  float* coef_a = reinterpret_cast<float*>(smem);
  float* coef_d = coef_a + p.coef_ch;
  uint8_t* smem_w = smem + p.coef_bytes;  // resident weights (if any)
  uint8_t* smem_ring = smem_w + wres_bytes; // stage ring

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&xf_bar[i], 4); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
    mbar_init(&wres_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmap0);
    if (p.ksplit < p.nk) tma_prefetch_desc(&p.tmap1);
    if (p.nk2 > 0) { tma_prefetch_desc(&p.tmap2); if (p.ksplit2 < p.nk2) tma_prefetch_desc(&p.tmap3); }
  }
  if (warp == kAllocWarp) tmem_alloc<Tr::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (R2DM_DBG(p.trace != nullptr && blockIdx.x == p.trace_block && threadIdx.x == 0 && p.trace_cap >= 8)) {
    p.trace[p.trace_cap - 4] = static_cast<unsigned long long>(clock64());   // (clock, globaltimer) at start
    p.trace[p.trace_cap - 3] = gtime();
  }
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) and the resident-weight preload below overlap the tail of the previous kernel; no
  // activation / statistics / output address is touched before pdl_wait().
  pdl_launch_dependents();

  // p.reverse: this launch walks the tiles back to front, so that it starts with what the previous launch
  // wrote last (still in L2) instead of with what it wrote first (evicted when a tensor exceeds L2)
  auto decode = [&](int t, int& b, int& yt, int& xt, int& nt) {
    if (p.reverse) t = p.tiles_total - 1 - t;
    nt = t % p.ntiles; t /= p.ntiles;
    xt = t % p.xtiles; t /= p.xtiles;
    yt = t % p.ytiles;
    b = t / p.ytiles;
  };

  if (warp == kProdWarp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      if (p.wres) {
        mbar_expect_tx(&wres_bar, wres_bytes);
        for (int ks = 0; ks < p.nk; ++ks)
          bulk_load(smem_w + static_cast<size_t>(ks) * Tr::B_BYTES,
                    static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(ks) * Tr::B_BYTES, Tr::B_BYTES,
                    &wres_bar);
      }
      // (option prefetch_w, measured +-0) Streamed weights do not depend on the previous kernel either: the weight part of the first ring fill is
      // requested before the dependency wait (a CTA becomes resident as soon as its SM is free and then sits in
      // griddepcontrol.wait until the slowest CTA of the previous launch has finished), so that only the activation
      // boxes are left to fetch once the wait returns.  A complete_tx that lands before the matching
      // arrive.expect_tx is fine: the phase cannot complete while the arrival is pending.
      uint32_t pre = 0;
      if (!p.wres && t_begin < t_end && get_prefetch(p)) {
        int b, yt, xt, nt;
        decode(t_begin, b, yt, xt, nt);
        const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(nt) * p.nk * Tr::B_BYTES;
        pre = static_cast<uint32_t>(p.stages < p.nk ? p.stages : p.nk);
        for (uint32_t ks = 0; ks < pre; ++ks)
          bulk_load(smem_ring + static_cast<size_t>(ks) * p.stage_bytes + Tr::A_BYTES_AL,
                    wsrc + static_cast<size_t>(ks) * Tr::B_BYTES, Tr::B_BYTES, &full_bar[ks]);
      }
      pdl_wait();
      const uint64_t l2pol = p.l2_hint ? l2_policy_evict_first() : 0ull;
      uint32_t it = 0, ph = 0;
      int st = 0;   // ring slot / phase advance incrementally (no division per stage)
      for (int t = t_begin; t < t_end; ++t) {
        int b, yt, xt, nt;
        decode(t, b, yt, xt, nt);
        const int x0 = xt * 128, y0 = yt * HT;
        const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(nt) * p.nk * Tr::B_BYTES;
        for (int ks = 0; ks < p.nk; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait_relaxed(&empty_bar[st], ph ^ 1, 2000);
          uint8_t* sa = smem_ring + static_cast<size_t>(st) * p.stage_bytes;
          if (R2DM_DBG(p.debug & 4)) { mbar_arrive(&full_bar[st]); continue; }
          mbar_expect_tx(&full_bar[st], Tr::A_BYTES + (p.wres ? 0 : Tr::B_BYTES));
          const bool second = ks >= p.ksplit;
          const int plane0 = (second ? ks - p.ksplit : ks) * Tr::PLANES;
          const CUtensorMap* tm = second ? &p.tmap1 : &p.tmap0;
          // coordinates: (8-byte element within the padded row, half, row, plane, image)
          if (p.l2_hint) {
            // experiment (option l2_evict_first): activations are streamed once per launch - let them leave L2 first so
            // that the output this launch writes is still resident when the next launch reads it
            if (TAPS == 9) tma_load_5d_hint(sa, tm, &full_bar[st], 2 * x0, 0, y0 - 1, plane0, b, l2pol);
            else tma_load_5d_hint(sa, tm, &full_bar[st], 2 * (x0 + 1), 0, y0, plane0, b, l2pol);
          } else if (TAPS == 9) tma_load_5d(sa, tm, &full_bar[st], 2 * x0, 0, y0 - 1, plane0, b);
          else tma_load_5d(sa, tm, &full_bar[st], 2 * (x0 + 1), 0, y0, plane0, b);
          if (!p.wres && it >= pre)
            bulk_load(sa + Tr::A_BYTES_AL, wsrc + static_cast<size_t>(ks) * Tr::B_BYTES, Tr::B_BYTES, &full_bar[st]);
          R2DM_TRACE(0, it);
        }
        if constexpr (TAPS == 9) {
          // skip stages: centre pixels of the raw skip input(s) + their 1x1 weight rows
          const uint8_t* w2src = static_cast<const uint8_t*>(p.w2packed) + static_cast<size_t>(nt) * p.nk2 * Tr::SK_B_BYTES;
          for (int ks = 0; ks < p.nk2; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
            mbar_wait_relaxed(&empty_bar[st], ph ^ 1, 2000);
            uint8_t* sa = smem_ring + static_cast<size_t>(st) * p.stage_bytes;
            mbar_expect_tx(&full_bar[st], Tr::SK_A_BYTES + Tr::SK_B_BYTES);
            const bool second = ks >= p.ksplit2;
            const int plane0 = (second ? ks - p.ksplit2 : ks) * Tr::SK_PLANES;
            tma_load_5d(sa, second ? &p.tmap3 : &p.tmap2, &full_bar[st], 2 * (x0 + 1), 0, y0, plane0, b);
            bulk_load(sa + Tr::SK_A_BYTES, w2src + static_cast<size_t>(ks) * Tr::SK_B_BYTES, Tr::SK_B_BYTES, &full_bar[st]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // the whole warp runs this loop convergently; one elected lane issues (see umma_f16_warp)
    {
      const uint32_t idesc = make_idesc(128, NT, Elem<T>::kFmt);
      const uint32_t idesc2 = make_idesc(128, 2 * NT <= 256 ? 2 * NT : NT, Elem<T>::kFmt);
      const uint32_t idesc3 = make_idesc(128, 3 * NT <= 256 ? 3 * NT : NT, Elem<T>::kFmt);
      (void)idesc2; (void)idesc3;
      // descriptor halves: lo = start>>4 | LBO>>4 << 16 ; hi = SBO>>4 | version<<14 (no swizzle)
      const uint32_t a_lo_const = static_cast<uint32_t>(Tr::A_PLANE_BYTES >> 4) << 16;
      const uint32_t b_lo_const = static_cast<uint32_t>(Tr::B_PLANE_BYTES >> 4) << 16;
      if (p.wres) mbar_wait(&wres_bar, 0);
      uint32_t it = 0, ph = 0;
      int st = 0;
      int j = 0;
      for (int t = t_begin; t < t_end; ++t, ++j) {
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dbase = tmem + buf * Tr::ACC_COLS;
        for (int ks = 0; ks < p.nk; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait(p.xf.enabled ? &xf_bar[st] : &full_bar[st], ph);
          tc_fence_after();
          if (lane == 0) R2DM_TRACE(1, 2 * it);
          const uint32_t sa = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes);
          const uint32_t sb = p.wres ? smem_u32(smem_w) + static_cast<uint32_t>(ks) * Tr::B_BYTES : sa + Tr::A_BYTES_AL;
          // low descriptor words of the stage bases; every operand below is base + compile-time
          // constant (shared memory < 256 KB, so the 14-bit start-address field cannot carry)
          const uint32_t a_lo0 = a_lo_const | ((sa >> 4) & 0x3FFFu);
          const uint32_t b_lo0 = b_lo_const | ((sb >> 4) & 0x3FFFu);
          constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
          if (R2DM_DBG(p.debug & 2)) {
            // developer ablation: no MMAs issued
          } else if constexpr (Tr::FUSE) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int kk = 0; kk < KS; ++kk) {
                const uint32_t a_k = static_cast<uint32_t>((kk * 2 * Tr::A_PLANE_BYTES + kx * 16) >> 4);
                const uint32_t b_k = static_cast<uint32_t>((kx * Tr::B_TAP_BYTES + kk * 2 * Tr::B_PLANE_BYTES) >> 4);
                // First touch of the accumulators (ks = kx = kk = 0): the MMAs are issued in an order in
                // which the first one(s) cover a partition of the output rows, so those clear the
                // accumulators (accumulate = 0) and every other MMA accumulates: input row
                // min(2, HT-1) feeds rows 0..min(2, HT-1), and for HT = 4 input row 5 feeds row 3 alone.
                static_assert(HT <= 4, "first-touch order assumes at most four output rows per tile");
                constexpr int IR_A = HT - 1 < 2 ? HT - 1 : 2;
                constexpr int IR_B = HT - 1 > IR_A ? HT + 1 : -1;
                const bool first = ks == 0 && kx == 0 && kk == 0;
#pragma unroll
                for (int idx = 0; idx < HT + 2; ++idx) {
                  // issue order: IR_A, IR_B (if any), then the remaining input rows ascending
                  int ir = idx;
                  if (idx == 0) ir = IR_A;
                  else if (IR_B >= 0 && idx == 1) ir = IR_B;
                  else {
                    int k = idx - (IR_B >= 0 ? 2 : 1);   // k-th row of the remaining ones
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
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const int dy = TAPS == 9 ? tap / 3 : 0, dx = TAPS == 9 ? tap % 3 : 0;
#pragma unroll
              for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
                for (int r = 0; r < HT; ++r) {
                  const uint32_t a_add = static_cast<uint32_t>((kk * 2 * Tr::A_PLANE_BYTES + ((r + dy) * Tr::APITCH + dx) * 16) >> 4);
                  const uint32_t b_add = static_cast<uint32_t>((tap * Tr::B_TAP_BYTES + kk * 2 * Tr::B_PLANE_BYTES) >> 4);
                  const uint64_t adesc = (static_cast<uint64_t>(kHi) << 32) | (a_lo0 + a_add);
                  const uint64_t bdesc = (static_cast<uint64_t>(kHi) << 32) | (b_lo0 + b_add);
                  const uint32_t acc = (ks > 0 || tap > 0 || kk > 0) ? 1u : 0u;
                  if (Elem<T>::kFmt == 2) umma_tf32_warp(dbase + r * NT, adesc, bdesc, idesc, acc);
                  else umma_f16_warp(dbase + r * NT, adesc, bdesc, idesc, acc);
                }
              }
            }
          }
          umma_commit_warp(&empty_bar[st]);  // frees this smem stage once the MMAs above have read it
          if (lane == 0) R2DM_TRACE(1, 2 * it + 1);
        }
        if constexpr (TAPS == 9) {
          // skip stages: plain 1x1 MMAs (one per output row and 16-byte K step) into the same accumulators
          for (int ks = 0; ks < p.nk2; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
            mbar_wait(p.xf.enabled ? &xf_bar[st] : &full_bar[st], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes);
            const uint32_t sb = sa + Tr::SK_A_BYTES;
            const uint32_t a_lo0 = (static_cast<uint32_t>(Tr::SK_A_PLANE_BYTES >> 4) << 16) | ((sa >> 4) & 0x3FFFu);
            const uint32_t b_lo0 = (static_cast<uint32_t>(Tr::SK_B_PLANE_BYTES >> 4) << 16) | ((sb >> 4) & 0x3FFFu);
            constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
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
        }
        umma_commit_warp(&acc_full[buf]);
      }
    }
  } else if (warp >= kXfWarp0 && warp < kXfWarp0 + 8) {
    // ------------------------------------------------------------------ operand transform
    // Two groups of four warps; group g transforms the pipeline stages with (global index % 2) == g,
    // so two stages are in flight and the XU (tanh) pipe of every SM sub-partition always has a
    // second warp to issue from while the first waits on shared memory.
    if (p.xf.enabled) {
      const double inv_cnt = 1.0 / (static_cast<double>((p.xf.C0 + p.xf.C1) / p.xf.groups) * p.H * p.W);
      pdl_wait();
      const int grp = (warp - kXfWarp0) >> 2;           // 0 / 1
      const int t256 = threadIdx.x - kXfWarp0 * 32;     // 0..255 over both groups
      const int tt = t256 & 127;                        // 0..127 within the group
      constexpr int TPP = 128 / Tr::PLANES;             // threads per channel plane
      const int my_plane = tt / TPP, tip = tt % TPP;
      const int Ctot = p.xf.C0 + p.xf.C1;
      const int gsize = Ctot / p.xf.groups;
      uint32_t it = 0, ph = 0;
      int st = 0;
      int cur_b = -1;
      for (int t = t_begin; t < t_end; ++t) {
        int b, yt, xt, nt;
        decode(t, b, yt, xt, nt);
        if (b != cur_b && !R2DM_DBG(p.xf.debug & 4)) {   // (debug 4: developer ablation, no statistics fold)
          // ---- fold statistics + affine/FiLM into per-channel (a, d) for image b (all 8 warps; the
          // first barrier also guarantees that nobody still reads the previous image's table)
          cur_b = b;
          if (t256 == 0 && it == 0) R2DM_TRACE(4, 0);   // statistics fold begins
          // affine / FiLM parameters of this thread's channels: loaded first so that their latency
          // overlaps the statistics loads (this fold sits on the critical path of every launch)
          const float* fl = nullptr;
          if (p.xf.film != nullptr) {
            const int row = (p.xf.step_ptr ? *p.xf.step_ptr : 0) * p.xf.rows_per_step + b * p.xf.row_batch_stride;
            fl = p.xf.film + static_cast<size_t>(row) * p.xf.film_stride + p.xf.film_off;
          }
          constexpr int CPT = kMaxCin / 256;             // channels per thread
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
          {
            // one warp per GroupNorm group (groups == 8): every lane issues all its loads before the first add
            const int g = t256 >> 5;
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
                for (int i = lane; i < n; i += 128) {
                  float2 v[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) v[u] = i + 32 * u < n ? st2[i + 32 * u] : make_float2(0.f, 0.f);
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
            // the previous image's table may still be in use until everybody is here
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (lane == 0) {
              const double mean = s1 * inv_cnt;
              double var = s2 * inv_cnt - mean * mean;
              if (var < 0.0) var = 0.0;
              grp_s[0][g] = static_cast<float>(mean);
              grp_s[1][g] = rsqrtf(static_cast<float>(var) + p.xf.eps);
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          // with the fast SiLU the coefficients produce h = t/2 directly
          const float fold = (p.xf.silu && (sizeof(T) == 2 || R2DM_F32_TANH)) ? 0.5f : 1.f;
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
          if (t256 == 0 && it == 0) R2DM_TRACE(4, 1);   // coefficient table ready
        }
        const int y_first = TAPS == 9 ? yt * HT - 1 : yt * HT;   // image row of tile row 0
        const int row_lo = max(0, -y_first), row_hi = min(Tr::AROWS, p.H - y_first);
        const int n_units = (row_hi - row_lo) * Tr::APITCH;
        for (int ks = 0; ks < p.nk; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
          // per-channel coefficients of this thread's plane as fp32 pairs (channels 2i, 2i+1); every plane of a
          // transformed layer is real (the host rejects channel padding together with the transform)
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
          if (tt == 0) R2DM_TRACE(grp ? 4 : 2, 2 * it);
          const uint32_t sbase = smem_u32(smem_ring + static_cast<size_t>(st) * p.stage_bytes +
                                          my_plane * Tr::A_PLANE_BYTES) + row_lo * Tr::APITCH * 16;
          // y = silu(a x + d): packed fp32 FMAs (one issue slot per two channels), one MUFU.TANH per channel
          auto xform_unit = [&](uint4 raw) {
            f32x2 v2[CW / 2];
            Elem<T>::unpack2x(raw, v2);
#pragma unroll
            for (int k = 0; k < CW / 2; ++k) {
              const f32x2 t2 = fma2(v2[k], ca2[k], cd2[k]);
              if (p.xf.silu) {
                float lo, hi;
                unpack2(t2, lo, hi);
                if (sizeof(T) == 2 || R2DM_F32_TANH) {
                  float tl, th;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(tl) : "f"(lo));
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hi));
                  v2[k] = fma2(t2, pack2(tl, th), t2);      // h + h tanh(h), h = t / 2 (folded into a, d)
                } else {
                  v2[k] = pack2(silu_fast(lo), silu_fast(hi));
                }
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
          if (!R2DM_DBG(p.xf.debug != 0)) {
            // full batches of XB units per thread: all loads first, then the math, then the stores;
            // the remainder (n_units is not a multiple of XB * TPP) goes one unit at a time so that no
            // XU-pipe slots are spent on padding
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
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&xf_bar[st]);
          if (tt == 0) R2DM_TRACE(grp ? 4 : 2, 2 * it + 1);
        }
        if constexpr (TAPS == 9) {
          // skip stages carry raw data: nothing to transform, only forward "landed" to the MMA warp
          for (int ks = 0; ks < p.nk2; ++ks, ++it, st = (st + 1 == p.stages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
            if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
            mbar_wait_relaxed(&full_bar[st], ph, 500);
            __syncwarp();
            if (lane == 0) mbar_arrive(&xf_bar[st]);
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 8) {
    // ------------------------------------------------------------------ epilogue
    // Compact on purpose: the column-chunk loop is NOT unrolled (the fully unrolled version was
    // 96 KB of SASS and the eight epilogue warps stalled on instruction fetch).
    const int ew = warp - kEpiWarp0;    // 0..7
    const int q = ew & 3;               // TMEM lane quarter (must equal warp % 4)
    const int half = ew >> 2;           // which rows (HT > 1) or which column half (HT == 1)
    const int m = q * 32 + lane;
    const int Wp = p.W + 2;
    const int planes_out = p.cout_pad / CW;
    // 16-byte units between channel planes; unit indices fit in 32 bits (a tensor of 2^32 units would be 64 GB)
    const uint32_t plane_stride = static_cast<uint32_t>(p.H) * static_cast<uint32_t>(Wp);
    constexpr int CB = NT >= 32 ? 32 : 16;              // columns per chunk
    constexpr int CCOLS = HT > 1 ? NT : NT / 2;         // columns visited by this warp
    constexpr int NCHUNK = CCOLS / CB;
    constexpr int NSUB = CB / 8;                        // 8-channel statistics sub-chunks per chunk
    constexpr int RSTEP = HT > 1 ? 2 : 1;
    const int r_begin = HT > 1 ? half : 0;
    const int c_begin = HT > 1 ? 0 : half * (NT / 2);
    const uint4* res = static_cast<const uint4*>(p.residual);
    uint4* out = static_cast<uint4*>(p.out);
    const int ethread = threadIdx.x - kEpiWarp0 * 32;
    pdl_wait();
    int j = 0, cur_nt = -1;
    int pw_b = -1;
    auto pw_flush = [&](int bb) {     // all 256 epilogue threads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = ethread; i < p.cout_pad; i += 256) {
        const float v = pw_max[i];
        if (bb >= 0 && v > -INFINITY) {
          float* dst = p.colmax + static_cast<size_t>(bb) * p.cout_pad + i;
          // float maximum with integer atomics: non-negative values order like ints, negative ones like
          // reversed unsigned ints; the buffer starts at -inf
          if (v >= 0.f) atomicMax(reinterpret_cast<int*>(dst), __float_as_int(v));
          else atomicMin(reinterpret_cast<unsigned int*>(dst), __float_as_uint(v));
        }
        pw_max[i] = -INFINITY;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int t = t_begin; t < t_end; ++t, ++j) {
      int b, yt, xt, nt;
      decode(t, b, yt, xt, nt);
      if constexpr (PW) {
        if (p.colmax != nullptr && b != pw_b) { pw_flush(pw_b); pw_b = b; }
      }
      const int n0 = nt * NT, x = xt * 128 + m, y0 = yt * HT;
      const int buf = j & 1, par = j & 1;
      // the thread that holds pixel 0 (W-1) also writes the wrap halo column xp = W+1 (xp = 0): one predicated
      // store at a signed offset, no divergent block inside the unit loop
      const bool seam = (x == 0) || (x == p.W - 1);
      const int seam_off = (x == 0) ? p.W : -p.W;
      if (nt != cur_nt) {   // bias of this N tile -> shared memory (once per CTA when ntiles == 1)
        cur_nt = nt;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = ethread; i < NT; i += 256)   // pre-scaled: one FMA per value; bias2 = the folded skip projection's
          bias_s[i] = (p.bias[n0 + i] + (p.bias2 ? p.bias2[n0 + i] : 0.f)) * p.scale;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (!NCHW && res != nullptr) {
        // pull the residual tile into L2 while the MMAs of this tile run (one 128-byte line per request)
        constexpr int SEGS = HT * (NT / CW);             // (row, plane) segments of 128 px = 2 KB
        for (int l = ethread; l < SEGS * 16; l += 256) {
          const int seg = l >> 4, row = seg / (NT / CW), pl = seg % (NT / CW);
          if (y0 + row < p.H) {
            const uint4* a = res + pt_index(b, planes_out, n0 / CW + pl, p.H, Wp, y0 + row, xt * 128 + 1) + (l & 15) * 8;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
      // unit index of (image b, plane of channel n0, row y0, this thread's pixel); rows and planes are offsets from it
      const uint32_t tile_idx = static_cast<uint32_t>(pt_index(b, planes_out, n0 / CW, p.H, Wp, y0, x + 1));
      mbar_wait_relaxed(&acc_full[buf], (j >> 1) & 1, 1000);
      tc_fence_after();
      if (ethread == 0) R2DM_TRACE(3, 3 * j);
      const uint32_t tbase = tmem + buf * Tr::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < (R2DM_DBG(p.debug & 1) ? 0 : NCHUNK); ++ch) {
        const int c0 = c_begin + ch * CB;
        // partial (sum, sum of squares) of this warp's pixels per 8-channel sub-chunk, as fp32 pairs over
        // (even, odd) channels: packed FADD2 / FFMA2, one issue slot per two values
        f32x2 s1p[NSUB], s2p[NSUB];
#pragma unroll
        for (int u = 0; u < NSUB; ++u) { s1p[u] = pack2(0.f, 0.f); s2p[u] = pack2(0.f, 0.f); }
        const f32x2 scale2 = pack2(p.scale, p.scale);
        // 1x1 launches: per-channel maximum over this thread's pixels (point-wise MLP + global max pool)
        float cm[PW ? CB : 1];
        if constexpr (PW) {
#pragma unroll
          for (int i = 0; i < CB; ++i) cm[i] = -INFINITY;
        }
#pragma unroll
        for (int r = r_begin; r < HT; r += RSTEP) {
          const int y = y0 + r;
          if (y < p.H) {
            const uint32_t idx0 = tile_idx + static_cast<uint32_t>(r) * static_cast<uint32_t>(Wp) +
                                  static_cast<uint32_t>(c0 / CW) * plane_stride;
            // residual prefetch (independent loads in flight while TMEM is read)
            uint4 rr[CB / CW];
            if (!NCHW && res != nullptr) {
#pragma unroll
              for (int u = 0; u < CB / CW; ++u) rr[u] = R2DM_DBG(p.debug & 8) ? make_uint4(0u, 0u, 0u, 0u) : res[idx0 + u * plane_stride];
            }
            float v[CB];
#pragma unroll
            for (int h16 = 0; h16 < CB / 16; ++h16) tmem_ld16(tbase + r * NT + c0 + h16 * 16, v + h16 * 16);
            tmem_ld_wait();
            const float4* bias4 = reinterpret_cast<const float4*>(bias_s + c0);
            f32x2 v2[CB / 2];
#pragma unroll
            for (int i4 = 0; i4 < CB / 4; ++i4) {
              const float4 bv = bias4[i4];      // (acc + bias) * scale = acc * scale + bias * scale
              v2[2 * i4] = fma2(pack2(v[4 * i4], v[4 * i4 + 1]), scale2, pack2(bv.x, bv.y));
              v2[2 * i4 + 1] = fma2(pack2(v[4 * i4 + 2], v[4 * i4 + 3]), scale2, pack2(bv.z, bv.w));
            }
            if constexpr (NCHW) {
              float* dst = p.out_nchw + ((static_cast<size_t>(b) * p.cout + n0 + c0) * p.H + y) * p.W + x;
              const size_t cstride = static_cast<size_t>(p.H) * p.W;
#pragma unroll
              for (int i = 0; i < CB / 2; ++i) {
                float lo, hi;
                unpack2(v2[i], lo, hi);
                if (n0 + c0 + 2 * i < p.cout) dst[(2 * i) * cstride] = lo;
                if (n0 + c0 + 2 * i + 1 < p.cout) dst[(2 * i + 1) * cstride] = hi;
              }
            } else {
              constexpr int PPU = CW / 2;                   // pairs per 16-byte unit
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
                if constexpr (PW) {
                  {
#pragma unroll
                    for (int i = 0; i < PPU; ++i) {
                      float lo, hi;
                      unpack2(o2[i], lo, hi);
                      if (p.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); o2[i] = pack2(lo, hi); }
                      cm[u * CW + 2 * i] = fmaxf(cm[u * CW + 2 * i], lo);
                      cm[u * CW + 2 * i + 1] = fmaxf(cm[u * CW + 2 * i + 1], hi);
                    }
                  }
                }
                const int sub = (u * CW) / 8;
#pragma unroll
                for (int i = 0; i < PPU; ++i) {
                  s1p[sub] = add2(s1p[sub], o2[i]);
                  s2p[sub] = fma2(o2[i], o2[i], s2p[sub]);
                }
                if constexpr (PW) {
                  if (out == nullptr) continue;      // max-pool only: the tensor itself is not needed
                }
                uint4 pk = Elem<T>::pack2x(o2);
                if (sizeof(T) == 4 && p.round_out) {
                  asm("cvt.rna.tf32.f32 %0, %0;" : "+r"(pk.x));
                  asm("cvt.rna.tf32.f32 %0, %0;" : "+r"(pk.y));
                  asm("cvt.rna.tf32.f32 %0, %0;" : "+r"(pk.z));
                  asm("cvt.rna.tf32.f32 %0, %0;" : "+r"(pk.w));
                }
                out[idx] = pk;
                if (seam) out[static_cast<int>(idx) + seam_off] = pk;
              }
            }
          }
        }
        if constexpr (PW && !NCHW) {
          if (p.colmax != nullptr) {
            // transposing butterfly: afterwards lane i holds the warp maximum of channel c0 + i (CB = 32) -
            // 31 shuffles instead of 5 per channel
            static_assert(CB == 32, "column maximum assumes 32-channel chunks");
#pragma unroll
            for (int rd = 0; rd < 5; ++rd) {
              const int nv = 32 >> rd, off = 16 >> rd;
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < nv / 2; ++i) {
                const float send = upper ? cm[i] : cm[i + nv / 2];
                const float keep = upper ? cm[i + nv / 2] : cm[i];
                cm[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
              }
            }
            float* dst = pw_max + n0 + c0 + lane;      // shared-memory atomics: eight warps share a channel
            const float v = cm[0];
            if (v >= 0.f) atomicMax(reinterpret_cast<int*>(dst), __float_as_int(v));
            else atomicMin(reinterpret_cast<unsigned int*>(dst), __float_as_uint(v));
          }
        }
        float ssum[NSUB][2];
#pragma unroll
        for (int u = 0; u < NSUB; ++u) {
          float lo, hi;
          unpack2(s1p[u], lo, hi); ssum[u][0] = lo + hi;
          unpack2(s2p[u], lo, hi); ssum[u][1] = lo + hi;
        }
        if (!NCHW && p.stats != nullptr) {
          // warp totals of the 2*NSUB partial sums with a transposing butterfly (NSUB*2 - 1 + 2
          // shuffles instead of 5 per value): afterwards value i lives in the lanes whose bits
          // [4:..] spell i; lane with (lane & (32/(2*NSUB) - 1)) == 0 stores it
          float vals[2 * NSUB];
#pragma unroll
          for (int u = 0; u < NSUB; ++u) { vals[2 * u] = ssum[u][0]; vals[2 * u + 1] = ssum[u][1]; }
          constexpr int LOGV = NSUB == 4 ? 3 : (NSUB == 2 ? 2 : 1);   // log2(2 * NSUB)
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
          constexpr int LPV = 32 / (2 * NSUB);          // lanes per value
          if ((lane & (LPV - 1)) == 0) {
            const int vi = lane / LPV;                  // = 2 * sub + k
            stat_w[par][ew][(c0 >> 3) + (vi >> 1)][vi & 1] = vals[0];
          }
        }
      }
      // this warp is done with the TMEM buffer: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (ethread == 0) R2DM_TRACE(3, 3 * j + 1);
      if (!NCHW && p.stats != nullptr) {
        // fold 8-channel sub-chunks into statistics units (unit_ch = 8 << k channels) and across warps
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int units_here = NT / p.unit_ch;
        if (ethread < units_here * 2) {
          const int u = ethread >> 1, k = ethread & 1;
          const int cpu = p.unit_ch >> 3;                // sub-chunks per unit (power of two)
          float tot = 0.f;
          for (int sc = u * cpu; sc < (u + 1) * cpu; ++sc) {
            // HT > 1: every warp visited all columns; HT == 1: warps 0-3 the lower half, 4-7 the upper
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              const bool has = HT > 1 ? true : (w >> 2) == (sc >= NT / 16 ? 1 : 0);
              if (has) tot += stat_w[par][w][sc][k];
            }
          }
          const int unit = n0 / p.unit_ch + u;
          const int slot = yt * p.xtiles + xt;
          p.stats[((static_cast<size_t>(b) * kNU + unit) * p.slots + slot) * 2 + k] = tot;
        }
      }
    }
    if constexpr (PW) {
      if (p.colmax != nullptr) pw_flush(pw_b);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (R2DM_DBG(p.trace != nullptr && blockIdx.x == p.trace_block && threadIdx.x == 0 && p.trace_cap >= 8)) {
    p.trace[p.trace_cap - 2] = static_cast<unsigned long long>(clock64());   // ... and at the end: SM clock rate
    p.trace[p.trace_cap - 1] = gtime();
  }
  if (warp == kAllocWarp) tmem_dealloc<Tr::TMEM_COLS>(tmem);
  if (p.ktime != nullptr && threadIdx.x == 0) atomicMax(p.ktime + 1, gtime());
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int ks_for(int taps) { return taps == 9 ? 1 : 4; }
int conv_stage_channels(int dtype, int taps) { return ks_for(taps) * 2 * dtype_cw(dtype); }
int conv_skip_planes(int nt) { return nt >= 128 ? 8 : 2; }   // = ConvTraits::SK_PLANES

size_t conv_packed_weight_bytes(int dtype, int taps, int nt, int cin_pad, int cout_pad) {
  (void)nt;
  return static_cast<size_t>(taps) * cin_pad * cout_pad * dtype_size(dtype);
}

int conv_stat_slots(const ConvLaunch& l) { return (l.out.H / l.ht) * (l.out.W / 128); }

// The planar-16 tensor [B][planes][H][Wp][16 B] is described to the TMA unit with 8-byte elements
// and the pixel axis split in two halves (dim1, stride = half a box row) so that one inner box
// segment is box_w*8 bytes (1040 B) instead of 16 B: a 16-byte inner dimension makes the TMA
// unit issue one request per pixel-plane and was the bottleneck of the first version.
//   dims (fastest first): [2*Wp elements of 8 B][2 halves][H][planes][B]
//   box                 : [box_w elements][2][box_h][box_planes][1]   -> box_w pixels x 16 B per row
static int make_one_tmap(CUtensorMap* tm, int dtype, const PT& t, int box_w, int box_h, int box_planes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const int cw = dtype_cw(dtype);
  const cuuint64_t Wp = t.W + 2;
  cuuint64_t dims[5] = {2 * Wp, 2, (cuuint64_t)t.H, (cuuint64_t)(t.C / cw), (cuuint64_t)t.B};
  cuuint64_t strides[4] = {(cuuint64_t)box_w * 8, Wp * 16, (cuuint64_t)t.H * Wp * 16,
                           (cuuint64_t)(t.C / cw) * t.H * Wp * 16};
  cuuint32_t box[5] = {(cuuint32_t)box_w, 2, (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, t.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

// qkv tensor maps for the attention kernel: boxes of 128 (Q) / key-tile (K, V) pixels of one row x (hd/CW) planes
int attention_make_tmaps(CUtensorMap* tm_q, CUtensorMap* tm_kv, int dtype, const PT& qkv, int heads) {
  const int hd = qkv.C / 3 / heads;
  int rc = make_one_tmap(tm_q, dtype, qkv, 128, 1, hd / dtype_cw(dtype));
  if (rc) return rc;
  return make_one_tmap(tm_kv, dtype, qkv, attention_key_tile(dtype), 1, hd / dtype_cw(dtype));
}

int conv_make_tmaps(ConvLaunch& l) {
  const int planes = 2 * ks_for(l.taps);
  const int bw = l.taps == 9 ? 130 : 128, bh = l.taps == 9 ? l.ht + 2 : l.ht;
  int rc = make_one_tmap(&l.tmap0, l.dtype, l.in0, bw, bh, planes);
  if (rc) return rc;
  if (l.in1.ptr) rc = make_one_tmap(&l.tmap1, l.dtype, l.in1, bw, bh, planes);
  else l.tmap1 = l.tmap0;
  if (rc) return rc;
  l.tmap2 = l.tmap3 = l.tmap0;
  if (l.sk0.ptr) {
    if (l.taps != 9) return -2;
    rc = make_one_tmap(&l.tmap2, l.dtype, l.sk0, 128, l.ht, conv_skip_planes(l.nt));
    if (rc) return rc;
    if (l.sk1.ptr) rc = make_one_tmap(&l.tmap3, l.dtype, l.sk1, 128, l.ht, conv_skip_planes(l.nt));
    else l.tmap3 = l.tmap2;
  }
  return rc;
}

int conv_num_sms() {
  static int n[64] = {0};   // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (v == 0 && (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)) v = 148;
  return v;
}

bool pdl_enabled() {
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("R2DM_PDL"); pdl = e ? atoi(e) : 1; }
  return pdl != 0;
}

static unsigned long long* g_trace = nullptr;
static int g_trace_cap = 0;
static int g_trace_skip = 0;      // developer: trace only the (skip+1)-th conv launch after conv_set_trace
void conv_set_trace(unsigned long long* buf, int cap) {
  g_trace = buf; g_trace_cap = cap;
  const char* e = getenv("R2DM_TRACE_SKIP");
  g_trace_skip = (buf && e) ? atoi(e) : 0;
}
// the trace buffer for this launch (null unless it is the selected one)
unsigned long long* conv_trace_for_this_launch() {
  if (g_trace == nullptr) return nullptr;
  if (getenv("R2DM_TRACE_SKIP") == nullptr) return g_trace;     // legacy: every launch writes (last one wins)
  return g_trace_skip-- == 0 ? g_trace : nullptr;
}
int conv_trace_cap() { return g_trace_cap; }

constexpr int kSmemBudget = 224 * 1024;     // dynamic smem per CTA (227 KB limit minus < 3 KB static)
constexpr int kWresMaxBytes = 80 * 1024;    // keep the filter bank resident below this size

template <typename T, int NT, int HT, int TAPS, int KS, bool NCHW = false, bool PW = false>
static cudaError_t launch_one(const ConvLaunch& l, cudaStream_t s) {
  using Tr = ConvTraits<T, NT, HT, TAPS, KS>;
  auto kern = conv_umma_kernel<T, NT, HT, TAPS, KS, NCHW, PW>;
  constexpr int kBudget = kSmemBudget - (PW ? kMaxPwChannels * static_cast<int>(sizeof(float)) : 0);   // PW: static pw_max
  if ((l.out_nchw != nullptr) != NCHW) return cudaErrorInvalidConfiguration;
  static unsigned long long configured = 0;   // one bit per device (the attribute is per device)
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBudget);
    if (e != cudaSuccess) return e;
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tmap0 = l.tmap0; p.tmap1 = l.tmap1; p.tmap2 = l.tmap2; p.tmap3 = l.tmap3;
  p.wpacked = l.wpacked; p.bias = l.bias; p.residual = l.residual;
  if (l.sk0.ptr != nullptr) {
    if (TAPS != 9 || l.w2packed == nullptr || l.cin2_pad % Tr::SK_KCH != 0 || l.sk0.C % Tr::SK_KCH != 0)
      return cudaErrorInvalidValue;
    p.w2packed = l.w2packed; p.bias2 = l.bias2;
    p.nk2 = l.cin2_pad / Tr::SK_KCH;
    p.ksplit2 = l.sk1.ptr ? l.sk0.C / Tr::SK_KCH : p.nk2;
  }
  p.out = l.out.ptr; p.out_nchw = l.out_nchw; p.stats = l.out_nchw ? nullptr : l.out.stats;
  p.B = l.out.B; p.H = l.out.H; p.W = l.out.W;
  p.cout = l.cout; p.cout_pad = l.cout_pad;
  p.nk = l.cin_pad / Tr::KCH;
  p.ksplit = l.in1.ptr ? l.in0.C / Tr::KCH : p.nk;
  p.xtiles = l.out.W / 128; p.ytiles = (l.out.H + HT - 1) / HT; p.ntiles = l.cout_pad / NT;
  p.tiles_total = p.B * p.ytiles * p.xtiles * p.ntiles;
  p.unit_ch = l.cout / kNU >= 8 ? l.cout / kNU : 8;
  p.unit_shift = 0;
  while ((1 << p.unit_shift) < p.unit_ch) ++p.unit_shift;
  if (p.stats != nullptr && ((1 << p.unit_shift) != p.unit_ch || p.unit_ch > NT)) return cudaErrorInvalidValue;
  p.slots = l.out.slots;
  p.scale = l.scale;
  // operand transform (fused GroupNorm / AdaGN + SiLU)
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
    if (x.C0 + x.C1 != l.cin_pad) return cudaErrorInvalidValue;   // no channel padding under the transform
  }
  // shared-memory plan: resident weights when the whole bank of this N tile fits, then as many
  // ring stages as the budget allows
  const size_t wbytes = static_cast<size_t>(p.nk) * Tr::B_BYTES;
  p.wres = (p.ntiles == 1 && wbytes <= static_cast<size_t>(kWresMaxBytes)) ? 1 : 0;
  if (p.nk2 > 0 && p.wres && Tr::SK_A_BYTES + Tr::SK_B_BYTES > Tr::A_BYTES_AL) p.wres = 0;   // a skip stage must fit in a slot
  p.stage_bytes = Tr::A_BYTES_AL + (p.wres ? 0 : Tr::B_BYTES);
  if (p.nk2 > 0 && Tr::SK_A_BYTES + Tr::SK_B_BYTES > p.stage_bytes) return cudaErrorInvalidConfiguration;
  p.coef_ch = l.xf.enabled ? (p.xf.C0 + p.xf.C1 + 31) / 32 * 32 : 0;
  p.coef_bytes = 2 * p.coef_ch * static_cast<int>(sizeof(float));
  p.reverse = l.reverse;
  p.round_out = l.round_out;
  p.relu = l.relu; p.colmax = l.colmax;
  p.l2_hint = get_option("l2_evict_first", 0);
  p.prefetch_w = get_option("prefetch_w", 0);   // experiment: +-0 (2.375 vs 2.376 ms per forward), off by default
  if ((l.relu || l.colmax != nullptr) != PW) return cudaErrorInvalidConfiguration;
  if (PW && l.cout_pad > kMaxPwChannels) return cudaErrorInvalidValue;
  p.ktime = l.ktime;
  const int avail = kBudget - 256 - p.coef_bytes - (p.wres ? static_cast<int>(wbytes) : 0);
  int stages = avail / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  { const int cap = get_option("max_stages", kMaxStages); if (stages > cap) stages = cap; }
  // The two transform groups take alternating stages.  With an ODD ring depth the owner of a slot would alternate
  // between its uses, and a group that is ahead (e.g. after pass-through skip stages, whose small loads can land
  // before the preceding larger one) would wait for phase k+1 of a slot whose phase k has not completed yet - an
  // mbarrier parity wait cannot tell these apart and passes immediately.  An even depth keeps one owner per slot.
  if (l.xf.enabled && stages > 2) stages &= ~1;
  if (stages < 2) return cudaErrorInvalidConfiguration;
  p.stages = stages;
  const int smem = 256 + p.coef_bytes + (p.wres ? static_cast<int>(wbytes) : 0) + stages * p.stage_bytes;
  p.trace = conv_trace_for_this_launch(); p.trace_cap = g_trace_cap;
  { const char* e = getenv("R2DM_TRACE_BLOCK"); p.trace_block = e ? static_cast<unsigned>(atoi(e)) : 0u; }
  {
    static int cdbg = -1;
    if (cdbg < 0) { const char* e = getenv("R2DM_CONV_DEBUG"); cdbg = e ? atoi(e) : 0; }
    p.debug = cdbg;
  }
  int grid = conv_num_sms();
  {
    static int cap = -1;   // developer knob: cap the persistent grid (R2DM_CONV_GRID), e.g. to share the GPU between streams
    if (cap < 0) { const char* e = getenv("R2DM_CONV_GRID"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && grid > cap) grid = cap;
  }
  if (grid > p.tiles_total) grid = p.tiles_total;
  if (get_option("compact_grid", 0)) {
    // experiment (measured: no gain, 2.43 vs 2.42 ms per forward): the launch ends with the CTAs that own
    // ceil(tiles / grid) tiles, so ceil(tiles / that) CTAs give the same critical path (512 tiles: 128 x 4 instead
    // of 68 x 4 + 80 x 3) and the idle SMs would return their share of the power budget
    const int per = (p.tiles_total + grid - 1) / grid;
    grid = (p.tiles_total + per - 1) / per;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}


template <typename T>
static cudaError_t dispatch(const ConvLaunch& l, cudaStream_t s) {
  if (l.taps == 9) {
    if (l.nt == 64 && l.ht == 4) return launch_one<T, 64, 4, 9, 1>(l, s);
    if (l.nt == 64 && l.ht == 2) return launch_one<T, 64, 2, 9, 1>(l, s);
    if (l.nt == 128 && l.ht == 2) return launch_one<T, 128, 2, 9, 1>(l, s);
    if (l.nt == 128 && l.ht == 1) return launch_one<T, 128, 1, 9, 1>(l, s);
    if (l.nt == 16 && l.ht == 4)
      return l.out_nchw ? launch_one<T, 16, 4, 9, 1, true>(l, s) : launch_one<T, 16, 4, 9, 1>(l, s);
  } else if (l.taps == 1 && (l.relu || l.colmax != nullptr)) {
    if (l.nt == 64 && l.ht == 2) return launch_one<T, 64, 2, 1, 4, false, true>(l, s);
    if (l.nt == 128 && l.ht == 2) return launch_one<T, 128, 2, 1, 4, false, true>(l, s);
    if (l.nt == 128 && l.ht == 1) return launch_one<T, 128, 1, 1, 4, false, true>(l, s);
    if (l.nt == 64 && l.ht == 1) return launch_one<T, 64, 1, 1, 4, false, true>(l, s);
  } else if (l.taps == 1) {
    if (l.nt == 64 && l.ht == 2) return launch_one<T, 64, 2, 1, 4>(l, s);
    if (l.nt == 128 && l.ht == 2) return launch_one<T, 128, 2, 1, 4>(l, s);
    if (l.nt == 128 && l.ht == 1) return launch_one<T, 128, 1, 1, 4>(l, s);
    if (l.nt == 64 && l.ht == 1) return launch_one<T, 64, 1, 1, 4>(l, s);
  }
  return cudaErrorInvalidConfiguration;
}

cudaError_t conv_launch(const ConvLaunch& l, cudaStream_t s) {
  if (l.out.W % 128 != 0 || l.cout_pad % l.nt != 0 || l.cin_pad % conv_stage_channels(l.dtype, l.taps) != 0)
    return cudaErrorInvalidValue;
  return l.dtype == kBF16 ? dispatch<__nv_bfloat16>(l, s) : dispatch<float>(l, s);
}

// ------------------------------------------------------------------------------------ weights
// dst[nt][ks][tap][plane][co][cw]  <-  w[co][ci][tap]   (zero padded); with fuse_dy (3x3, nt = 64)
// the order inside a stage is [kx][plane][ky descending][co][cw] (see ConvTraits::FUSE)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int taps, int nt,
                                   int cout, int cin, int cin_pad, int cout_pad, int planes, int fuse_dy) {
  constexpr int CW = Elem<T>::CW;
  const size_t total = static_cast<size_t>(taps) * cin_pad * cout_pad;
  const int kch = planes * CW;
  const int nk = cin_pad / kch;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int cw = r % CW; r /= CW;
    const int co = r % nt; r /= nt;
    int pl, tap;
    if (fuse_dy) {
      const int kyd = r % 3; r /= 3;
      pl = r % planes; r /= planes;
      const int kx = r % 3; r /= 3;
      tap = (2 - kyd) * 3 + kx;
    } else {
      pl = r % planes; r /= planes;
      tap = r % taps; r /= taps;
    }
    const int ks = r % nk; r /= nk;
    const int nti = static_cast<int>(r);
    const int ci = ks * kch + pl * CW + cw;
    const int o = nti * nt + co;
    float v = 0.f;
    if (ci < cin && o < cout) v = w[(static_cast<size_t>(o) * cin + ci) * taps + tap];
    if (sizeof(T) == 4) {  // tf32 operand: round to nearest instead of the tensor core's truncation
      uint32_t rr;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v));
      v = __uint_as_float(rr);
    }
    dst[i] = static_cast<T>(v);
  }
}

cudaError_t pack_conv_weight(int dtype, int taps, int nt, const float* w, int cout, int cin, int cin_pad,
                             int cout_pad, void* dst, cudaStream_t s, int planes_arg) {
  const int planes = planes_arg > 0 ? planes_arg : 2 * ks_for(taps);
  const int fuse = (taps == 9 && (nt == 64 || nt == 128 || nt == 16)) ? 1 : 0;
  const size_t total = static_cast<size_t>(taps) * cin_pad * cout_pad;
  const int grid = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  if (dtype == kBF16)
    pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(dst), taps, nt, cout,
                                                           cin, cin_pad, cout_pad, planes, fuse);
  else
    pack_weight_kernel<float><<<grid, 256, 0, s>>>(w, static_cast<float*>(dst), taps, nt, cout, cin, cin_pad,
                                                   cout_pad, planes, fuse);
  return cudaGetLastError();
}

}  // namespace r2dm
