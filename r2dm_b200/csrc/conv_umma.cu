// Ring 3x3 / 1x1 convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces, on the sampling path, models/ops.py:149-173 (Conv2d + Pad: circular azimuth padding,
// zero elevation padding, 3x3 cross-correlation + bias), the 1x1 skip convolution of
// models/efficient_unet.py:87-91 and the nn.MultiheadAttention in/out projections (:34-38).
//
// One CTA computes an output tile of HT rows x 128 pixels x NT output channels:
//   * warp 0   : producer.  Per pipeline stage (KCH input channels) one 5-D TMA box load brings the
//                halo tile [(HT+2) x 130 px] of those channels into shared memory (planar-16
//                layout, see common.cuh) and one bulk copy brings the pre-packed weights of all
//                taps for those channels.
//   * warp 1   : MMA issuer.  For every output row and every tap it issues tcgen05.mma with the A
//                descriptor pointing at the *shifted* start address inside the halo tile, so the
//                input is read from L2 once per tile instead of once per tap.  Accumulators for
//                the HT rows live in TMEM (HT x NT fp32 columns).
//   * warp 2   : TMEM allocator.
//   * warps 4-7: epilogue.  tcgen05.ld -> + bias (+ residual) -> * scale -> GroupNorm partial
//                sums for the consumer -> bf16/fp32 planar-16 store (incl. the wrap halo columns),
//                or fp32 NCHW store for the network output.
// Elevation borders come from TMA out-of-bounds zero fill; the azimuth wrap from the halo columns.
#include <cstdio>
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace r2dm {

struct ConvParams {
  CUtensorMap tmap0, tmap1;
  const void* wpacked;
  const float* bias;
  const void* residual;
  void* out;
  float* out_nchw;
  float* stats;
  int B, H, W;
  int cout, cout_pad;     // real / padded output channels
  int nk, ksplit;         // pipeline stages over K; first stage that reads from tmap1
  int xtiles, ytiles, ntiles;
  int unit_ch;            // output channels per statistics unit (cout / kNU)
  int slots;
  float scale;
};

template <typename T, int NT, int HT, int TAPS, int KS, int STAGES>
struct ConvTraits {
  static constexpr int CW = Elem<T>::CW;
  static constexpr int KCH = KS * 2 * CW;
  static constexpr int PLANES = 2 * KS;
  static constexpr int AROWS = TAPS == 9 ? HT + 2 : HT;
  static constexpr int APITCH = TAPS == 9 ? 130 : 128;
  static constexpr int A_PLANE_BYTES = AROWS * APITCH * 16;
  static constexpr int A_BYTES = PLANES * A_PLANE_BYTES;
  static constexpr int A_BYTES_AL = (A_BYTES + 127) / 128 * 128;
  static constexpr int B_PLANE_BYTES = NT * 16;
  static constexpr int B_TAP_BYTES = PLANES * B_PLANE_BYTES;
  static constexpr int B_BYTES = TAPS * B_TAP_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES_AL + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128;
  static constexpr int ACC_COLS = HT * NT;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(B_BYTES % 128 == 0, "weight stage must stay 128B aligned");
};

template <typename T, int NT, int HT, int TAPS, int KS, int STAGES, int MINB>
__global__ void __launch_bounds__(256, MINB)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
  using Tr = ConvTraits<T, NT, HT, TAPS, KS, STAGES>;
  constexpr int CW = Tr::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[NT];
  __shared__ float stat_s[4][kNU][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bid = blockIdx.x;
  const int nt = bid % p.ntiles; bid /= p.ntiles;
  const int xt = bid % p.xtiles; bid /= p.xtiles;
  const int yt = bid % p.ytiles;
  const int b = bid / p.ytiles;
  const int n0 = nt * NT, x0 = xt * 128, y0 = yt * HT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmap0);
    if (p.ksplit < p.nk) tma_prefetch_desc(&p.tmap1);
  }
  if (warp == 2) tmem_alloc<Tr::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(nt) * p.nk * Tr::B_BYTES;
      for (int ks = 0; ks < p.nk; ++ks) {
        const int st = ks % STAGES;
        const uint32_t ph = (ks / STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t* sa = smem + st * Tr::STAGE_BYTES;
        uint8_t* sb = sa + Tr::A_BYTES_AL;
        mbar_expect_tx(&full_bar[st], Tr::A_BYTES + Tr::B_BYTES);
        const bool second = ks >= p.ksplit;
        const int plane0 = (second ? ks - p.ksplit : ks) * Tr::PLANES;
        const CUtensorMap* tm = second ? &p.tmap1 : &p.tmap0;
        if (TAPS == 9) tma_load_5d(sa, tm, &full_bar[st], 0, x0, y0 - 1, plane0, b);
        else tma_load_5d(sa, tm, &full_bar[st], 0, x0 + 1, y0, plane0, b);
        bulk_load(sb, wsrc + static_cast<size_t>(ks) * Tr::B_BYTES, Tr::B_BYTES, &full_bar[st]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, NT, Elem<T>::kFmt);
      // descriptor halves: lo = start>>4 | LBO>>4 << 16 ; hi = SBO>>4 | version<<14 | layout<<29
      const uint32_t a_lo_const = static_cast<uint32_t>(Tr::A_PLANE_BYTES >> 4) << 16;
      const uint32_t b_lo_const = static_cast<uint32_t>(Tr::B_PLANE_BYTES >> 4) << 16;
      const uint32_t hi = (128u >> 4) | (1u << 14);
      for (int ks = 0; ks < p.nk; ++ks) {
        const int st = ks % STAGES;
        const uint32_t ph = (ks / STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * Tr::STAGE_BYTES);
        const uint32_t sb = sa + Tr::A_BYTES_AL;
#pragma unroll
        for (int r = 0; r < HT; ++r) {
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            const int dy = TAPS == 9 ? tap / 3 : 0, dx = TAPS == 9 ? tap % 3 : 0;
            const uint32_t a_off = static_cast<uint32_t>(((r + dy) * Tr::APITCH + dx) * 16);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              const uint32_t aa = sa + kk * 2 * Tr::A_PLANE_BYTES + a_off;
              const uint32_t ba = sb + tap * Tr::B_TAP_BYTES + kk * 2 * Tr::B_PLANE_BYTES;
              const uint64_t adesc = (static_cast<uint64_t>(hi) << 32) | (a_lo_const | ((aa >> 4) & 0x3FFFu));
              const uint64_t bdesc = (static_cast<uint64_t>(hi) << 32) | (b_lo_const | ((ba >> 4) & 0x3FFFu));
              const uint32_t acc = (ks > 0 || tap > 0 || kk > 0) ? 1u : 0u;
              if (Elem<T>::kFmt == 2) umma_tf32(tmem + r * NT, adesc, bdesc, idesc, acc);
              else umma_f16(tmem + r * NT, adesc, bdesc, idesc, acc);
            }
          }
        }
        umma_commit(&empty_bar[st]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(&accum_bar);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int m = ew * 32 + lane;
    const int x = x0 + m;
    for (int i = m; i < NT; i += 128) bias_s[i] = p.bias[n0 + i];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int Wp = p.W + 2;
    const int planes_out = p.cout_pad / CW;
    constexpr int NUT = NT >= 8 * kNU ? kNU : (NT / 8 > 0 ? NT / 8 : 1);  // max units per tile
    float sacc[NUT][2];
#pragma unroll
    for (int u = 0; u < NUT; ++u) { sacc[u][0] = 0.f; sacc[u][1] = 0.f; }
    const uint4* res = static_cast<const uint4*>(p.residual);
    uint4* out = static_cast<uint4*>(p.out);

    mbar_wait(&accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int r = 0; r < HT; ++r) {
      const int y = y0 + r;
      if (y >= p.H) break;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(ew * 32) << 16) + r * NT + c0, v);
        tmem_ld_wait();
        if (p.out_nchw != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ch = n0 + c0 + i;
            if (ch < p.cout)
              p.out_nchw[((static_cast<size_t>(b) * p.cout + ch) * p.H + y) * p.W + x] =
                  (v[i] + bias_s[c0 + i]) * p.scale;
          }
          continue;
        }
#pragma unroll
        for (int u = 0; u < 16 / CW; ++u) {
          const int cl = c0 + u * CW;  // channel within the N tile
          const int plane = (n0 + cl) / CW;
          const size_t idx = pt_index(b, planes_out, plane, p.H, Wp, y, x + 1);
          float o[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) o[i] = v[u * CW + i] + bias_s[cl + i];
          if (res != nullptr) {
            float rv[CW];
            Elem<T>::unpack(res[idx], rv);
#pragma unroll
            for (int i = 0; i < CW; ++i) o[i] += rv[i];
          }
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < CW; ++i) { o[i] *= p.scale; s1 += o[i]; s2 += o[i] * o[i]; }
          if (p.stats != nullptr) {
            const int un = cl / p.unit_ch;
#pragma unroll
            for (int q = 0; q < NUT; ++q)
              if (q == un) { sacc[q][0] += s1; sacc[q][1] += s2; }
          }
          const uint4 pk = Elem<T>::pack(o);
          out[idx] = pk;
          if (x == 0) out[idx + p.W] = pk;              // xp = W+1 mirrors pixel 0
          if (x == p.W - 1) out[idx - p.W] = pk;        // xp = 0 mirrors pixel W-1
        }
      }
    }
    if (p.stats != nullptr) {
#pragma unroll
      for (int u = 0; u < NUT; ++u) {
        const float a = warp_sum(sacc[u][0]), q = warp_sum(sacc[u][1]);
        if (lane == 0) { stat_s[ew][u][0] = a; stat_s[ew][u][1] = q; }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int units_here = NT / p.unit_ch;
      if (m < units_here * 2) {
        const int u = m >> 1, k = m & 1;
        const float tot = stat_s[0][u][k] + stat_s[1][u][k] + stat_s[2][u][k] + stat_s[3][u][k];
        const int unit = n0 / p.unit_ch + u;
        const int slot = yt * p.xtiles + xt;
        p.stats[((static_cast<size_t>(b) * kNU + unit) * p.slots + slot) * 2 + k] = tot;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Tr::TMEM_COLS>(tmem);
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

size_t conv_packed_weight_bytes(int dtype, int taps, int nt, int cin_pad, int cout_pad) {
  (void)nt;
  return static_cast<size_t>(taps) * cin_pad * cout_pad * dtype_size(dtype);
}

int conv_stat_slots(const ConvLaunch& l) { return (l.out.H / l.ht) * (l.out.W / 128); }

static int make_one_tmap(CUtensorMap* tm, int dtype, const PT& t, int box_w, int box_h, int box_planes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const int cw = dtype_cw(dtype);
  const cuuint64_t Wp = t.W + 2;
  cuuint64_t dims[5] = {(cuuint64_t)cw, Wp, (cuuint64_t)t.H, (cuuint64_t)(t.C / cw), (cuuint64_t)t.B};
  cuuint64_t strides[4] = {16, Wp * 16, (cuuint64_t)t.H * Wp * 16, (cuuint64_t)(t.C / cw) * t.H * Wp * 16};
  cuuint32_t box[5] = {(cuuint32_t)cw, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                   t.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

int conv_make_tmaps(ConvLaunch& l) {
  const int planes = 2 * ks_for(l.taps);
  const int bw = l.taps == 9 ? 130 : 128, bh = l.taps == 9 ? l.ht + 2 : l.ht;
  int rc = make_one_tmap(&l.tmap0, l.dtype, l.in0, bw, bh, planes);
  if (rc) return rc;
  if (l.in1.ptr) rc = make_one_tmap(&l.tmap1, l.dtype, l.in1, bw, bh, planes);
  else l.tmap1 = l.tmap0;
  return rc;
}

template <typename T, int NT, int HT, int TAPS, int KS, int STAGES, int MINB>
static cudaError_t launch_one(const ConvLaunch& l, cudaStream_t s) {
  using Tr = ConvTraits<T, NT, HT, TAPS, KS, STAGES>;
  auto kern = conv_umma_kernel<T, NT, HT, TAPS, KS, STAGES, MINB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tr::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  ConvParams p;
  p.tmap0 = l.tmap0; p.tmap1 = l.tmap1;
  p.wpacked = l.wpacked; p.bias = l.bias; p.residual = l.residual;
  p.out = l.out.ptr; p.out_nchw = l.out_nchw; p.stats = l.out_nchw ? nullptr : l.out.stats;
  p.B = l.out.B; p.H = l.out.H; p.W = l.out.W;
  p.cout = l.cout; p.cout_pad = l.cout_pad;
  p.nk = l.cin_pad / Tr::KCH;
  p.ksplit = l.in1.ptr ? l.in0.C / Tr::KCH : p.nk;
  p.xtiles = l.out.W / 128; p.ytiles = (l.out.H + HT - 1) / HT; p.ntiles = l.cout_pad / NT;
  p.unit_ch = l.cout / kNU > 0 ? l.cout / kNU : 1;
  p.slots = l.out.slots;
  p.scale = l.scale;
  const int grid = p.B * p.ytiles * p.xtiles * p.ntiles;
  kern<<<grid, 256, Tr::SMEM_BYTES, s>>>(p);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t dispatch(const ConvLaunch& l, cudaStream_t s) {
  if (l.taps == 9) {
    if (l.nt == 64 && l.ht == 4) return launch_one<T, 64, 4, 9, 1, 2, 2>(l, s);
    if (l.nt == 64 && l.ht == 2) return launch_one<T, 64, 2, 9, 1, 3, 2>(l, s);
    if (l.nt == 128 && l.ht == 2) return launch_one<T, 128, 2, 9, 1, 2, 2>(l, s);
    if (l.nt == 128 && l.ht == 1) return launch_one<T, 128, 1, 9, 1, 2, 2>(l, s);
    if (l.nt == 16 && l.ht == 4) return launch_one<T, 16, 4, 9, 1, 3, 2>(l, s);
  } else if (l.taps == 1) {
    if (l.nt == 64 && l.ht == 2) return launch_one<T, 64, 2, 1, 4, 2, 2>(l, s);
    if (l.nt == 128 && l.ht == 2) return launch_one<T, 128, 2, 1, 4, 2, 2>(l, s);
    if (l.nt == 128 && l.ht == 1) return launch_one<T, 128, 1, 1, 4, 3, 2>(l, s);
    if (l.nt == 64 && l.ht == 1) return launch_one<T, 64, 1, 1, 4, 3, 2>(l, s);
  }
  return cudaErrorInvalidConfiguration;
}

cudaError_t conv_launch(const ConvLaunch& l, cudaStream_t s) {
  if (l.out.W % 128 != 0 || l.cout_pad % l.nt != 0 || l.cin_pad % conv_stage_channels(l.dtype, l.taps) != 0)
    return cudaErrorInvalidValue;
  return l.dtype == kBF16 ? dispatch<__nv_bfloat16>(l, s) : dispatch<float>(l, s);
}

// ------------------------------------------------------------------------------------ weights
// dst[nt][ks][tap][plane][co][cw]  <-  w[co][ci][tap]   (zero padded)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int taps, int nt,
                                   int cout, int cin, int cin_pad, int cout_pad, int planes) {
  constexpr int CW = Elem<T>::CW;
  const size_t total = static_cast<size_t>(taps) * cin_pad * cout_pad;
  const int kch = planes * CW;
  const int nk = cin_pad / kch;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int cw = r % CW; r /= CW;
    const int co = r % nt; r /= nt;
    const int pl = r % planes; r /= planes;
    const int tap = r % taps; r /= taps;
    const int ks = r % nk; r /= nk;
    const int nti = static_cast<int>(r);
    const int ci = ks * kch + pl * CW + cw;
    const int o = nti * nt + co;
    float v = 0.f;
    if (ci < cin && o < cout) v = w[(static_cast<size_t>(o) * cin + ci) * taps + tap];
    if (sizeof(T) == 4) {  // tf32 operand: round to nearest instead of the tensor core's truncation
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
      v = __uint_as_float(r);
    }
    dst[i] = static_cast<T>(v);
  }
}

cudaError_t pack_conv_weight(int dtype, int taps, int nt, const float* w, int cout, int cin, int cin_pad,
                             int cout_pad, void* dst, cudaStream_t s) {
  const int planes = 2 * ks_for(taps);
  const size_t total = static_cast<size_t>(taps) * cin_pad * cout_pad;
  const int grid = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  if (dtype == kBF16)
    pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(dst), taps, nt, cout,
                                                           cin, cin_pad, cout_pad, planes);
  else
    pack_weight_kernel<float><<<grid, 256, 0, s>>>(w, static_cast<float*>(dst), taps, nt, cout, cin, cin_pad,
                                                   cout_pad, planes);
  return cudaGetLastError();
}

}  // namespace r2dm
