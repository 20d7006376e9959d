// Host-side launcher declarations for the hand-written sm_100a kernels (internal header).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace r2dm {

constexpr int kMaxConvCin = 1024;  // widest convolution input (incl. channel concat) the conv kernel is built for
constexpr int kNU = 8;  // GroupNorm statistic units per tensor (= gn_num_groups of the reference)

enum DType : int { kF32 = 0, kBF16 = 1 };

// cudaFuncSetAttribute and the SM count are per device: a launcher keeps one bit per device ordinal
// (one process may drive several GPUs, e.g. a model moved from cuda:0 to cuda:1).
inline bool first_use_on_this_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  const bool first = (mask & bit) == 0;
  mask |= bit;
  return first;
}
inline int dtype_size(int dt) { return dt == kBF16 ? 2 : 4; }
inline int dtype_cw(int dt) { return 16 / dtype_size(dt); }

// A planar-16 activation tensor (see common.cuh) plus its GroupNorm partial statistics.
struct PT {
  void* ptr = nullptr;
  int B = 0, C = 0, H = 0, W = 0;
  float* stats = nullptr;  // [B][kNU][slots][2] (sum, sum of squares), or null
  int slots = 0;
  size_t bytes(int dt) const { return static_cast<size_t>(B) * C * H * (W + 2) * dtype_size(dt); }
};

// ------------------------------------------------------------------ implicit-GEMM convolution
// GroupNorm / AdaGN (+SiLU) applied to the conv input inside the conv kernel (statistics come from
// in0.stats / in1.stats, i.e. from the kernel that produced the input tensors).
struct ConvXform {
  int enabled;
  int silu;
  int groups;
  float eps;
  const float* gamma; const float* beta;  // affine GN, or null for AdaGN
  const float* film;                      // AdaGN table [rows][film_stride]; scale at film_off, shift at +C
  int film_stride, film_off;
  const int* step_ptr; int rows_per_step, row_batch_stride;
  int c0_real;                            // real channel count of in0 if it is zero padded (0 = in0.C)
};

struct ConvLaunch {
  int dtype;          // kF32 (tf32 tensor cores) or kBF16
  int taps;           // 9 (3x3 ring conv) or 1 (1x1 conv / linear over tokens)
  int nt, ht;         // N tile (output channels per CTA) and output rows per tile
  PT in0, in1;        // input(s); in1.ptr == nullptr unless the input is a channel concat
  int cin_pad;        // total input channels incl. zero padding (multiple of the stage K)
  const void* wpacked;  // packed weights (pack_conv_weight)
  const float* bias;  // [cout] fp32 (zero padded to cout_pad)
  int cout, cout_pad;
  PT out;             // output tensor (planar-16), unless out_nchw
  const void* residual;  // planar-16 tensor added before scaling, or null
  float scale;        // applied after bias (+ residual)
  float* out_nchw;    // if non-null: write fp32 [B][cout][H][W] instead of `out`
  // folded skip projection (3x3 launches only): out += conv1x1(sk0 ++ sk1) + bias2, accumulated as extra K
  // stages of the same tile from the RAW (untransformed) tensors; sk0.ptr == nullptr: none
  PT sk0, sk1;
  int cin2_pad;
  const void* w2packed;   // pack_conv_weight(taps = 1, ..., planes = conv_skip_planes(nt))
  const float* bias2;
  ConvXform xf;       // fused input normalisation (enabled = 0: plain convolution)
  unsigned long long* ktime;  // developer: device [2] receiving (first CTA start, last CTA end), globaltimer ns
  int reverse;        // tile order back to front (alternated between consecutive launches, see conv_umma.cu)
  int round_out;      // fp32 engine: store the output rounded to nearest tf32 (its only consumer is a tensor-core operand)
  // 1x1 launches only (point-wise MLPs over the range image, pointnet.cu): ReLU after bias / scale, and a running
  // per-(image, channel) maximum over all pixels into colmax [B][cout_pad] (float bits, initialised to -inf);
  // with colmax set and out.ptr == nullptr the output tensor itself is not stored
  int relu;
  float* colmax;
  CUtensorMap tmap0, tmap1, tmap2, tmap3;
};
// Fills l.tmap0/tmap1 for the current in0/in1 pointers.  Returns 0 on success.
int conv_make_tmaps(ConvLaunch& l);
int conv_stage_channels(int dtype, int taps);  // K per pipeline stage
int conv_skip_planes(int nt);                  // channel planes per skip stage of a 3x3 launch with N tile nt
size_t conv_packed_weight_bytes(int dtype, int taps, int nt, int cin_pad, int cout_pad);
int conv_stat_slots(const ConvLaunch& l);
cudaError_t conv_launch(const ConvLaunch& l, cudaStream_t s);
// Several consecutive 3x3 convolutions of one resolution level in ONE persistent launch (conv_chain.cu): same tile
// geometry, fused GroupNorm on every layer.  done: [n][B] ints of scratch; batch_split: images in the first group.
constexpr int kMaxChainLayers = 6;
bool conv_chain_supported(const ConvLaunch& l);
cudaError_t conv_chain_launch(const ConvLaunch* const* ls, int n, int* done, int batch_split, cudaStream_t s);
// Launch with programmatic dependent launch (PDL) allowed: the kernel may become resident while its stream
// predecessor drains; it must execute griddepcontrol.wait (pdl_wait) before touching anything a predecessor
// wrote or still reads.  All kernels of the sampling loop use it, so a CUDA graph of the loop has
// programmatic edges end to end (R2DM_PDL=0 switches it off).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// process-wide developer options (r2dm_set_option / R2DM_OPT_<NAME> in the environment)
int get_option(const char* name, int dflt);
int set_option(const char* name, int value);
// developer timeline of CTA 0 of subsequent conv launches: buf[4 roles][cap] (globaltimer ns), or null
void conv_set_trace(unsigned long long* buf, int cap);
// w: fp32 [cout][cin][k][k] (OIHW), k*k == taps
// planes: channel planes per K stage (0 = the kernel's own stage width for `taps`)
cudaError_t pack_conv_weight(int dtype, int taps, int nt, const float* w, int cout, int cin,
                             int cin_pad, int cout_pad, void* dst, cudaStream_t s, int planes = 0);

// ------------------------------------------------------------------ layout conversion
cudaError_t pack_nchw(int dtype, const float* src, int B, int Csrc, int H, int W, PT dst,
                      int c_off, cudaStream_t s);
cudaError_t unpack_nchw(int dtype, PT src, float* dst, int c_off, int Cdst, cudaStream_t s);
// x [B][Cx][H][W] fp32 and constant enc [Ce][H][W] fp32 -> planes [plane_begin, plane_end) of dst
cudaError_t pack_input(int dtype, const float* x, int Cx, const float* enc, int Ce, PT dst,
                       int plane_begin, int plane_end, cudaStream_t s);

// ------------------------------------------------------------------ GroupNorm / AdaGN (+SiLU)
struct GnApply {
  int dtype;
  PT src0, src1;      // channel concat [src0, src1] (src1.ptr null if single)
  PT dst;
  int groups;
  float eps;
  const float* gamma; const float* beta;  // affine GN, or null for AdaGN
  const float* film;  // AdaGN: table [rows][film_stride]; scale at film_off, shift at +C
  int film_stride, film_off;
  const int* step_ptr;  // device int: current table row base (null -> 0)
  int row_batch_stride; // row = step * rows_per_step + b * row_batch_stride
  int rows_per_step;
  int silu;
};
cudaError_t gn_apply_launch(const GnApply& g, cudaStream_t s);
// statistics of a tensor that was not produced by a stats-emitting kernel (tests / external input)
cudaError_t tensor_stats_launch(int dtype, PT t, cudaStream_t s);
int tensor_stats_slots(int dtype, const PT& t);

// ------------------------------------------------------------------ FIR resampling
cudaError_t down2_launch(int dtype, PT src, PT dst, cudaStream_t s);  // dst gets stats
int down2_stat_slots(int dtype, const PT& dst);
cudaError_t up2_launch(int dtype, PT src, PT dst, cudaStream_t s);

// ------------------------------------------------------------------ attention core
// qkv: planar-16 tensor with 3E channels ([q;k;v]); out: E channels. softmax(q k^T / sqrt(hd)) v
cudaError_t attention_launch(int dtype, PT qkv, PT out, int heads, cudaStream_t s);   // exact fp32 FMA-pipe version
// bf16 tensor-core version (tcgen05): needs a TMA map of the qkv tensor
int attention_key_tile(int dtype);   // keys per tile of the tensor-core kernel (128 bf16 / 64 tf32)
int attention_make_tmaps(CUtensorMap* tm_q, CUtensorMap* tm_kv, int dtype, const PT& qkv, int heads);
cudaError_t attention_umma_launch(int dtype, PT qkv, PT out, int heads, const CUtensorMap& tm_q,
                                  const CUtensorMap& tm_kv, cudaStream_t s);

// ------------------------------------------------------------------ conditioning table
struct CondEmbed {
  const float* cond; int rows;        // [rows] network conditions (log-SNR or integer step)
  int base_ch, temb_ch;               // sinusoid width, MLP width
  const float* w1; const float* b1;   // [temb][base], [temb]
  const float* w2; const float* b2;   // [temb][temb], [temb]
  const float* wf; const float* bf;   // all AdaGN projections stacked: [F][temb], [F]
  int F;
  float* temb_scratch;                // [rows][temb]
  float* film;                        // [rows][F]
};
cudaError_t cond_embed_launch(const CondEmbed& c, cudaStream_t s);

// ------------------------------------------------------------------ sampler elementwise
// x0 = clamp(ux*x + up*pred); x' = kx*x + k0*x0 + kn*noise ; coef rows of 5 floats, row index =
// step*rows_per_step + b*row_batch_stride.  Optional RePaint blend with the re-noised known image:
// x' = mask*(qa*known + qs*noise2) + (1-mask)*x'   (coef row has 7 floats then).
// Noise drawn inside the kernel, bit-identical to torch.randn(per_sample elements, generator=cuda_gen_b)
// (models/diffusion/base.py:71-94 with a list of per-sample CUDA generators): element l of the tensor is
// component (l / threads) % 4 of curand_normal4 of the Philox4x32-10 stream (seed_b, subsequence
// l % threads, offset_b + offset_per_draw * draw + 4 * (l / threads / 4)), exactly what ATen's
// distribution_nullary_kernel computes for its launch width `threads`.
struct PhiloxDraw {
  const unsigned long long* seeds;    // [B] device; null = noise comes from memory
  const unsigned long long* offsets;  // [B] device: generator offset before the first draw of the loop
  const int* ctr0; const int* ctr1;   // draw number = mul0 * *ctr0 + mul1 * *ctr1 + index (null -> 0)
  int mul0, mul1;
  unsigned offset_per_draw, threads;
};

struct SamplerUpdate {
  float* x; const float* pred; const float* noise;
  const float* coef; int coef_cols;
  const int* step_ptr; int rows_per_step, row_batch_stride;
  float clip;          // <= 0: no clipping
  const float* known; const float* mask; const float* noise2;  // RePaint (or null)
  float* x_out;        // may alias x
  int B; size_t per_sample;
  PhiloxDraw ph;       // ph.seeds != null: noise / noise2 are generated (draw indices below), not read
  int draw_noise, draw_noise2;
};
cudaError_t sampler_update_launch(const SamplerUpdate& u, cudaStream_t s);
// y = a*x + c*noise  (q_step / q_step_from_x_0), coefficients on device: ac[row][2],
// row = step*rows_per_step + b*row_batch_stride (step_ptr may be null -> 0)
cudaError_t axpby_launch(const float* x, const float* noise, const float* ac, float* y, int B,
                         size_t per_sample, const int* step_ptr, int rows_per_step, int row_batch_stride,
                         cudaStream_t s, const PhiloxDraw* ph = nullptr, int draw = 0);
// out[b] = the `draw`-th torch.randn of sample b's generator (tests, x_T)
cudaError_t philox_normal_launch(float* out, const PhiloxDraw& ph, int draw, int B, size_t per_sample,
                                 cudaStream_t s);
cudaError_t advance_step_launch(int* step_ptr, int delta, cudaStream_t s);
// [depth, x, y, z, reflectance] from a sample in [-1, 1] (sample_and_save.py:52-57)
cudaError_t lidar_postprocess_launch(const float* sample, const float* angles, float* out, int B,
                                     int H, int W, int depth_format, float min_depth,
                                     float max_depth, cudaStream_t s);

// PointNet feature extractor helpers (pointnet.cu; metrics/extractor/pointnet.py)
cudaError_t fill_launch(float* p, float v, size_t n, cudaStream_t s);
cudaError_t dense_launch(const float* in, int in_stride, const float* w, const float* bias, float* out, int out_stride,
                         int B, int K, int N, int relu, cudaStream_t s);
cudaError_t point_transform_launch(const float* x, const float* trans, float* y, int B, int N, cudaStream_t s);
// caller-side consumers of the generated point clouds (render.cu; utils/render.py, metrics/bev.py)
cudaError_t render_splat_launch(const float* points, const float* colors, const float* R, const float* t,
                                float* acc, float* out, int B, int N, int size, float focal, cudaStream_t s);
cudaError_t rasterize_launch(const float* coords, const float* values, float* out, int B, int N, int C, int H, int W,
                             cudaStream_t s);
cudaError_t surface_normal_launch(const float* points, float* out, int B, int H, int W, int d, int mode,
                                  cudaStream_t s);
cudaError_t bev_histogram_launch(const float* points, const float* edges, unsigned int* counts, float* hist, int B,
                                 int N, int bins, float min_depth, float max_depth, cudaStream_t s);

}  // namespace r2dm
