/* r2dm_b200 — C ABI of the B200-native R2DM sampling hot path.
 *
 * The reference (kazuto1011/r2dm) is pure PyTorch and has no FFI layer; its boundary for this path
 * is the Python object protocol of models/diffusion/{base,continuous_time,discrete_time}.py and
 * models/efficient_unet.py.  The entry points below are what a binding for that path calls; the
 * Python shim in r2dm_b200/ (ctypes) mirrors the reference classes on top of them.  Each function
 * cites the reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions: all pointers are CUDA device pointers owned by the caller (PyTorch) unless marked
 * "host"; tensors are contiguous fp32 NCHW at the boundary; work is enqueued on `stream`
 * (a cudaStream_t passed as void*) and never synchronises; every function returns 0 on success or
 * a negative code, with a message available from r2dm_last_error().  No function allocates device
 * memory: weights live in a caller-provided arena, activations in a caller-provided workspace.
 *
 * Devices and threads: a handle, its arena / workspace and every tensor passed with it belong to ONE CUDA
 * device, which must be current in the calling thread.  One process may use several devices (per-device
 * kernel attributes and SM counts are tracked inside the library); the intended deployment is one
 * process per GPU (r2dm_b200/parallel.py).  r2dm_last_error() is thread-local; r2dm_set_option() and the
 * r2dm_debug_* hooks are process-wide and not synchronised.
 */
#ifndef R2DM_B200_H
#define R2DM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r2dm_model* r2dm_handle;

enum { R2DM_F32 = 0, R2DM_BF16 = 1 }; /* compute/storage type of the U-Net: tf32 or bf16 tensor cores */

/* EfficientUNet constructor arguments (models/efficient_unet.py:194-209, utils/inference.py:38-51). */
typedef struct {
  int in_channels;          /* image channels (depth, reflectance) */
  int height, width;        /* resolution, e.g. 64 x 1024 */
  int base_channels;
  int temb_channels;        /* 0 -> 4 * base_channels */
  int channel_multiplier[4];
  int num_residual_blocks[4];
  int gn_num_groups;        /* must be 8 */
  float gn_eps;
  int attn_num_heads;
  int extra_channels;       /* channels of the constant coordinate encoding appended to the input */
  float residual_scale;     /* ResidualBlock / SelfAttentionBlock `scale` buffer (1/sqrt 2) */
  int dtype;                /* R2DM_F32 or R2DM_BF16 */
} r2dm_config;

const char* r2dm_last_error(void);
int r2dm_version(void);

/* --- lifetime ---------------------------------------------------------------------------------- */
int r2dm_create(const r2dm_config* cfg, r2dm_handle* out);
int r2dm_destroy(r2dm_handle h);

/* --- weights: replaces nn.Module.load_state_dict (utils/inference.py:80-81) ------------------------
 * `name` is the reference state-dict key without the leading "model." (e.g.
 * "d_block1.residual_blocks.0.conv1.weight"); `src` is the fp32 device tensor.  Convolution and
 * projection weights are re-packed into the tensor-core operand layout inside the arena.  The
 * constant coordinate encoding (models/encoding.py) is loaded under the name "coords_encoding.table"
 * with shape [extra_channels, H, W].  Unknown names return 1 (ignored buffer), not an error. */
size_t r2dm_weight_arena_bytes(r2dm_handle h);
int r2dm_bind_weight_arena(r2dm_handle h, void* arena, size_t bytes);
int r2dm_load_tensor(r2dm_handle h, const char* name, const float* src, const int64_t* shape, int ndim,
                     void* stream);
int r2dm_missing_tensors(r2dm_handle h, char* names_out, size_t names_cap); /* returns count */

/* --- workspace for a given batch size ------------------------------------------------------------ */
size_t r2dm_workspace_bytes(r2dm_handle h, int batch);
int r2dm_bind_workspace(r2dm_handle h, void* workspace, size_t bytes, int batch, void* stream);

/* --- conditioning: time embedding MLP + all AdaGN projections --------------------------------------
 * (models/efficient_unet.py:232-237,273-275; models/ops.py:14-29,190-198).  For `rows` network
 * conditions (log-SNR values or integer steps) writes film[rows][r2dm_film_width()].
 * scratch: rows * temb_channels floats. */
int r2dm_film_width(r2dm_handle h);
int r2dm_cond_embed(r2dm_handle h, const float* cond, int rows, float* scratch, float* film, void* stream);

/* --- EfficientUNet.forward (models/efficient_unet.py:269-295) --------------------------------------
 * x, pred: [B][in_channels][H][W] fp32.  Sample b uses film row
 *   (step_ptr ? *step_ptr : 0) * rows_per_step + b * row_batch_stride
 * so a captured CUDA graph can walk a precomputed per-step table with a device-side counter. */
int r2dm_unet_forward(r2dm_handle h, const float* x, const float* film, const int* step_ptr,
                      int rows_per_step, int row_batch_stride, float* pred, void* stream);
int r2dm_num_launches(r2dm_handle h); /* kernels enqueued by one r2dm_unet_forward */
/* Measurement aid (bench.py): one eager forward with a CUDA-event pair around every launch; this
 * call SYNCHRONISES and is never used on the product path.  Host arrays of capacity `cap` receive
 * per launch: kind (0 pack_input, 1 conv3x3, 2 conv1x1, 3 GroupNorm/AdaGN apply, 4 down2, 5 up2,
 * 6 attention), device milliseconds, algorithmic FLOPs and algorithmic bytes.  Returns #launches. */
int r2dm_profile_forward(r2dm_handle h, const float* x, const float* film, float* pred, void* stream,
                         int cap, int* kind, float* ms, double* flops, double* bytes);

/* Measurement aid (bench.py): enqueue only the launches of one forward whose kind bit is set (bit k = kind k
 * above), in program order, on the buffers of the last real forward - e.g. the 56 3x3 convolutions back to
 * back inside a CUDA graph, exactly as the product launches them.  Results are meaningless. */
int r2dm_debug_forward_kinds(r2dm_handle h, const float* x, const float* film, float* pred,
                             unsigned kind_mask, void* stream);

/* --- sampler step arithmetic (models/diffusion/continuous_time.py:208-229, 296-299;
 *     discrete_time.py:140-179).  coef rows: {ux, up, kx, k0, kn [, qa, qs]}:
 *       x0  = clamp(ux*x + up*pred, +-clip)        (clip <= 0 disables)
 *       x'  = kx*x + k0*x0 + kn*noise
 *       x'  = mask*(qa*known + qs*noise2) + (1-mask)*x'     when known != NULL (RePaint)
 * row index as in r2dm_unet_forward. */
int r2dm_sampler_update(float* x_out, const float* x, const float* pred, const float* noise,
                        const float* coef, int coef_cols, const int* step_ptr, int rows_per_step,
                        int row_batch_stride, float clip, const float* known, const float* mask,
                        const float* noise2, int batch, size_t per_sample, void* stream);
/* y[b] = ac[b][0]*x[b] + ac[b][1]*noise[b]  (q_step / q_step_from_x_0, continuous_time.py:169-190) */
int r2dm_axpby(float* y, const float* x, const float* noise, const float* ac, int batch,
               size_t per_sample, void* stream);
/* same with a per-step table: row = (step_ptr ? *step_ptr : 0)*rows_per_step + b*row_batch_stride,
 * y[b] = table[row][0]*x[b] + table[row][1]*noise[b]  (RePaint re-noising under a CUDA graph) */
int r2dm_axpby_table(float* y, const float* x, const float* noise, const float* table, const int* step_ptr,
                     int rows_per_step, int row_batch_stride, int batch, size_t per_sample, void* stream);
int r2dm_advance_step(int* step_ptr, int delta, void* stream);

/* --- noise drawn on the device (models/diffusion/base.py:71-94 with `rng` = a list of per-sample CUDA
 * generators): the same arithmetic as above with the noise generated inside the kernel, bit-identical
 * to `torch.randn(per_sample elements, generator=g_b)` on this device, so that a whole sampling loop
 * replays as CUDA graphs without host-side draws in between.  Sample b's j-th draw of the loop reads
 * the Philox4x32-10 stream (seeds[b], offsets[b] + offset_per_draw * j), j = mul0 * *ctr0 +
 * mul1 * *ctr1 + <draw index of the call> (NULL counters read as 0); `threads` / `offset_per_draw`
 * are ATen's launch width and per-call offset increment for a tensor of per_sample elements
 * (r2dm_b200/diffusion.py::_torch_randn_geometry). */
typedef struct {
  const uint64_t* seeds;    /* device [batch] */
  const uint64_t* offsets;  /* device [batch] */
  const int* ctr0; const int* ctr1;
  int mul0, mul1;
  uint32_t offset_per_draw, threads;
} r2dm_philox;
/* r2dm_sampler_update with generated noise; draw_noise2 (known-region draw, RePaint) is ignored when known == NULL */
int r2dm_sampler_update_philox(float* x_out, const float* x, const float* pred, const float* coef, int coef_cols,
                               const int* step_ptr, int rows_per_step, int row_batch_stride, float clip,
                               const float* known, const float* mask, const r2dm_philox* philox, int draw_noise,
                               int draw_noise2, int batch, size_t per_sample, void* stream);
/* r2dm_axpby_table with generated noise */
int r2dm_axpby_table_philox(float* y, const float* x, const float* table, const int* step_ptr, int rows_per_step,
                            int row_batch_stride, const r2dm_philox* philox, int draw, int batch,
                            size_t per_sample, void* stream);
/* out[b][per_sample] = that draw itself (x_T, tests) */
int r2dm_philox_normal(float* out, const r2dm_philox* philox, int draw, int batch, size_t per_sample, void* stream);

/* --- LiDAR post-processing (sample_and_save.py:52-57, utils/lidar.py:49-70,99-120):
 * sample [B][2][H][W] in [-1,1] -> out [B][5][H][W] = depth, x, y, z, reflectance.
 * depth_format: 0 log_depth, 1 inverse_depth, 2 depth.  angles: [2][H][W] (elevation, azimuth). */
int r2dm_lidar_postprocess(const float* sample, const float* angles, float* out, int batch, int H,
                           int W, int depth_format, float min_depth, float max_depth, void* stream);

/* --- caller-side consumers of the generated point clouds -------------------
 * utils/render.py:32-80 render_point_clouds: points [B][N][3] (already divided by max_depth by the caller),
 * colors [B][N][3] or NULL (= white), R [3][3] row-major applied as p @ R or NULL, t [3] or NULL ->
 * out [B][3][size][size].  acc: scratch of B*size*size*16 bytes (weighted colour / weight sums). */
int r2dm_render_point_clouds(const float* points, const float* colors, const float* R, const float* t, float* acc,
                             float* out, int batch, int num_points, int size, float focal_length, void* stream);
/* utils/render.py:83-142 bilinear_rasterizer: coords [B][N][2] = (row, column), values [B][N][C] ->
 * out [B][C][H][W] (four-neighbour scatter-add, weights below 1e-3 and neighbours outside dropped). */
int r2dm_bilinear_rasterize(const float* coords, const float* values, float* out, int batch, int num_points,
                            int channels, int H, int W, void* stream);
/* utils/render.py:145-234 estimate_surface_normal: points [B][3][H][W] -> unit normals [B][3][H][W];
 * d = neighbour distance, mode 0 "closest", 1 "mean". */
int r2dm_surface_normal(const float* points, float* out, int batch, int H, int W, int d, int mode, void* stream);
/* metrics/bev.py:5-24 point_cloud_to_histogram, batched: points [B][N][3], edges [bins+1] (bin edges of both
 * axes, as torch.histogramdd builds them) -> hist [B][bins][bins] (fp32 counts of x-bin, y-bin);
 * counts: scratch of B*bins*bins uint32. */
int r2dm_bev_histogram(const float* points, const float* edges, unsigned int* counts, float* hist, int batch,
                       int num_points, int bins, float min_depth, float max_depth, void* stream);

/* metrics/extractor/pointnet.py:61-81 PointNet1.forward (the FPD feature extractor of evaluate.py): points
 * [B][3][N] (N a multiple of 128; the 64 x 1024 range image flattened) -> features [B][1024 + 512 + 256 + k].
 * The point-wise layers run as 1x1 tensor-core convolutions with the global max pool fused into the last one.
 * Weights are fp32 device pointers with BatchNorm (eval) folded in, in the order stn.conv1-3, stn.fc1-3 (the
 * identity of pointnet.py:32 added to the last bias), feat.conv1-3, fc1-3; weight[i] is [out][in] row-major. */
typedef struct {
  const float* weight[12];
  const float* bias[12];
  int num_classes; /* k of the classifier head (16 for the ShapeNet checkpoint the reference downloads) */
} r2dm_pointnet_weights;
size_t r2dm_pointnet_scratch_bytes(int dtype, int batch, int num_points);
int r2dm_pointnet_features(int dtype, const float* points, const r2dm_pointnet_weights* w, float* features,
                           int batch, int num_points, void* scratch, size_t scratch_bytes, void* stream);

/* --- single-operator entry points (used by the parity tests; same kernels as the network) ----------
 * All take fp32 NCHW tensors and a scratch buffer for the packed intermediates. */
size_t r2dm_op_scratch_bytes(int batch, int max_channels, int H, int W);
/* ops.Conv2d ring 3x3 (taps=9) or 1x1 (taps=1); w OIHW; residual may be NULL; y = (conv+bias+res)*scale */
int r2dm_op_conv(int dtype, int taps, const float* x, const float* w, const float* bias,
                 const float* residual, float scale, float* y, int B, int Cin, int Cout, int H, int W,
                 void* scratch, size_t scratch_bytes, void* stream);
/* the fused form the network actually runs: y = conv(act(GN(x))) + bias, GroupNorm/AdaGN(+SiLU) applied
 * to the operand tile inside the conv kernel.  film: [B][2*Cin] = [scale || shift] (AdaGN) or NULL. */
int r2dm_op_gn_conv(int dtype, int taps, const float* x, const float* gamma, const float* beta,
                    const float* film, float eps, int silu, const float* w, const float* bias, float* y,
                    int B, int Cin, int Cout, int H, int W, void* scratch, size_t scratch_bytes, void* stream);
/* ResidualBlock tail with the skip projection folded into conv2 (efficient_unet.py:99-110 for blocks with a skip):
 * y = (conv3x3(silu(adagn(h, film))) + bias + conv1x1(xs, w2) + bias2) * scale; h [B][C][H][W], xs [B][Cs][H][W] */
int r2dm_op_gn_conv_skip(int dtype, const float* h, const float* film, float eps, const float* w, const float* bias,
                         const float* xs, const float* w2, const float* bias2, float scale, float* y, int B, int C,
                         int Cs, int H, int W, void* scratch, size_t scratch_bytes, void* stream);
/* GroupNorm(8 groups) [+ FiLM: y = gn(x)*(1+fs)+fb when film_scale != NULL, per sample [B][C]] [+ SiLU] */
int r2dm_op_groupnorm(int dtype, const float* x, const float* gamma, const float* beta,
                      const float* film_scale_shift, float eps, int silu, float* y, int B, int C, int H,
                      int W, void* scratch, size_t scratch_bytes, void* stream);
/* ops.Resample: dir = +2 (up) or -2 (down) */
int r2dm_op_resample(int dtype, int dir, const float* x, float* y, int B, int C, int H, int W,
                     void* scratch, size_t scratch_bytes, void* stream);
/* attention core on packed qkv [B][3E][H][W] -> [B][E][H][W] */
int r2dm_op_attention(int dtype, const float* qkv, float* y, int B, int E, int heads, int H, int W,
                      void* scratch, size_t scratch_bytes, void* stream);

/* copy a named intermediate activation of the last forward to fp32 NCHW (debugging / tests);
 * returns channel count via *C_out etc.  Names: "in_conv", "<block>", "<block>.rb<i>". */
/* developer aid: record a per-role globaltimer timeline of CTA 0 of subsequent conv launches into
 * buf[4][cap] (uint64 ns; roles: producer, MMA, transform, epilogue); NULL disables. */
/* Process-wide developer options, read when a model is created / an r2dm_op_* call is planned
 * (see INTEGRATION.md "Options"); unknown names return an error. */
int r2dm_set_option(const char* name, int value);
int r2dm_debug_set_trace(void* buf, int cap);
/* developer aid: every convolution launch of the forward records (first CTA start, last CTA end) in
 * globaltimer ns into buf[launch index][2] (device uint64; works inside CUDA graphs); NULL disables. */
int r2dm_debug_set_ktime(r2dm_handle h, void* buf);
int r2dm_debug_tensor(r2dm_handle h, const char* name, float* out, int* C, int* H, int* W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2DM_B200_H */
