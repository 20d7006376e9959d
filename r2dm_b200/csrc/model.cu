// Host runtime + C ABI: weight arena, workspace planning and the per-forward launch program of the
// EfficientUNet (models/efficient_unet.py:188-295) built from the kernels in this directory.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/r2dm_b200.h"
#include "kernels.h"

using namespace r2dm;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return fail(-2, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct ConvW {
  std::string name;
  int taps = 9, cin = 0, cout = 0, cin_pad = 0, cout_pad = 0, nt = 64;
  size_t w_off = 0, b_off = 0;
  bool w_ok = false, b_ok = false;
  int sk_planes = 0;   // > 0: a skip projection folded into its block's conv2 (packed with this stage width)
};
struct RawW {  // fp32 tensor copied verbatim into the arena at dst_off (+ row offset for stacked)
  std::string name;
  size_t off = 0, numel = 0;
  bool ok = false;
};

struct BlockSpec {
  std::string name;
  int cin, cout, nres, down, up, attn;
};

// A planned activation tensor.
struct TRef { int id = -1; };

struct Op {
  enum Kind { PACK_INPUT, CONV, GN, DOWN, UP, ATTN } kind;
  ConvLaunch conv;
  GnApply gn;
  PT a, b;
  int heads = 0;
  int conv_w = -1;        // index into convs (weights/bias resolved at bind)
  int skip_w = -1;        // index of the skip projection folded into this launch, or -1
  int gn_gamma = -1;      // raw index of gamma (beta = +1), or -1 for AdaGN
  bool is_output = false; // network output conv (writes pred NCHW)
  bool xf_film = false;   // conv with fused AdaGN: film pointers patched per forward
  CUtensorMap attn_tmap, attn_tmap_kv;  // ATTN: TMA maps of the packed qkv tensor (Q tile / K, V tiles)
  bool attn_exact = false;              // fp32 engine with option attn_exact: FMA-pipe kernel instead of kind::tf32
  int chain_len = 0;                    // CONV: > 0 = this op and the next chain_len - 1 ops run as ONE conv_chain launch
  int* chain_done = nullptr;            //       its [chain_len][B] tile counters (workspace tail)
};

}  // namespace

struct r2dm_model {
  r2dm_config cfg;
  int dtype, cw, T, F = 0;
  int C[5];
  std::vector<BlockSpec> blocks;
  std::vector<ConvW> convs;
  std::vector<RawW> raws;
  std::map<std::string, int> conv_by_name, raw_by_name;
  std::map<std::string, std::pair<int, int>> film_rows;  // proj name -> (row offset, rows)
  size_t arena_bytes = 0;
  uint8_t* arena = nullptr;
  int raw_w1, raw_b1, raw_w2, raw_b2, raw_wf, raw_bf, raw_enc;
  int cin0_pad;
  bool keep_all = false;

  // workspace / program
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  int batch = 0;
  std::vector<Op> prog;
  std::map<std::string, PT> named;
  int n_launches = 0;

  float* raw_ptr(int i) const { return reinterpret_cast<float*>(arena + raws[i].off); }
};

namespace {

int add_raw(r2dm_model* m, const std::string& name, size_t numel) {
  RawW r;
  r.name = name;
  r.numel = numel;
  r.off = m->arena_bytes;
  m->arena_bytes += align_up(numel * sizeof(float), 256);
  m->raws.push_back(r);
  m->raw_by_name[name] = static_cast<int>(m->raws.size()) - 1;
  return static_cast<int>(m->raws.size()) - 1;
}

}  // namespace
namespace r2dm {
// Developer options: name -> value; the environment (R2DM_OPT_<NAME>, upper case) seeds a name on first use.
static std::map<std::string, int>& option_map() { static std::map<std::string, int> m; return m; }
static const char* const kOptionNames[] = {"serpentine", "ht1_max_tiles", "max_stages", "fold_skip", "attn_exact", "compact_grid", "prefetch_w", "chain", "chain_mask", "chain_nosplit", "chain_noclamp", "chain_rot", "unfuse_hw", "l2_evict_first"};
int get_option(const char* name, int dflt) {
  auto& m = option_map();
  auto it = m.find(name);
  if (it != m.end()) return it->second;
  std::string env = "R2DM_OPT_";
  for (const char* c = name; *c; ++c) env += static_cast<char>(toupper(*c));
  const char* e = getenv(env.c_str());
  const int v = e ? atoi(e) : dflt;
  m[name] = v;
  return v;
}
int set_option(const char* name, int value) {
  for (const char* n : kOptionNames)
    if (std::string(n) == name) { option_map()[name] = value; return 0; }
  return -1;
}
}  // namespace r2dm
namespace {
// N tile: 128 / 64 / 16 output channels per tile
int pick_nt(int cout) {
  if (cout <= 16) return 16;
  return cout % 128 == 0 ? 128 : 64;
}

int add_conv(r2dm_model* m, const std::string& name, int taps, int cin, int cout, bool folded_skip = false) {
  ConvW c;
  c.name = name;
  c.taps = taps; c.cin = cin; c.cout = cout;
  c.nt = pick_nt(cout);
  c.sk_planes = folded_skip ? conv_skip_planes(c.nt) : 0;
  c.cin_pad = round_up(cin, folded_skip ? c.sk_planes * dtype_cw(m->dtype) : conv_stage_channels(m->dtype, taps));
  c.cout_pad = round_up(cout, c.nt);
  c.w_off = m->arena_bytes;
  m->arena_bytes += align_up(conv_packed_weight_bytes(m->dtype, taps, c.nt, c.cin_pad, c.cout_pad), 256);
  c.b_off = m->arena_bytes;
  m->arena_bytes += align_up(static_cast<size_t>(c.cout_pad) * sizeof(float), 256);
  m->convs.push_back(c);
  m->conv_by_name[name] = static_cast<int>(m->convs.size()) - 1;
  return static_cast<int>(m->convs.size()) - 1;
}

// Mirrors the module tree of EfficientUNet.__init__ (efficient_unet.py:212-267).
int plan_weights(r2dm_model* m) {
  const r2dm_config& c = m->cfg;
  m->C[0] = c.base_channels;
  for (int i = 0; i < 4; ++i) m->C[i + 1] = c.base_channels * c.channel_multiplier[i];
  const int* C = m->C;
  const int* N = c.num_residual_blocks;
  m->blocks = {
      {"d_block1", C[0], C[1], N[0], 1, 1, 0}, {"d_block2", C[1], C[2], N[1], 2, 1, 0},
      {"d_block3", C[2], C[3], N[2], 2, 1, 0}, {"d_block4", C[3], C[4], N[3], 2, 1, 1},
      {"u_block4", C[4], C[3], N[3], 1, 2, 1}, {"u_block3", C[3] + C[3], C[2], N[2], 1, 2, 0},
      {"u_block2", C[2] + C[2], C[1], N[1], 1, 2, 0}, {"u_block1", C[1] + C[1], C[0], N[0], 1, 1, 0},
  };
  const int T = m->T;
  m->raw_w1 = add_raw(m, "time_embedding.1.weight", static_cast<size_t>(T) * c.base_channels);
  m->raw_b1 = add_raw(m, "time_embedding.1.bias", T);
  m->raw_w2 = add_raw(m, "time_embedding.3.weight", static_cast<size_t>(T) * T);
  m->raw_b2 = add_raw(m, "time_embedding.3.bias", T);
  m->raw_enc = c.extra_channels > 0
                   ? add_raw(m, "coords_encoding.table", static_cast<size_t>(c.extra_channels) * c.height * c.width)
                   : -1;
  add_conv(m, "in_conv", 9, c.in_channels + c.extra_channels, C[0]);
  int F = 0;
  for (const BlockSpec& b : m->blocks) {
    if (b.down > 1) add_conv(m, b.name + ".downsample.0", 9, b.cin, b.cout);
    for (int i = 0; i < b.nres; ++i) {
      const int ci = (i != 0 || b.down > 1) ? b.cout : b.cin;
      const std::string p = b.name + ".residual_blocks." + std::to_string(i);
      add_raw(m, p + ".norm1.weight", ci);
      add_raw(m, p + ".norm1.bias", ci);
      add_conv(m, p + ".conv1", 9, ci, b.cout);
      m->film_rows[p + ".norm2.proj.1"] = {F, 2 * b.cout};
      F += 2 * b.cout;
      add_conv(m, p + ".conv2", 9, b.cout, b.cout);
      if (ci != b.cout) add_conv(m, p + ".skip", 1, ci, b.cout, get_option("fold_skip", 1) != 0);
    }
    if (b.attn) {
      const std::string p = b.name + ".self_attn_block";
      add_raw(m, p + ".norm.weight", b.cout);
      add_raw(m, p + ".norm.bias", b.cout);
      add_conv(m, p + ".attn.in_proj", 1, b.cout, 3 * b.cout);
      add_conv(m, p + ".attn.out_proj", 1, b.cout, b.cout);
    }
    if (b.up > 1) add_conv(m, b.name + ".upsample.1", 9, b.cout, b.cout);
  }
  add_conv(m, "out_conv", 9, C[0], c.in_channels);
  m->F = F;
  m->raw_wf = add_raw(m, "__film_weight", static_cast<size_t>(F) * T);
  m->raw_bf = add_raw(m, "__film_bias", F);
  m->raws[m->raw_wf].ok = m->raws[m->raw_bf].ok = true;  // filled piecewise via film_rows
  m->cin0_pad = m->convs[m->conv_by_name["in_conv"]].cin_pad;
  return 0;
}

// ------------------------------------------------------------------------------ workspace planner
struct Planner {
  r2dm_model* m;
  uint8_t* base;       // nullptr for the dry run
  size_t top = 0;
  int B;
  struct Buf { size_t off, bytes; int refs; };
  std::vector<Buf> bufs;
  std::multimap<size_t, size_t> free_list;  // bytes -> offset
  std::vector<PT> tensors;
  std::vector<int> tensor_buf;

  size_t alloc(size_t bytes) {
    bytes = align_up(bytes, 1024);
    if (!m->keep_all) {
      auto it = free_list.find(bytes);
      if (it != free_list.end()) {
        size_t off = it->second;
        free_list.erase(it);
        return off;
      }
    }
    size_t off = top;
    top += bytes;
    return off;
  }
  int new_tensor(int C, int H, int W, int slots) {
    PT t;
    t.B = B; t.C = C; t.H = H; t.W = W; t.slots = slots;
    const size_t data = align_up(t.bytes(m->dtype), 256);
    const size_t st = slots > 0 ? static_cast<size_t>(B) * kNU * slots * 2 * sizeof(float) : 0;
    const size_t bytes = align_up(data + st, 1024);
    const size_t off = alloc(bytes);
    bufs.push_back({off, bytes, 1});
    t.ptr = base ? base + off : reinterpret_cast<void*>(off + 4096);  // dry run: fake non-null
    t.stats = slots > 0 ? reinterpret_cast<float*>(static_cast<uint8_t*>(t.ptr) + data) : nullptr;
    tensors.push_back(t);
    tensor_buf.push_back(static_cast<int>(bufs.size()) - 1);
    return static_cast<int>(tensors.size()) - 1;
  }
  void retain(int id) { bufs[tensor_buf[id]].refs++; }
  void release(int id) {
    Buf& b = bufs[tensor_buf[id]];
    if (--b.refs == 0) free_list.insert({b.bytes, b.off});
  }
};

struct Builder {
  r2dm_model* m;
  Planner pl;
  std::vector<Op> prog;
  std::map<std::string, PT> named;
  int n_convs = 0;

  const PT& T(int id) const { return pl.tensors[id]; }

  // xf: 0 none, 1 GroupNorm(affine from raw index gamma_raw)+SiLU, 2 AdaGN(film_name)+SiLU,
  //     3 GroupNorm(affine) without SiLU (attention)
  int conv(const std::string& wname, int in0, int in1, int residual, float scale, bool want_stats,
           bool is_output = false, int xf = 0, int gamma_raw = -1, const std::string& film_name = "",
           const std::string& skip_w = "", int sk0 = -1, int sk1 = -1) {
    const int wi = m->conv_by_name.at(wname);
    const ConvW& w = m->convs[wi];
    const PT& a = T(in0);
    Op op;
    op.kind = Op::CONV;
    op.conv_w = wi;
    op.is_output = is_output;
    ConvLaunch& l = op.conv;
    memset(&l, 0, sizeof(l));
    l.dtype = m->dtype; l.taps = w.taps; l.nt = w.nt;
    if (w.taps == 9) l.ht = (w.nt == 128) ? (a.H >= 2 ? 2 : 1) : (a.H % 4 == 0 ? 4 : (a.H >= 2 ? 2 : 1));
    else l.ht = a.H >= 2 ? 2 : 1;
    if (w.nt == 16 && l.ht != 4) l.ht = 4;
    // low-resolution levels: with two-row tiles fewer than half of the 148 SMs would get a tile, so
    // halve the tile (the persistent grid then covers twice as many SMs with half the K loop each)
    if (w.taps == 9 && w.nt == 128 && l.ht == 2) {
      // (decided per image, never from the batch size: the accumulation order depends on the tile
      //  shape and results must not depend on the batch composition)
      const int tiles_per_image = (a.H / 2) * (a.W / 128) * (w.cout_pad / 128);
      if (tiles_per_image <= get_option("ht1_max_tiles", 9)) l.ht = 1;
    }
    l.in0 = a;
    if (in1 >= 0) l.in1 = T(in1);
    l.cin_pad = w.cin_pad; l.cout = w.cout; l.cout_pad = w.cout_pad;
    l.scale = scale;
    // consecutive convolutions walk their tiles in opposite directions (L2 reuse of the previous output)
    l.reverse = get_option("serpentine", 0) ? (n_convs & 1) : 0;
    ++n_convs;
    int out_id = -1;
    if (!is_output) {
      int slots = 0;
      if (want_stats) slots = ((a.H + l.ht - 1) / l.ht) * (a.W / 128);
      out_id = pl.new_tensor(w.cout_pad, a.H, a.W, slots);
      l.out = T(out_id);
    } else {
      l.out = a;  // geometry only
      l.out.C = w.cout_pad;
      l.out.stats = nullptr; l.out.slots = 0;
    }
    l.residual = residual >= 0 ? T(residual).ptr : nullptr;
    if (!skip_w.empty()) {   // folded skip projection: extra K stages over the raw block input
      op.skip_w = m->conv_by_name.at(skip_w);
      l.sk0 = T(sk0);
      if (sk1 >= 0) l.sk1 = T(sk1);
      l.cin2_pad = m->convs[op.skip_w].cin_pad;
    }
    if (xf != 0) {
      l.xf.enabled = 1;
      l.xf.silu = xf != 3;
      l.xf.groups = m->cfg.gn_num_groups;
      l.xf.eps = m->cfg.gn_eps;
      if (xf == 2) {
        op.xf_film = true;
        l.xf.film_off = m->film_rows.at(film_name).first;
        l.xf.film_stride = m->F;
      } else {
        op.gn_gamma = gamma_raw;
      }
    }
    prog.push_back(op);
    return out_id;
  }
  int gn(int src0, int src1, int gamma_raw, const std::string& film_name, bool silu) {
    const PT& a = T(src0);
    const int Ctot = a.C + (src1 >= 0 ? T(src1).C : 0);
    const int out = pl.new_tensor(Ctot, a.H, a.W, 0);
    Op op;
    op.kind = Op::GN;
    GnApply& g = op.gn;
    memset(&g, 0, sizeof(g));
    g.dtype = m->dtype;
    g.src0 = T(src0);
    if (src1 >= 0) g.src1 = T(src1);
    g.dst = T(out);
    g.groups = m->cfg.gn_num_groups; g.eps = m->cfg.gn_eps;
    g.silu = silu ? 1 : 0;
    op.gn_gamma = gamma_raw;
    if (gamma_raw < 0) {
      g.film_off = m->film_rows.at(film_name).first;
      g.film_stride = m->F;
    }
    prog.push_back(op);
    return out;
  }
  int resample(int src, int dir) {
    const PT& a = T(src);
    Op op;
    int out;
    if (dir < 0) {
      PT tmp; tmp.B = a.B; tmp.C = a.C; tmp.H = a.H / 2; tmp.W = a.W / 2;
      out = pl.new_tensor(a.C, a.H / 2, a.W / 2, down2_stat_slots(m->dtype, tmp));
      op.kind = Op::DOWN;
    } else {
      out = pl.new_tensor(a.C, a.H * 2, a.W * 2, 0);
      op.kind = Op::UP;
    }
    op.a = T(src); op.b = T(out);
    prog.push_back(op);
    return out;
  }

  int build() {
    const r2dm_config& c = m->cfg;
    const float rs = c.residual_scale;
    // input staging tensor: [x | coords encoding | 0]
    const int xin = pl.new_tensor(m->cin0_pad, c.height, c.width, 0);
    named["__input"] = T(xin);
    {
      Op op; op.kind = Op::PACK_INPUT; op.a = T(xin);
      prog.push_back(op);
    }
    int h = conv("in_conv", xin, -1, -1, 1.f, true);
    named["in_conv"] = T(h);
    std::vector<int> skips;
    for (const BlockSpec& b : m->blocks) {
      int h2 = -1;  // second half of a concat input
      if (b.name[0] == 'u' && b.name != "u_block4") { h2 = skips.back(); skips.pop_back(); }
      if (b.down > 1) {
        const int t = conv(b.name + ".downsample.0", h, -1, -1, 1.f, false);
        pl.release(h);
        h = resample(t, -2);
        pl.release(t);
      }
      for (int i = 0; i < b.nres; ++i) {
        const std::string p = b.name + ".residual_blocks." + std::to_string(i);
        const int x0 = h, x1 = (i == 0) ? h2 : -1;
        // GroupNorm+SiLU and AdaGN+SiLU run inside the consumer convolutions (operand transform)
        // experiment (option unfuse_hw, off): below this many pixels per image apply GroupNorm / AdaGN + SiLU ONCE in
        // a stand-alone pass instead of once per N tile and halo row inside the conv (12x redundant at 8 x 128)
        const bool unfuse = T(x0).H * T(x0).W <= get_option("unfuse_hw", 0);
        int h1;
        if (unfuse) {
          const int n1 = gn(x0, x1, m->raw_by_name.at(p + ".norm1.weight"), "", true);
          h1 = conv(p + ".conv1", n1, -1, -1, 1.f, true);
          pl.release(n1);
        } else {
          h1 = conv(p + ".conv1", x0, x1, -1, 1.f, true, false, 1, m->raw_by_name.at(p + ".norm1.weight"));
        }
        int res = x0, sk = -1;
        std::string fold;
        if (m->conv_by_name.count(p + ".skip")) {
          if (m->convs[m->conv_by_name.at(p + ".skip")].sk_planes > 0) {
            fold = p + ".skip";      // runs inside conv2 (no launch, no skip tensor, no residual read)
            res = -1;
          } else {
            sk = conv(p + ".skip", x0, x1, -1, 1.f, false);
            res = sk;
          }
        }
        int o;
        if (unfuse) {
          const int n2 = gn(h1, -1, -1, p + ".norm2.proj.1", true);
          o = conv(p + ".conv2", n2, -1, res, rs, true, false, 0, -1, "", fold, x0, x1);
          pl.release(n2);
        } else {
          o = conv(p + ".conv2", h1, -1, res, rs, true, false, 2, -1, p + ".norm2.proj.1", fold, x0, x1);
        }
        pl.release(h1);
        if (sk >= 0) pl.release(sk);
        pl.release(x0);
        if (x1 >= 0) pl.release(x1);
        h = o;
        named[b.name + ".rb" + std::to_string(i)] = T(h);
      }
      if (b.attn) {
        const std::string p = b.name + ".self_attn_block";
        const int qkv = conv(p + ".attn.in_proj", h, -1, -1, 1.f, false, false, 3, m->raw_by_name.at(p + ".norm.weight"));
        prog.back().conv.round_out = 1;   // q, k, v are only ever read as tensor-core operands (fp32 engine: tf32)
        const int att = pl.new_tensor(b.cout, T(h).H, T(h).W, 0);
        {
          Op op; op.kind = Op::ATTN; op.a = T(qkv); op.b = T(att); op.heads = c.attn_num_heads;
          prog.push_back(op);
        }
        pl.release(qkv);
        const int o = conv(p + ".attn.out_proj", att, -1, h, rs, true);
        pl.release(att);
        pl.release(h);
        h = o;
      }
      if (b.up > 1) {
        const int t = resample(h, +2);
        pl.release(h);
        h = conv(b.name + ".upsample.1", t, -1, -1, 1.f, true);
        pl.release(t);
      }
      named[b.name] = T(h);
      if (b.name == "d_block1" || b.name == "d_block2" || b.name == "d_block3") {
        skips.push_back(h);
        pl.retain(h);
      }
    }
    conv("out_conv", h, -1, -1, 1.f, false, true);
    pl.release(h);
    return 0;
  }
};

constexpr size_t kChainCounterBytes = 64 * 1024;   // tile counters of the conv chains, after the planned buffers

int validate_config(const r2dm_config& c) {
  if (c.gn_num_groups != kNU) return fail(-1, "gn_num_groups must be %d", kNU);
  if (c.base_channels % 64 != 0) return fail(-1, "base_channels must be a multiple of 64");
  if (c.width % 1024 != 0) return fail(-1, "width must be a multiple of 1024 (bottleneck rows of >=128 px)");
  if (c.height % 8 != 0 || c.height < 8) return fail(-1, "height must be a multiple of 8");
  if (c.dtype != R2DM_F32 && c.dtype != R2DM_BF16) return fail(-1, "dtype must be R2DM_F32 or R2DM_BF16");
  for (int i = 0; i < 4; ++i) {
    if (c.channel_multiplier[i] < 1 || c.num_residual_blocks[i] < 1) return fail(-1, "bad multiplier / block count");
  }
  const int E4 = c.base_channels * c.channel_multiplier[3], E3 = c.base_channels * c.channel_multiplier[2];
  for (int E : {E4, E3}) {
    if (E % c.attn_num_heads) return fail(-1, "attention width %d not divisible by heads", E);
    const int hd = E / c.attn_num_heads;
    if (hd != 32 && hd != 64) return fail(-1, "head dim %d unsupported (32 or 64)", hd);
  }
  if (c.in_channels < 1 || c.in_channels > 16) return fail(-1, "in_channels out of range");
  // widest convolution input: the bottleneck (C4) or a concatenated up-block input (2 * C[k]); the fused
  // GroupNorm coefficient table and the per-thread channel loop of the transform are sized for kMaxConvCin
  int C[5] = {c.base_channels, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) C[i + 1] = c.base_channels * c.channel_multiplier[i];
  int widest = C[4];
  for (int k = 1; k <= 3; ++k) widest = std::max(widest, 2 * C[k]);
  if (widest > kMaxConvCin)
    return fail(-1, "convolution input of %d channels exceeds the supported maximum of %d (base_channels x multipliers too large)",
                widest, kMaxConvCin);
  // up-blocks 3..1 read the concat [h, skip_k] (2 * C[k] channels) and produce C[k-1]; when the two are equal
  // the reference's ResidualBlock uses an identity skip over the CONCATENATED tensor (efficient_unet.py:87-91),
  // which this launch program (residual = first concat half) does not implement
  for (int k = 1; k <= 3; ++k)
    if (2 * C[k] == C[k - 1])
      return fail(-1, "channel_multiplier makes up-block %d's concat input as wide as its output (%d): identity skip "
                      "over a concat is not supported", k, C[k - 1]);
  const int T = c.temb_channels > 0 ? c.temb_channels : 4 * c.base_channels;
  if (static_cast<size_t>(8) * T * sizeof(float) > 48 * 1024 || static_cast<size_t>(c.base_channels + T) * sizeof(float) > 48 * 1024)
    return fail(-1, "temb_channels = %d needs more than 48 KB of shared memory in the conditioning kernels (max 1536)", T);
  return 0;
}

}  // namespace

// ================================================================================== C ABI
extern "C" {

const char* r2dm_last_error(void) { return g_err; }
int r2dm_version(void) { return 100; }

int r2dm_create(const r2dm_config* cfg, r2dm_handle* out) {
  if (!cfg || !out) return fail(-1, "null argument");
  int rc = validate_config(*cfg);
  if (rc) return rc;
  r2dm_model* m = new r2dm_model();
  m->cfg = *cfg;
  if (m->cfg.temb_channels <= 0) m->cfg.temb_channels = 4 * cfg->base_channels;
  if (m->cfg.residual_scale <= 0.f) m->cfg.residual_scale = 0.70710678118654752f;
  m->dtype = cfg->dtype;
  m->cw = dtype_cw(m->dtype);
  m->T = m->cfg.temb_channels;
  plan_weights(m);
  const char* dbg = getenv("R2DM_KEEP_ACTIVATIONS");
  m->keep_all = dbg && dbg[0] == '1';
  *out = m;
  return 0;
}

int r2dm_destroy(r2dm_handle h) {
  delete h;
  return 0;
}

size_t r2dm_weight_arena_bytes(r2dm_handle h) { return h ? h->arena_bytes : 0; }

int r2dm_bind_weight_arena(r2dm_handle h, void* arena, size_t bytes) {
  if (!h || !arena) return fail(-1, "null argument");
  if (bytes < h->arena_bytes) return fail(-1, "weight arena too small: %zu < %zu", bytes, h->arena_bytes);
  if (reinterpret_cast<uintptr_t>(arena) % 256) return fail(-1, "weight arena must be 256-byte aligned");
  h->arena = static_cast<uint8_t*>(arena);
  return 0;
}

static size_t numel_of(const int64_t* shape, int ndim) {
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= static_cast<size_t>(shape[i]);
  return n;
}

int r2dm_load_tensor(r2dm_handle h, const char* name_c, const float* src, const int64_t* shape, int ndim,
                     void* stream) {
  if (!h || !name_c || !src) return fail(-1, "null argument");
  if (!h->arena) return fail(-1, "bind the weight arena first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::string name(name_c);
  const size_t n = numel_of(shape, ndim);
  // attention projections are stored as 1x1 convolutions
  auto ends_with = [&](const char* suf) {
    const size_t l = strlen(suf);
    return name.size() >= l && name.compare(name.size() - l, l, suf) == 0;
  };
  std::string cname;
  bool is_w = false, is_b = false;
  if (ends_with(".attn.in_proj_weight")) { cname = name.substr(0, name.size() - 7); is_w = true; }
  else if (ends_with(".attn.in_proj_bias")) { cname = name.substr(0, name.size() - 5); is_b = true; }
  else if (ends_with(".weight")) { cname = name.substr(0, name.size() - 7); is_w = true; }
  else if (ends_with(".bias")) { cname = name.substr(0, name.size() - 5); is_b = true; }
  auto ci = h->conv_by_name.find(cname);
  if ((is_w || is_b) && ci != h->conv_by_name.end()) {
    ConvW& c = h->convs[ci->second];
    if (is_w) {
      if (n != static_cast<size_t>(c.cout) * c.cin * c.taps)
        return fail(-3, "%s: expected %d x %d x %d elements, got %zu", name_c, c.cout, c.cin, c.taps, n);
      CUDA_TRY(pack_conv_weight(h->dtype, c.taps, c.nt, src, c.cout, c.cin, c.cin_pad, c.cout_pad,
                                h->arena + c.w_off, s, c.sk_planes));
      c.w_ok = true;
    } else {
      if (n != static_cast<size_t>(c.cout)) return fail(-3, "%s: expected %d elements, got %zu", name_c, c.cout, n);
      CUDA_TRY(cudaMemsetAsync(h->arena + c.b_off, 0, static_cast<size_t>(c.cout_pad) * sizeof(float), s));
      CUDA_TRY(cudaMemcpyAsync(h->arena + c.b_off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
      c.b_ok = true;
    }
    return 0;
  }
  auto fi = h->film_rows.find(cname);
  if ((is_w || is_b) && fi != h->film_rows.end()) {
    const int row0 = fi->second.first, rows = fi->second.second;
    if (is_w) {
      if (n != static_cast<size_t>(rows) * h->T) return fail(-3, "%s: bad shape", name_c);
      CUDA_TRY(cudaMemcpyAsync(h->raw_ptr(h->raw_wf) + static_cast<size_t>(row0) * h->T, src, n * sizeof(float),
                               cudaMemcpyDeviceToDevice, s));
    } else {
      if (n != static_cast<size_t>(rows)) return fail(-3, "%s: bad shape", name_c);
      CUDA_TRY(cudaMemcpyAsync(h->raw_ptr(h->raw_bf) + row0, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    // bookkeeping: mark via a pseudo raw entry
    h->raw_by_name.emplace(name, -1);
    return 0;
  }
  auto ri = h->raw_by_name.find(name);
  if (ri != h->raw_by_name.end() && ri->second >= 0) {
    RawW& r = h->raws[ri->second];
    if (n != r.numel) return fail(-3, "%s: expected %zu elements, got %zu", name_c, r.numel, n);
    CUDA_TRY(cudaMemcpyAsync(h->arena + r.off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    r.ok = true;
    return 0;
  }
  return 1;  // not a tensor this path consumes (buffers such as coords, scale, kernel, _dummy)
}

int r2dm_missing_tensors(r2dm_handle h, char* names_out, size_t cap) {
  if (!h) return fail(-1, "null argument");
  std::string acc;
  int n = 0;
  auto add = [&](const std::string& s) { ++n; acc += s; acc += '\n'; };
  for (const ConvW& c : h->convs) {
    if (!c.w_ok) add(c.name + ".weight");
    if (!c.b_ok) add(c.name + ".bias");
  }
  for (const RawW& r : h->raws)
    if (!r.ok) add(r.name);
  for (const auto& kv : h->film_rows) {
    if (!h->raw_by_name.count(kv.first + ".weight")) add(kv.first + ".weight");
    if (!h->raw_by_name.count(kv.first + ".bias")) add(kv.first + ".bias");
  }
  if (names_out && cap > 0) {
    strncpy(names_out, acc.c_str(), cap - 1);
    names_out[cap - 1] = 0;
  }
  return n;
}

size_t r2dm_workspace_bytes(r2dm_handle h, int batch) {
  if (!h || batch < 1) return 0;
  Builder b;
  b.m = h;
  b.pl.m = h; b.pl.base = nullptr; b.pl.B = batch;
  b.build();
  return b.pl.top + 4096 + kChainCounterBytes;
}

int r2dm_bind_workspace(r2dm_handle h, void* workspace, size_t bytes, int batch, void* stream) {
  if (!h || !workspace || batch < 1) return fail(-1, "bad argument");
  if (!h->arena) return fail(-1, "bind the weight arena first");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(-1, "workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Builder b;
  b.m = h;
  b.pl.m = h; b.pl.base = static_cast<uint8_t*>(workspace); b.pl.B = batch;
  b.build();
  if (b.pl.top > bytes) return fail(-1, "workspace too small: %zu < %zu", bytes, b.pl.top);
  h->ws = static_cast<uint8_t*>(workspace);
  h->ws_bytes = bytes;
  h->batch = batch;
  h->prog = b.prog;
  h->named = b.named;
  // resolve weights, build tensor maps
  int launches = 0;
  for (Op& op : h->prog) {
    if (op.kind == Op::CONV) {
      const ConvW& w = h->convs[op.conv_w];
      op.conv.wpacked = h->arena + w.w_off;
      op.conv.bias = reinterpret_cast<const float*>(h->arena + w.b_off);
      if (op.skip_w >= 0) {
        const ConvW& sw = h->convs[op.skip_w];
        op.conv.w2packed = h->arena + sw.w_off;
        op.conv.bias2 = reinterpret_cast<const float*>(h->arena + sw.b_off);
      }
      int rc = conv_make_tmaps(op.conv);
      if (rc) return fail(-4, "cuTensorMapEncodeTiled failed for %s (%d)", w.name.c_str(), rc);
      if (op.conv.xf.enabled && op.gn_gamma >= 0) {
        op.conv.xf.gamma = h->raw_ptr(op.gn_gamma);
        op.conv.xf.beta = h->raw_ptr(op.gn_gamma + 1);
      }
    } else if (op.kind == Op::ATTN) {
      op.attn_exact = h->dtype != kBF16 && get_option("attn_exact", 0) != 0;
      int rc = op.attn_exact ? 0 : attention_make_tmaps(&op.attn_tmap, &op.attn_tmap_kv, h->dtype, op.a, op.heads);
      if (rc) return fail(-4, "cuTensorMapEncodeTiled failed for attention (%d)", rc);
    } else if (op.kind == Op::GN && op.gn_gamma >= 0) {
      op.gn.gamma = h->raw_ptr(op.gn_gamma);
      op.gn.beta = h->raw_ptr(op.gn_gamma + 1);
    }
    ++launches;
  }
  // Chains (option chain = 1, off by default): maximal runs of consecutive 3x3 convolutions with the same tile
  // geometry (the ResidualBlock convs of one resolution level) become ONE persistent launch each (conv_chain.cu).
  // Bit-identical results; measured +-0 on the forward, so the product keeps one launch per layer.
  if (get_option("chain", 0)) {
    size_t ctr_off = align_up(b.pl.top, 256);
    const int n_ops = static_cast<int>(h->prog.size());
    int n_chains = 0;
    for (int i = 0; i < n_ops;) {
      Op& first = h->prog[i];
      int len = 0;
      if (first.kind == Op::CONV && conv_chain_supported(first.conv)) {
        len = 1;
        while (i + len < n_ops && len < kMaxChainLayers) {
          const Op& o = h->prog[i + len];
          if (o.kind != Op::CONV || !conv_chain_supported(o.conv) || o.conv.nt != first.conv.nt ||
              o.conv.ht != first.conv.ht || o.conv.out.H != first.conv.out.H || o.conv.out.W != first.conv.out.W ||
              o.conv.cout_pad != first.conv.cout_pad || o.conv.cout != first.conv.cout)
            break;
          ++len;
        }
      }
      if (len >= 2) {
        const size_t need = static_cast<size_t>(len) * batch * sizeof(int);
        const bool enabled = (get_option("chain_mask", -1) >> n_chains++) & 1;   // developer: pick chains by order
        if (enabled && ctr_off + need <= bytes) {
          first.chain_len = len;
          first.chain_done = reinterpret_cast<int*>(h->ws + ctr_off);
          ctr_off += align_up(need, 256);
          launches -= len - 1;
        }
        i += len;
      } else {
        ++i;
      }
    }
  }
  h->n_launches = launches;
  // constant planes of the input staging tensor (coordinate encoding + zero padding)
  const PT& xin = h->named.at("__input");
  CUDA_TRY(cudaMemsetAsync(xin.ptr, 0, xin.bytes(h->dtype), s));
  // all planes once (x = null -> image channels read as 0); forwards rewrite only the image planes
  CUDA_TRY(pack_input(h->dtype, nullptr, h->cfg.in_channels, h->raw_enc >= 0 ? h->raw_ptr(h->raw_enc) : nullptr,
                      h->cfg.extra_channels, xin, 0, xin.C / h->cw, s));
  return 0;
}

int r2dm_film_width(r2dm_handle h) { return h ? h->F : 0; }
int r2dm_num_launches(r2dm_handle h) { return h ? h->n_launches : 0; }

int r2dm_cond_embed(r2dm_handle h, const float* cond, int rows, float* scratch, float* film, void* stream) {
  if (!h || !cond || !scratch || !film || rows < 1) return fail(-1, "bad argument");
  if (!h->arena) return fail(-1, "bind the weight arena first");
  CondEmbed c;
  c.cond = cond; c.rows = rows;
  c.base_ch = h->cfg.base_channels; c.temb_ch = h->T;
  c.w1 = h->raw_ptr(h->raw_w1); c.b1 = h->raw_ptr(h->raw_b1);
  c.w2 = h->raw_ptr(h->raw_w2); c.b2 = h->raw_ptr(h->raw_b2);
  c.wf = h->raw_ptr(h->raw_wf); c.bf = h->raw_ptr(h->raw_bf);
  c.F = h->F;
  c.temb_scratch = scratch; c.film = film;
  CUDA_TRY(cond_embed_launch(c, static_cast<cudaStream_t>(stream)));
  return 0;
}

static int launch_op(r2dm_handle h, Op& op, const float* x, const float* film, const int* step_ptr,
                     int rows_per_step, int row_batch_stride, float* pred, cudaStream_t s) {
  const r2dm_config& c = h->cfg;
  switch (op.kind) {
    case Op::PACK_INPUT: {
      // rewrite only the planes that contain image channels; the rest is constant
      const int pe = (c.in_channels + h->cw - 1) / h->cw;
      CUDA_TRY(pack_input(h->dtype, x, c.in_channels, h->raw_enc >= 0 ? h->raw_ptr(h->raw_enc) : nullptr,
                          c.extra_channels, op.a, 0, pe, s));
      break;
    }
    case Op::CONV: {
      if (op.is_output) op.conv.out_nchw = pred;
      if (op.xf_film) {
        op.conv.xf.film = film; op.conv.xf.step_ptr = step_ptr;
        op.conv.xf.rows_per_step = rows_per_step; op.conv.xf.row_batch_stride = row_batch_stride;
      }
      CUDA_TRY(conv_launch(op.conv, s));
      break;
    }
    case Op::GN: {
      if (op.gn_gamma < 0) {
        op.gn.film = film; op.gn.step_ptr = step_ptr;
        op.gn.rows_per_step = rows_per_step; op.gn.row_batch_stride = row_batch_stride;
      }
      CUDA_TRY(gn_apply_launch(op.gn, s));
      break;
    }
    case Op::DOWN: CUDA_TRY(down2_launch(h->dtype, op.a, op.b, s)); break;
    case Op::UP: CUDA_TRY(up2_launch(h->dtype, op.a, op.b, s)); break;
    case Op::ATTN:
      if (!op.attn_exact)
        CUDA_TRY(attention_umma_launch(h->dtype, op.a, op.b, op.heads, op.attn_tmap, op.attn_tmap_kv, s));
      else CUDA_TRY(attention_launch(h->dtype, op.a, op.b, op.heads, s));
      break;
  }
  return 0;
}

// Launches op i - or, when it starts a chain, the whole chain - and reports how many program entries that covered.
static int launch_at(r2dm_handle h, int i, const float* x, const float* film, const int* step_ptr, int rows_per_step,
                     int row_batch_stride, float* pred, cudaStream_t s, int* consumed) {
  Op& op = h->prog[i];
  *consumed = 1;
  if (op.kind != Op::CONV || op.chain_len < 2)
    return launch_op(h, op, x, film, step_ptr, rows_per_step, row_batch_stride, pred, s);
  const ConvLaunch* ls[kMaxChainLayers];
  for (int k = 0; k < op.chain_len; ++k) {
    Op& o = h->prog[i + k];
    if (o.xf_film) {
      o.conv.xf.film = film; o.conv.xf.step_ptr = step_ptr;
      o.conv.xf.rows_per_step = rows_per_step; o.conv.xf.row_batch_stride = row_batch_stride;
    }
    ls[k] = &o.conv;
  }
  // two image groups keep every CTA busy across layer boundaries; a single image has nothing to interleave with
  const int B = op.conv.out.B;
  const int split = (B >= 2 && !get_option("chain_nosplit", 0)) ? (B + 1) / 2 : B;
  CUDA_TRY(conv_chain_launch(ls, op.chain_len, op.chain_done, split, s));
  *consumed = op.chain_len;
  return 0;
}

int r2dm_unet_forward(r2dm_handle h, const float* x, const float* film, const int* step_ptr, int rows_per_step,
                      int row_batch_stride, float* pred, void* stream) {
  if (!h || !x || !film || !pred) return fail(-1, "null argument");
  if (!h->ws) return fail(-1, "bind a workspace first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(h->prog.size());
  for (int i = 0; i < n;) {
    int used = 1;
    int rc = launch_at(h, i, x, film, step_ptr, rows_per_step, row_batch_stride, pred, s, &used);
    if (rc) return rc;
    i += used;
  }
  return 0;
}

// Measurement aid (bench.py): enqueue only the launches of one forward whose kind bit is set in kind_mask
// (bit k = kind k of r2dm_profile_forward), in program order, on the buffers of the last real forward.  Lets
// bench.py time e.g. the 56 3x3 convolutions back to back inside a CUDA graph exactly as the product
// launches them (programmatic dependent launch, no host gaps).  Results are meaningless; never on the product path.
int r2dm_debug_forward_kinds(r2dm_handle h, const float* x, const float* film, float* pred, unsigned kind_mask,
                             void* stream) {
  if (!h || !x || !film || !pred) return fail(-1, "null argument");
  if (!h->ws) return fail(-1, "bind a workspace first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(h->prog.size());
  for (int i = 0; i < n;) {
    Op& op = h->prog[i];
    int k = 0;
    switch (op.kind) {
      case Op::PACK_INPUT: k = 0; break;
      case Op::CONV: k = h->convs[op.conv_w].taps == 9 ? 1 : 2; break;
      case Op::GN: k = 3; break;
      case Op::DOWN: k = 4; break;
      case Op::UP: k = 5; break;
      case Op::ATTN: k = 6; break;
    }
    int used = (op.kind == Op::CONV && op.chain_len >= 2) ? op.chain_len : 1;
    if (kind_mask >> k & 1u) {
      int rc = launch_at(h, i, x, film, nullptr, 0, 1, pred, s, &used);
      if (rc) return rc;
    }
    i += used;
  }
  return 0;
}

// Measurement aid (bench.py / profiles): one eager forward with a CUDA event pair around every
// launch.  Synchronises; never used on the product path.  kind: 0 pack_input, 1 conv3x3, 2 conv1x1,
// 3 GroupNorm/AdaGN apply, 4 down2, 5 up2, 6 attention.  flops / bytes are ALGORITHMIC (unpadded).
int r2dm_profile_forward(r2dm_handle h, const float* x, const float* film, float* pred, void* stream,
                         int cap, int* kind, float* ms, double* flops, double* bytes) {
  if (!h || !x || !film || !pred) return fail(-1, "null argument");
  if (!h->ws) return fail(-1, "bind a workspace first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(h->prog.size());
  if (cap < n) return fail(-1, "capacity %d < %d ops", cap, n);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
  const double es = dtype_size(h->dtype);
  int covered = 0;   // program entries already executed by a chain launch (their time is booked on its first entry)
  for (int i = 0; i < n; ++i) {
    Op& op = h->prog[i];
    CUDA_TRY(cudaEventRecord(ev[i], s));
    if (covered > 0) {
      --covered;
    } else {
      int used = 1;
      int rc = launch_at(h, i, x, film, nullptr, 0, 1, pred, s, &used);
      if (rc) return rc;
      covered = used - 1;
    }
    double fl = 0, by = 0;
    int k = 0;
    auto tb = [&](const PT& t) { return static_cast<double>(t.B) * t.C * t.H * t.W * es; };
    switch (op.kind) {
      case Op::PACK_INPUT: k = 0; by = static_cast<double>(op.a.B) * h->cfg.in_channels * op.a.H * op.a.W * 4 +
                                      static_cast<double>(op.a.B) * h->cw * op.a.H * op.a.W * es; break;
      case Op::CONV: {
        const ConvW& w = h->convs[op.conv_w];
        const PT& o = op.conv.out;
        k = w.taps == 9 ? 1 : 2;
        fl = 2.0 * o.B * o.H * o.W * w.cin * w.cout * w.taps;
        if (op.skip_w >= 0) fl += 2.0 * o.B * o.H * o.W * h->convs[op.skip_w].cin * w.cout;   // folded 1x1 skip
        by = static_cast<double>(o.B) * o.H * o.W * (w.cin * es + w.cout * (op.is_output ? 4.0 : es)) +
             static_cast<double>(w.cin) * w.cout * w.taps * es + (op.conv.residual ? static_cast<double>(o.B) * o.H * o.W * w.cout * es : 0.0);
        break;
      }
      case Op::GN: k = 3; by = 2.0 * tb(op.gn.dst); fl = 8.0 * op.gn.dst.B * op.gn.dst.C * op.gn.dst.H * op.gn.dst.W; break;
      case Op::DOWN: k = 4; by = tb(op.a) + tb(op.b); break;
      case Op::UP: k = 5; by = tb(op.a) + tb(op.b); break;
      case Op::ATTN: {
        k = 6;
        const double Lt = static_cast<double>(op.b.H) * op.b.W;
        fl = 4.0 * op.b.B * Lt * Lt * op.b.C;
        by = tb(op.a) + tb(op.b);
        break;
      }
    }
    kind[i] = k; flops[i] = fl; bytes[i] = by;
  }
  CUDA_TRY(cudaEventRecord(ev[n], s));
  CUDA_TRY(cudaEventSynchronize(ev[n]));
  for (int i = 0; i < n; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return n;
}

int r2dm_sampler_update(float* x_out, const float* x, const float* pred, const float* noise, const float* coef,
                        int coef_cols, const int* step_ptr, int rows_per_step, int row_batch_stride, float clip,
                        const float* known, const float* mask, const float* noise2, int batch, size_t per_sample,
                        void* stream) {
  if (!x_out || !x || !pred || !noise || !coef) return fail(-1, "null argument");
  if (coef_cols < (known ? 7 : 5)) return fail(-1, "coef_cols too small");
  if (known && (!mask || !noise2)) return fail(-1, "RePaint blend needs mask and noise2");
  SamplerUpdate u;
  u.x = const_cast<float*>(x); u.pred = pred; u.noise = noise; u.coef = coef; u.coef_cols = coef_cols;
  u.step_ptr = step_ptr; u.rows_per_step = rows_per_step; u.row_batch_stride = row_batch_stride;
  u.clip = clip; u.known = known; u.mask = mask; u.noise2 = noise2; u.x_out = x_out;
  u.B = batch; u.per_sample = per_sample;
  memset(&u.ph, 0, sizeof(u.ph)); u.draw_noise = u.draw_noise2 = 0;
  CUDA_TRY(sampler_update_launch(u, static_cast<cudaStream_t>(stream)));
  return 0;
}

static int to_philox(const r2dm_philox* p, PhiloxDraw* out) {
  if (!p || !p->seeds || !p->offsets) return fail(-1, "philox: null seeds / offsets");
  if (p->threads == 0 || p->offset_per_draw == 0) return fail(-1, "philox: threads / offset_per_draw must be > 0");
  out->seeds = reinterpret_cast<const unsigned long long*>(p->seeds);
  out->offsets = reinterpret_cast<const unsigned long long*>(p->offsets);
  out->ctr0 = p->ctr0; out->ctr1 = p->ctr1; out->mul0 = p->mul0; out->mul1 = p->mul1;
  out->offset_per_draw = p->offset_per_draw; out->threads = p->threads;
  return 0;
}

int r2dm_sampler_update_philox(float* x_out, const float* x, const float* pred, const float* coef, int coef_cols,
                               const int* step_ptr, int rows_per_step, int row_batch_stride, float clip,
                               const float* known, const float* mask, const r2dm_philox* philox, int draw_noise,
                               int draw_noise2, int batch, size_t per_sample, void* stream) {
  if (!x_out || !x || !pred || !coef) return fail(-1, "null argument");
  if (coef_cols < (known ? 7 : 5)) return fail(-1, "coef_cols too small");
  if (known && !mask) return fail(-1, "RePaint blend needs a mask");
  SamplerUpdate u;
  u.x = const_cast<float*>(x); u.pred = pred; u.noise = nullptr; u.coef = coef; u.coef_cols = coef_cols;
  u.step_ptr = step_ptr; u.rows_per_step = rows_per_step; u.row_batch_stride = row_batch_stride;
  u.clip = clip; u.known = known; u.mask = mask; u.noise2 = nullptr; u.x_out = x_out;
  u.B = batch; u.per_sample = per_sample;
  int rc = to_philox(philox, &u.ph);
  if (rc) return rc;
  u.draw_noise = draw_noise; u.draw_noise2 = draw_noise2;
  CUDA_TRY(sampler_update_launch(u, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_axpby_table_philox(float* y, const float* x, const float* table, const int* step_ptr, int rows_per_step,
                            int row_batch_stride, const r2dm_philox* philox, int draw, int batch,
                            size_t per_sample, void* stream) {
  if (!y || !x || !table) return fail(-1, "null argument");
  PhiloxDraw ph;
  int rc = to_philox(philox, &ph);
  if (rc) return rc;
  CUDA_TRY(axpby_launch(x, nullptr, table, y, batch, per_sample, step_ptr, rows_per_step, row_batch_stride,
                        static_cast<cudaStream_t>(stream), &ph, draw));
  return 0;
}

int r2dm_philox_normal(float* out, const r2dm_philox* philox, int draw, int batch, size_t per_sample, void* stream) {
  if (!out) return fail(-1, "null argument");
  PhiloxDraw ph;
  int rc = to_philox(philox, &ph);
  if (rc) return rc;
  CUDA_TRY(philox_normal_launch(out, ph, draw, batch, per_sample, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_axpby(float* y, const float* x, const float* noise, const float* ac, int batch, size_t per_sample,
               void* stream) {
  if (!y || !x || !noise || !ac) return fail(-1, "null argument");
  CUDA_TRY(axpby_launch(x, noise, ac, y, batch, per_sample, nullptr, 0, 1, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_axpby_table(float* y, const float* x, const float* noise, const float* table, const int* step_ptr,
                     int rows_per_step, int row_batch_stride, int batch, size_t per_sample, void* stream) {
  if (!y || !x || !noise || !table) return fail(-1, "null argument");
  CUDA_TRY(axpby_launch(x, noise, table, y, batch, per_sample, step_ptr, rows_per_step, row_batch_stride,
                        static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_advance_step(int* step_ptr, int delta, void* stream) {
  if (!step_ptr) return fail(-1, "null argument");
  CUDA_TRY(advance_step_launch(step_ptr, delta, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_lidar_postprocess(const float* sample, const float* angles, float* out, int batch, int H, int W,
                           int depth_format, float min_depth, float max_depth, void* stream) {
  if (!sample || !angles || !out) return fail(-1, "null argument");
  if (depth_format < 0 || depth_format > 2) return fail(-1, "bad depth_format");
  CUDA_TRY(lidar_postprocess_launch(sample, angles, out, batch, H, W, depth_format, min_depth, max_depth,
                                    static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------- caller-side consumers
int r2dm_render_point_clouds(const float* points, const float* colors, const float* R, const float* t, float* acc,
                             float* out, int batch, int num_points, int size, float focal_length, void* stream) {
  if (!points || !acc || !out) return fail(-1, "null argument");
  if (batch < 1 || num_points < 1 || size < 2 || size > 4096) return fail(-1, "bad batch / num_points / size");
  CUDA_TRY(render_splat_launch(points, colors, R, t, acc, out, batch, num_points, size, focal_length,
                               static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_bilinear_rasterize(const float* coords, const float* values, float* out, int batch, int num_points,
                            int channels, int H, int W, void* stream) {
  if (!coords || !values || !out) return fail(-1, "null argument");
  if (batch < 1 || num_points < 1 || channels < 1 || H < 1 || W < 1 || static_cast<long long>(H) * W > (1 << 24))
    return fail(-1, "bad shape (H * W must stay below 2^24: the reference computes pixel indices in fp32)");
  CUDA_TRY(rasterize_launch(coords, values, out, batch, num_points, channels, H, W, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_surface_normal(const float* points, float* out, int batch, int H, int W, int d, int mode, void* stream) {
  if (!points || !out) return fail(-1, "null argument");
  if (batch < 1 || H < 1 || W < 1 || d < 1 || d > W) return fail(-1, "bad shape / neighbour distance");
  if (mode != 0 && mode != 1) return fail(-1, "mode must be 0 (closest) or 1 (mean)");
  CUDA_TRY(surface_normal_launch(points, out, batch, H, W, d, mode, static_cast<cudaStream_t>(stream)));
  return 0;
}

int r2dm_bev_histogram(const float* points, const float* edges, unsigned int* counts, float* hist, int batch,
                       int num_points, int bins, float min_depth, float max_depth, void* stream) {
  if (!points || !edges || !counts || !hist) return fail(-1, "null argument");
  if (batch < 1 || num_points < 1 || bins < 1) return fail(-1, "bad shape");
  CUDA_TRY(bev_histogram_launch(points, edges, counts, hist, batch, num_points, bins, min_depth, max_depth,
                                static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------- op hooks
namespace {
struct Scratch {
  uint8_t* base; size_t cap; size_t top = 0;
  void* take(size_t bytes) {
    bytes = align_up(bytes, 1024);
    if (top + bytes > cap) return nullptr;
    void* p = base + top;
    top += bytes;
    return p;
  }
};
PT make_pt(Scratch& sc, int dt, int B, int C, int H, int W, int slots) {
  PT t; t.B = B; t.C = C; t.H = H; t.W = W; t.slots = slots;
  t.ptr = sc.take(t.bytes(dt));
  t.stats = slots > 0 ? static_cast<float*>(sc.take(static_cast<size_t>(B) * kNU * slots * 2 * sizeof(float))) : nullptr;
  return t;
}
}  // namespace

size_t r2dm_op_scratch_bytes(int batch, int max_channels, int H, int W) {
  const size_t t = static_cast<size_t>(batch) * round_up(max_channels, 64) * H * (W + 2) * 4 + 8192;
  return 4 * t + (static_cast<size_t>(9) * round_up(max_channels, 64) * round_up(max_channels, 128) * 4) + (1 << 20);
}

int r2dm_op_conv(int dtype, int taps, const float* x, const float* w, const float* bias, const float* residual,
                 float scale, float* y, int B, int Cin, int Cout, int H, int W, void* scratch, size_t scratch_bytes,
                 void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (W % 128) return fail(-1, "W must be a multiple of 128");
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  ConvLaunch l;
  memset(&l, 0, sizeof(l));
  l.dtype = dtype; l.taps = taps;
  l.nt = pick_nt(Cout);
  l.cin_pad = round_up(Cin, conv_stage_channels(dtype, taps));
  l.cout = Cout; l.cout_pad = round_up(Cout, l.nt);
  if (taps == 9) l.ht = (l.nt == 128) ? (H >= 2 ? 2 : 1) : (H % 4 == 0 ? 4 : (H >= 2 ? 2 : 1));
  else l.ht = H >= 2 ? 2 : 1;
  if (l.nt == 16 && l.ht != 4) return fail(-1, "small-N conv needs H %% 4 == 0");
  l.in0 = make_pt(sc, dtype, B, l.cin_pad, H, W, 0);
  l.out = make_pt(sc, dtype, B, l.cout_pad, H, W, 0);
  PT res = make_pt(sc, dtype, B, l.cout_pad, H, W, 0);
  void* wp = sc.take(conv_packed_weight_bytes(dtype, taps, l.nt, l.cin_pad, l.cout_pad));
  float* bp = static_cast<float*>(sc.take(static_cast<size_t>(l.cout_pad) * 4));
  if (!l.in0.ptr || !l.out.ptr || !res.ptr || !wp || !bp) return fail(-1, "scratch too small");
  CUDA_TRY(cudaMemsetAsync(l.in0.ptr, 0, l.in0.bytes(dtype), s));
  CUDA_TRY(pack_nchw(dtype, x, B, Cin, H, W, l.in0, 0, s));
  CUDA_TRY(pack_conv_weight(dtype, taps, l.nt, w, Cout, Cin, l.cin_pad, l.cout_pad, wp, s));
  CUDA_TRY(cudaMemsetAsync(bp, 0, static_cast<size_t>(l.cout_pad) * 4, s));
  if (bias) CUDA_TRY(cudaMemcpyAsync(bp, bias, static_cast<size_t>(Cout) * 4, cudaMemcpyDeviceToDevice, s));
  if (residual) {
    CUDA_TRY(cudaMemsetAsync(res.ptr, 0, res.bytes(dtype), s));
    CUDA_TRY(pack_nchw(dtype, residual, B, Cout, H, W, res, 0, s));
    l.residual = res.ptr;
  }
  l.wpacked = wp; l.bias = bp; l.scale = scale;
  int rc = conv_make_tmaps(l);
  if (rc) return fail(-4, "tensor map encode failed (%d)", rc);
  CUDA_TRY(conv_launch(l, s));
  CUDA_TRY(unpack_nchw(dtype, l.out, y, 0, Cout, s));
  return 0;
}

// GroupNorm/AdaGN(+SiLU) fused into the following convolution, exactly as the network runs it:
// y = conv(silu(gn(x))) (+bias).  film: per-sample [B][2*Cin] = [scale || shift] or NULL (then gamma/beta).
int r2dm_op_gn_conv(int dtype, int taps, const float* x, const float* gamma, const float* beta, const float* film,
                    float eps, int silu, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                    int H, int W, void* scratch, size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (W % 128) return fail(-1, "W must be a multiple of 128");
  if (Cin % (kNU * dtype_cw(dtype))) return fail(-1, "Cin must be a multiple of %d", kNU * dtype_cw(dtype));
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  ConvLaunch l;
  memset(&l, 0, sizeof(l));
  l.dtype = dtype; l.taps = taps;
  l.nt = pick_nt(Cout);
  l.cin_pad = round_up(Cin, conv_stage_channels(dtype, taps));
  if (l.cin_pad != Cin) return fail(-1, "Cin must be a multiple of the stage K (%d)", conv_stage_channels(dtype, taps));
  l.cout = Cout; l.cout_pad = round_up(Cout, l.nt);
  if (taps == 9) l.ht = (l.nt == 128) ? (H >= 2 ? 2 : 1) : (H % 4 == 0 ? 4 : (H >= 2 ? 2 : 1));
  else l.ht = H >= 2 ? 2 : 1;
  PT probe; probe.B = B; probe.C = Cin; probe.H = H; probe.W = W;
  l.in0 = make_pt(sc, dtype, B, Cin, H, W, tensor_stats_slots(dtype, probe));
  l.out = make_pt(sc, dtype, B, l.cout_pad, H, W, 0);
  void* wp = sc.take(conv_packed_weight_bytes(dtype, taps, l.nt, l.cin_pad, l.cout_pad));
  float* bp = static_cast<float*>(sc.take(static_cast<size_t>(l.cout_pad) * 4));
  if (!l.in0.ptr || !l.in0.stats || !l.out.ptr || !wp || !bp) return fail(-1, "scratch too small");
  // NOTE: the input is stored unrounded here (it is a residual-stream tensor in the network)
  CUDA_TRY(pack_nchw(dtype, x, B, Cin, H, W, l.in0, 0, s));
  CUDA_TRY(tensor_stats_launch(dtype, l.in0, s));
  CUDA_TRY(pack_conv_weight(dtype, taps, l.nt, w, Cout, Cin, l.cin_pad, l.cout_pad, wp, s));
  CUDA_TRY(cudaMemsetAsync(bp, 0, static_cast<size_t>(l.cout_pad) * 4, s));
  if (bias) CUDA_TRY(cudaMemcpyAsync(bp, bias, static_cast<size_t>(Cout) * 4, cudaMemcpyDeviceToDevice, s));
  l.wpacked = wp; l.bias = bp; l.scale = 1.f;
  l.xf.enabled = 1; l.xf.silu = silu; l.xf.groups = kNU; l.xf.eps = eps;
  if (film) { l.xf.film = film; l.xf.film_stride = 2 * Cin; l.xf.film_off = 0; l.xf.row_batch_stride = 1; }
  else { l.xf.gamma = gamma; l.xf.beta = beta; }
  int rc = conv_make_tmaps(l);
  if (rc) return fail(-4, "tensor map encode failed (%d)", rc);
  CUDA_TRY(conv_launch(l, s));
  CUDA_TRY(unpack_nchw(dtype, l.out, y, 0, Cout, s));
  return 0;
}

// ResidualBlock tail with the folded skip projection, as the up-path blocks run it:
//   y = (conv3x3(silu(adagn(h, film))) + bias + conv1x1(xs, w2) + bias2) * scale
// h: [B][C][H][W] (C = Cout), xs: [B][Cs][H][W] raw block input, film: [B][2C], w: [C][C][3][3], w2: [C][Cs][1][1].
int r2dm_op_gn_conv_skip(int dtype, const float* hsrc, const float* film, float eps, const float* w, const float* bias,
                         const float* xs, const float* w2, const float* bias2, float scale, float* y, int B, int C,
                         int Cs, int H, int W, void* scratch, size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (W % 128) return fail(-1, "W must be a multiple of 128");
  if (C % (kNU * dtype_cw(dtype))) return fail(-1, "C must be a multiple of %d", kNU * dtype_cw(dtype));
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  ConvLaunch l;
  memset(&l, 0, sizeof(l));
  l.dtype = dtype; l.taps = 9;
  l.nt = pick_nt(C);
  l.cin_pad = round_up(C, conv_stage_channels(dtype, 9));
  if (l.cin_pad != C) return fail(-1, "C must be a multiple of the stage K");
  l.cout = C; l.cout_pad = round_up(C, l.nt);
  l.ht = (l.nt == 128) ? (H >= 2 ? 2 : 1) : (H % 4 == 0 ? 4 : (H >= 2 ? 2 : 1));
  const int skp = conv_skip_planes(l.nt) * dtype_cw(dtype);
  l.cin2_pad = round_up(Cs, skp);
  if (l.cin2_pad != Cs) return fail(-1, "Cs must be a multiple of the skip stage K (%d)", skp);
  PT probe; probe.B = B; probe.C = C; probe.H = H; probe.W = W;
  l.in0 = make_pt(sc, dtype, B, C, H, W, tensor_stats_slots(dtype, probe));
  l.sk0 = make_pt(sc, dtype, B, Cs, H, W, 0);
  l.out = make_pt(sc, dtype, B, l.cout_pad, H, W, 0);
  void* wp = sc.take(conv_packed_weight_bytes(dtype, 9, l.nt, l.cin_pad, l.cout_pad));
  void* w2p = sc.take(conv_packed_weight_bytes(dtype, 1, l.nt, l.cin2_pad, l.cout_pad));
  float* bp = static_cast<float*>(sc.take(static_cast<size_t>(l.cout_pad) * 4));
  float* b2p = static_cast<float*>(sc.take(static_cast<size_t>(l.cout_pad) * 4));
  if (!l.in0.ptr || !l.in0.stats || !l.sk0.ptr || !l.out.ptr || !wp || !w2p || !bp || !b2p) return fail(-1, "scratch too small");
  CUDA_TRY(pack_nchw(dtype, hsrc, B, C, H, W, l.in0, 0, s));
  CUDA_TRY(tensor_stats_launch(dtype, l.in0, s));
  CUDA_TRY(pack_nchw(dtype, xs, B, Cs, H, W, l.sk0, 0, s));
  CUDA_TRY(pack_conv_weight(dtype, 9, l.nt, w, C, C, l.cin_pad, l.cout_pad, wp, s));
  CUDA_TRY(pack_conv_weight(dtype, 1, l.nt, w2, C, Cs, l.cin2_pad, l.cout_pad, w2p, s, conv_skip_planes(l.nt)));
  CUDA_TRY(cudaMemsetAsync(bp, 0, static_cast<size_t>(l.cout_pad) * 4, s));
  CUDA_TRY(cudaMemsetAsync(b2p, 0, static_cast<size_t>(l.cout_pad) * 4, s));
  if (bias) CUDA_TRY(cudaMemcpyAsync(bp, bias, static_cast<size_t>(C) * 4, cudaMemcpyDeviceToDevice, s));
  if (bias2) CUDA_TRY(cudaMemcpyAsync(b2p, bias2, static_cast<size_t>(C) * 4, cudaMemcpyDeviceToDevice, s));
  l.wpacked = wp; l.bias = bp; l.w2packed = w2p; l.bias2 = b2p; l.scale = scale;
  l.xf.enabled = 1; l.xf.silu = 1; l.xf.groups = kNU; l.xf.eps = eps;
  l.xf.film = film; l.xf.film_stride = 2 * C; l.xf.film_off = 0; l.xf.row_batch_stride = 1;
  int rc = conv_make_tmaps(l);
  if (rc) return fail(-4, "tensor map encode failed (%d)", rc);
  CUDA_TRY(conv_launch(l, s));
  CUDA_TRY(unpack_nchw(dtype, l.out, y, 0, C, s));
  return 0;
}

int r2dm_op_groupnorm(int dtype, const float* x, const float* gamma, const float* beta, const float* film,
                      float eps, int silu, float* y, int B, int C, int H, int W, void* scratch,
                      size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C % (kNU * dtype_cw(dtype))) return fail(-1, "C must be a multiple of %d", kNU * dtype_cw(dtype));
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  PT probe; probe.B = B; probe.C = C; probe.H = H; probe.W = W;
  PT in = make_pt(sc, dtype, B, C, H, W, tensor_stats_slots(dtype, probe));
  PT out = make_pt(sc, dtype, B, C, H, W, 0);
  if (!in.ptr || !out.ptr || !in.stats) return fail(-1, "scratch too small");
  CUDA_TRY(pack_nchw(dtype, x, B, C, H, W, in, 0, s));
  CUDA_TRY(tensor_stats_launch(dtype, in, s));
  GnApply g;
  memset(&g, 0, sizeof(g));
  g.dtype = dtype; g.src0 = in; g.dst = out; g.groups = kNU; g.eps = eps; g.silu = silu;
  if (film) {
    g.film = film; g.film_stride = 2 * C; g.film_off = 0; g.row_batch_stride = 1; g.rows_per_step = 0;
  } else {
    g.gamma = gamma; g.beta = beta;
  }
  CUDA_TRY(gn_apply_launch(g, s));
  CUDA_TRY(unpack_nchw(dtype, out, y, 0, C, s));
  return 0;
}

int r2dm_op_resample(int dtype, int dir, const float* x, float* y, int B, int C, int H, int W, void* scratch,
                     size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C % dtype_cw(dtype)) return fail(-1, "C must be a multiple of %d", dtype_cw(dtype));
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  PT in = make_pt(sc, dtype, B, C, H, W, 0);
  const int Ho = dir > 0 ? H * 2 : H / 2, Wo = dir > 0 ? W * 2 : W / 2;
  PT out = make_pt(sc, dtype, B, C, Ho, Wo, 0);
  if (!in.ptr || !out.ptr) return fail(-1, "scratch too small");
  CUDA_TRY(pack_nchw(dtype, x, B, C, H, W, in, 0, s));
  if (dir > 0) CUDA_TRY(up2_launch(dtype, in, out, s));
  else CUDA_TRY(down2_launch(dtype, in, out, s));
  CUDA_TRY(unpack_nchw(dtype, out, y, 0, C, s));
  return 0;
}

int r2dm_op_attention(int dtype, const float* qkv, float* y, int B, int E, int heads, int H, int W, void* scratch,
                      size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  PT in = make_pt(sc, dtype, B, 3 * E, H, W, 0);
  PT out = make_pt(sc, dtype, B, E, H, W, 0);
  if (!in.ptr || !out.ptr) return fail(-1, "scratch too small");
  CUDA_TRY(pack_nchw(dtype, qkv, B, 3 * E, H, W, in, 0, s));
  if (dtype == kBF16 || get_option("attn_exact", 0) == 0) {
    CUtensorMap tm, tm_kv;
    int rc = attention_make_tmaps(&tm, &tm_kv, dtype, in, heads);
    if (rc) return fail(-4, "tensor map encode failed (%d)", rc);
    CUDA_TRY(attention_umma_launch(dtype, in, out, heads, tm, tm_kv, s));
  } else {
    CUDA_TRY(attention_launch(dtype, in, out, heads, s));
  }
  CUDA_TRY(unpack_nchw(dtype, out, y, 0, E, s));
  return 0;
}

// ---------------------------------------------------------------------------------- PointNet features
namespace {
int pointnet_image_width(int N) {
  for (int w : {1024, 512, 256, 128})
    if (N % w == 0) return w;
  return 0;
}
// one point-wise layer (Conv1d k=1 with folded BatchNorm) as a 1x1 convolution over the [H x W] point image
int pointwise_layer(int dtype, Scratch& sc, const PT& in, int cin, int cout, const float* w, const float* b, int relu,
                    PT* out, float* colmax, cudaStream_t s) {
  ConvLaunch l;
  memset(&l, 0, sizeof(l));
  l.dtype = dtype; l.taps = 1;
  l.nt = pick_nt(cout);
  l.cin_pad = round_up(cin, conv_stage_channels(dtype, 1));
  if (in.C != l.cin_pad) return fail(-1, "pointnet: layer input has %d channel planes, expected %d", in.C, l.cin_pad);
  l.cout = cout; l.cout_pad = round_up(cout, l.nt);
  l.ht = in.H >= 2 ? 2 : 1;
  l.in0 = in;
  if (out != nullptr) {
    // the consumer reads whole K stages: the padded width must already be a multiple of the stage K
    if (l.cout_pad % conv_stage_channels(dtype, 1)) return fail(-1, "pointnet: layer width %d not a stage multiple", l.cout_pad);
    *out = make_pt(sc, dtype, in.B, l.cout_pad, in.H, in.W, 0);
    if (!out->ptr) return fail(-1, "pointnet: scratch too small");
    l.out = *out;
  } else {
    l.out.B = in.B; l.out.C = l.cout_pad; l.out.H = in.H; l.out.W = in.W;   // dimensions only: nothing is stored
  }
  void* wp = sc.take(conv_packed_weight_bytes(dtype, 1, l.nt, l.cin_pad, l.cout_pad));
  float* bp = static_cast<float*>(sc.take(static_cast<size_t>(l.cout_pad) * 4));
  if (!wp || !bp) return fail(-1, "pointnet: scratch too small");
  CUDA_TRY(pack_conv_weight(dtype, 1, l.nt, w, cout, cin, l.cin_pad, l.cout_pad, wp, s));
  CUDA_TRY(cudaMemsetAsync(bp, 0, static_cast<size_t>(l.cout_pad) * 4, s));
  CUDA_TRY(cudaMemcpyAsync(bp, b, static_cast<size_t>(cout) * 4, cudaMemcpyDeviceToDevice, s));
  l.wpacked = wp; l.bias = bp; l.scale = 1.f;
  l.relu = relu; l.colmax = colmax;
  int rc = conv_make_tmaps(l);
  if (rc) return fail(-4, "tensor map encode failed (%d)", rc);
  CUDA_TRY(conv_launch(l, s));
  return 0;
}
}  // namespace

size_t r2dm_pointnet_scratch_bytes(int dtype, int batch, int num_points) {
  const int W = pointnet_image_width(num_points);
  if (W == 0 || batch < 1) return 0;
  PT t; t.B = batch; t.H = num_points / W; t.W = W;
  size_t total = 0;
  for (int c : {64, 64, 128}) { t.C = round_up(c, conv_stage_channels(dtype, 1)); total += align_up(t.bytes(dtype), 1024); }
  total *= 2;                                                          // both trunks
  total += static_cast<size_t>(batch) * 3 * num_points * sizeof(float);   // transformed points
  total += 8u << 20;                                                   // packed weights, biases, pooled vectors
  return total;
}

int r2dm_pointnet_features(int dtype, const float* points, const r2dm_pointnet_weights* w, float* features, int batch,
                           int num_points, void* scratch, size_t scratch_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!points || !w || !features || !scratch) return fail(-1, "null argument");
  if (dtype != kF32 && dtype != kBF16) return fail(-1, "dtype must be R2DM_F32 or R2DM_BF16");
  const int W = pointnet_image_width(num_points);
  if (W == 0 || batch < 1) return fail(-1, "num_points must be a multiple of 128");
  for (int i = 0; i < 12; ++i)
    if (!w->weight[i] || !w->bias[i]) return fail(-1, "pointnet weight %d missing", i);
  const int k = w->num_classes;
  if (k < 1 || k > 4096) return fail(-1, "bad num_classes");
  const int H = num_points / W, B = batch, FD = 1024 + 512 + 256 + k;
  Scratch sc{static_cast<uint8_t*>(scratch), scratch_bytes};
  float* pooled = static_cast<float*>(sc.take(static_cast<size_t>(B) * 1024 * 4));
  float* v512 = static_cast<float*>(sc.take(static_cast<size_t>(B) * 512 * 4));
  float* v256 = static_cast<float*>(sc.take(static_cast<size_t>(B) * 256 * 4));
  float* trans = static_cast<float*>(sc.take(static_cast<size_t>(B) * 9 * 4));
  float* moved = static_cast<float*>(sc.take(static_cast<size_t>(B) * 3 * num_points * 4));
  if (!pooled || !v512 || !v256 || !trans || !moved) return fail(-1, "scratch too small");
  const int cin0 = round_up(3, conv_stage_channels(dtype, 1));
  // trunk(points, first conv index) -> pooled [B][1024]; relu_last: STN3d applies ReLU before the pool (:26-27),
  // PointNetfeat does not (:56-57)
  auto trunk = [&](const float* pts, int wi, int relu_last) -> int {
    PT in = make_pt(sc, dtype, B, cin0, H, W, 0);
    if (!in.ptr) return fail(-1, "scratch too small");
    CUDA_TRY(cudaMemsetAsync(in.ptr, 0, in.bytes(dtype), s));
    CUDA_TRY(pack_nchw(dtype, pts, B, 3, H, W, in, 0, s));
    PT h1, h2;
    int rc = pointwise_layer(dtype, sc, in, 3, 64, w->weight[wi], w->bias[wi], 1, &h1, nullptr, s);
    if (rc) return rc;
    rc = pointwise_layer(dtype, sc, h1, 64, 128, w->weight[wi + 1], w->bias[wi + 1], 1, &h2, nullptr, s);
    if (rc) return rc;
    CUDA_TRY(fill_launch(pooled, -INFINITY, static_cast<size_t>(B) * 1024, s));
    return pointwise_layer(dtype, sc, h2, 128, 1024, w->weight[wi + 2], w->bias[wi + 2], relu_last, nullptr, pooled, s);
  };
  // STN3d (pointnet.py:22-33)
  int rc = trunk(points, 0, 1);
  if (rc) return rc;
  CUDA_TRY(dense_launch(pooled, 1024, w->weight[3], w->bias[3], v512, 512, B, 1024, 512, 1, s));
  CUDA_TRY(dense_launch(v512, 512, w->weight[4], w->bias[4], v256, 256, B, 512, 256, 1, s));
  CUDA_TRY(dense_launch(v256, 256, w->weight[5], w->bias[5], trans, 9, B, 256, 9, 0, s));
  // PointNetfeat (:47-58) on the transformed points
  CUDA_TRY(point_transform_launch(points, trans, moved, B, num_points, s));
  rc = trunk(moved, 6, 0);
  if (rc) return rc;
  // PointNet1 head (:74-81): feature = cat(x1, x2, x3, x4)
  CUDA_TRY(cudaMemcpy2DAsync(features, static_cast<size_t>(FD) * 4, pooled, 1024 * 4, 1024 * 4, B,
                             cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(dense_launch(pooled, 1024, w->weight[9], w->bias[9], features + 1024, FD, B, 1024, 512, 1, s));
  CUDA_TRY(dense_launch(features + 1024, FD, w->weight[10], w->bias[10], features + 1536, FD, B, 512, 256, 1, s));
  CUDA_TRY(dense_launch(features + 1536, FD, w->weight[11], w->bias[11], features + 1792, FD, B, 256, k, 0, s));
  return 0;
}

int r2dm_set_option(const char* name, int value) {
  if (name && set_option(name, value) == 0) return 0;
  return fail(-1, "unknown option %s", name ? name : "(null)");
}

// developer: per-launch (first CTA start, last CTA end) of every convolution of the forward, written by the
// kernels themselves (works inside CUDA graphs): buf = device uint64 [num_launches][2], or NULL to switch off.
// Takes effect for launches / graph captures made afterwards.
int r2dm_debug_set_ktime(r2dm_handle h, void* buf) {
  if (!h) return fail(-1, "null argument");
  int i = 0;
  for (Op& op : h->prog) {
    if (op.kind == Op::CONV)
      op.conv.ktime = buf ? static_cast<unsigned long long*>(buf) + 2 * i : nullptr;
    ++i;
  }
  return 0;
}

int r2dm_debug_set_trace(void* buf, int cap) {
  conv_set_trace(static_cast<unsigned long long*>(buf), cap);
  return 0;
}

int r2dm_debug_tensor(r2dm_handle h, const char* name, float* out, int* C, int* H, int* W, void* stream) {
  if (!h || !name) return fail(-1, "null argument");
  auto it = h->named.find(name);
  if (it == h->named.end()) return fail(-5, "no tensor named %s", name);
  const PT& t = it->second;
  if (C) *C = t.C;
  if (H) *H = t.H;
  if (W) *W = t.W;
  if (out) CUDA_TRY(unpack_nchw(h->dtype, t, out, 0, t.C, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
