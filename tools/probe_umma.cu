// Hardware probe (developer tool, not part of the product path): checks on a real B200 which
// tcgen05 shared-memory descriptor conventions hold for the "shifted view" trick the ring-conv
// kernel relies on (one halo tile in smem, 9 tap operands = 9 start addresses), plus the 5-D TMA
// box load with out-of-bounds zero fill.  Build: see tools/build_probe.sh.  Prints PASS/FAIL lines.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../r2dm_b200/csrc/ptx.cuh"

using namespace r2dm;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

struct MmaOp {
  uint32_t a_off, b_off;      // byte offsets from the 1024-aligned smem base
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;
  uint32_t layout;            // 0 none, 2 sw128
  uint32_t base_mode;         // 0: base_offset=0, 1: (addr>>7)&7
  uint32_t accumulate;
};
struct CaseParams {
  uint32_t image_bytes;
  uint32_t n_ops;
  uint32_t N;
  uint32_t fmt;  // 1 bf16, 2 tf32
  MmaOp ops[8];
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ image, CaseParams p, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32;
  for (uint32_t i = threadIdx.x * 16; i < p.image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, p.N, p.fmt);
    const uint32_t sbase = smem_u32(smem);
    for (uint32_t i = 0; i < p.n_ops; ++i) {
      const MmaOp& o = p.ops[i];
      uint32_t aa = sbase + o.a_off, ba = sbase + o.b_off;
      uint32_t boa = o.base_mode ? ((aa >> 7) & 7) : 0;
      uint32_t bob = o.base_mode ? ((ba >> 7) & 7) : 0;
      uint64_t da = make_smem_desc(aa, o.lbo_a, o.sbo_a, o.layout, boa);
      uint64_t db = make_smem_desc(ba, o.lbo_b, o.sbo_b, o.layout, bob);
      if (p.fmt == 2) umma_tf32(tmem, da, db, idesc, o.accumulate);
      else umma_f16(tmem, da, db, idesc, o.accumulate);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = threadIdx.x;  // TMEM lane
  for (uint32_t c = 0; c < p.N; c += 8) {
    float v[8];
    tmem_ld8(tmem + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[row * p.N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------ TMA probe
__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, int c4,
                 uint32_t box_bytes, uint8_t* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (uint32_t i = threadIdx.x; i < box_bytes; i += blockDim.x) smem[i] = 0xAB;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, box_bytes);
    tma_load_5d(smem, &tmap, &bar, c0, c1, c2, c3, c4);
  }
  mbar_wait(&bar, 0);
  for (uint32_t i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

// ------------------------------------------------------------------ host
static const int P = 400, K = 64, NMAX = 256;
static std::vector<float> X, Wt;  // X[P][K], Wt[NMAX][K]

static void put(std::vector<uint8_t>& img, size_t off, float v, int fmt) {
  if (fmt == 1) {
    __nv_bfloat16 b = __float2bfloat16(v);
    memcpy(&img[off], &b, 2);
  } else {
    memcpy(&img[off], &v, 4);
  }
}

int main() {
  srand(1);
  X.resize(P * K);
  Wt.resize(NMAX * K);
  for (auto& v : X) v = float(rand() % 9 - 4);
  for (auto& v : Wt) v = float(rand() % 7 - 3) * 0.5f;

  uint8_t* d_img;
  float* d_out;
  CK(cudaMalloc(&d_img, 200 * 1024));
  CK(cudaMalloc(&d_out, 128 * NMAX * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

  struct Spec { const char* name; int layout; int shift; int ksteps; int N; int base_mode; int fmt; };
  Spec specs[] = {
      {"planar  s=0   k=1 N=64", 0, 0, 1, 64, 0, 1},
      {"planar  s=0   k=4 N=64", 0, 0, 4, 64, 0, 1},
      {"planar  s=1   k=4 N=64", 0, 1, 4, 64, 0, 1},
      {"planar  s=131 k=4 N=64", 0, 131, 4, 64, 0, 1},
      {"planar  s=5   k=4 N=128", 0, 5, 4, 128, 0, 1},
      {"planar  s=5   k=4 N=256", 0, 5, 4, 256, 0, 1},
      {"planar  s=2   k=4 N=16", 0, 2, 4, 16, 0, 1},
      {"sw128   s=0   k=1 N=64", 2, 0, 1, 64, 0, 1},
      {"sw128   s=0   k=4 N=64", 2, 0, 4, 64, 0, 1},
      {"sw128   s=8   k=4 N=64", 2, 8, 4, 64, 0, 1},
      {"sw128   s=1   k=4 N=64 base0", 2, 1, 4, 64, 0, 1},
      {"sw128   s=1   k=4 N=64 baseC", 2, 1, 4, 64, 1, 1},
      {"sw128   s=3   k=4 N=64 base0", 2, 3, 4, 64, 0, 1},
      {"sw128   s=3   k=4 N=64 baseC", 2, 3, 4, 64, 1, 1},
      {"tf32 planar s=0 k=4 N=64", 0, 0, 4, 64, 0, 2},
      {"tf32 planar s=3 k=4 N=64", 0, 3, 4, 64, 0, 2},
      {"tf32 planar s=3 k=4 N=256", 0, 3, 4, 256, 0, 2},
  };
  for (const Spec& s : specs) {
    const int es = (s.fmt == 1) ? 2 : 4;     // element size
    const int cw = 16 / es;                  // channels per 16-byte chunk
    const int kk = 32 / es;                  // K per MMA
    const int Kuse = s.ksteps * kk;
    std::vector<uint8_t> img(200 * 1024, 0);
    CaseParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.N = s.N; cp.fmt = s.fmt; cp.n_ops = s.ksteps;
    uint32_t offX = 0, offW;
    if (s.layout == 0) {
      const uint32_t planeX = P * 16, planeW = s.N * 16;
      const int chunks = K / cw;
      offW = ((offX + chunks * planeX) + 1023) & ~1023u;
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < Kuse; ++k)
          put(img, offX + (k / cw) * planeX + p * 16 + (k % cw) * es, X[p * K + k], s.fmt);
      for (int n = 0; n < s.N; ++n)
        for (int k = 0; k < Kuse; ++k)
          put(img, offW + (k / cw) * planeW + n * 16 + (k % cw) * es, Wt[n * K + k], s.fmt);
      cp.image_bytes = offW + chunks * planeW;
      for (int j = 0; j < s.ksteps; ++j) {
        MmaOp& o = cp.ops[j];
        o.a_off = offX + 2 * j * planeX + s.shift * 16;
        o.b_off = offW + 2 * j * planeW;
        o.lbo_a = planeX; o.sbo_a = 128; o.lbo_b = planeW; o.sbo_b = 128;
        o.layout = 0; o.base_mode = 0; o.accumulate = j > 0;
      }
    } else {
      offW = ((offX + P * 128) + 1023) & ~1023u;
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k)
          put(img, offX + p * 128 + (((k / 8) ^ (p & 7)) * 16) + (k % 8) * 2, X[p * K + k], 1);
      for (int n = 0; n < s.N; ++n)
        for (int k = 0; k < K; ++k)
          put(img, offW + n * 128 + (((k / 8) ^ (n & 7)) * 16) + (k % 8) * 2, Wt[n * K + k], 1);
      cp.image_bytes = offW + s.N * 128;
      for (int j = 0; j < s.ksteps; ++j) {
        MmaOp& o = cp.ops[j];
        o.a_off = offX + s.shift * 128 + j * 32;
        o.b_off = offW + j * 32;
        o.lbo_a = 16; o.sbo_a = 1024; o.lbo_b = 16; o.sbo_b = 1024;
        o.layout = 2; o.base_mode = s.base_mode; o.accumulate = j > 0;
      }
    }
    cp.image_bytes = (cp.image_bytes + 15) & ~15u;
    CK(cudaMemcpy(d_img, img.data(), cp.image_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, 128 * NMAX * 4));
    probe_kernel<<<1, 128, cp.image_bytes + 2048>>>(d_img, cp, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("FAIL  %-34s launch error: %s\n", s.name, cudaGetErrorString(e));
      return 2;  // context is dead
    }
    std::vector<float> out(128 * s.N);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < s.N; ++n) {
        double ref = 0;
        for (int k = 0; k < Kuse; ++k) ref += double(X[(m + s.shift) * K + k]) * Wt[n * K + k];
        double err = fabs(ref - out[m * s.N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) ++bad;
      }
    printf("%s  %-34s maxerr=%g bad=%d/%d\n", bad ? "FAIL" : "PASS", s.name, maxerr, bad, 128 * s.N);
  }

  // ---------------- TMA 5-D probe: tensor [Bn][C8][H][Wp][8] bf16, box (8, bw, bh, bc, 1)
  {
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    const int Bn = 2, C8 = 3, H = 6, Wp = 34;
    std::vector<__nv_bfloat16> T(size_t(Bn) * C8 * H * Wp * 8);
    for (size_t i = 0; i < T.size(); ++i) T[i] = __float2bfloat16(float(i % 251) - 125.f);
    __nv_bfloat16* d_T;
    CK(cudaMalloc(&d_T, T.size() * 2));
    CK(cudaMemcpy(d_T, T.data(), T.size() * 2, cudaMemcpyHostToDevice));
    const int bw = 10, bh = 4, bc = 2;
    CUtensorMap tmap;
    cuuint64_t dims[5] = {8, (cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)Bn};
    cuuint64_t strides[4] = {16, (cuuint64_t)Wp * 16, (cuuint64_t)H * Wp * 16, (cuuint64_t)C8 * H * Wp * 16};
    cuuint32_t box[5] = {8, bw, bh, bc, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d_T, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("FAIL  tensor map encode: %d\n", (int)r); return 3; }
    const uint32_t box_bytes = 16 * bw * bh * bc;
    uint8_t* d_o;
    CK(cudaMalloc(&d_o, box_bytes));
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int coords[][5] = {{0, 3, -1, 1, 1}, {0, 0, 4, 0, 0}, {0, 26, 1, 1, 0}};
    for (auto& c : coords) {
      tma_probe_kernel<<<1, 128, box_bytes + 2048>>>(tmap, c[0], c[1], c[2], c[3], c[4], box_bytes, d_o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("FAIL  tma probe launch: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<__nv_bfloat16> o(box_bytes / 2);
      CK(cudaMemcpy(o.data(), d_o, box_bytes, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int cc = 0; cc < bc; ++cc)
        for (int y = 0; y < bh; ++y)
          for (int x = 0; x < bw; ++x)
            for (int e8 = 0; e8 < 8; ++e8) {
              int gy = c[2] + y, gx = c[1] + x, gc = c[3] + cc, gb = c[4];
              float ref = 0.f;
              if (gy >= 0 && gy < H && gx >= 0 && gx < Wp && gc < C8)
                ref = __bfloat162float(T[(((size_t(gb) * C8 + gc) * H + gy) * Wp + gx) * 8 + e8]);
              float got = __bfloat162float(o[((size_t(cc) * bh + y) * bw + x) * 8 + e8]);
              if (ref != got) ++bad;
            }
      printf("%s  tma5d box@(%d,%d,%d,%d,%d) bad=%d\n", bad ? "FAIL" : "PASS", c[0], c[1], c[2], c[3], c[4], bad);
    }
  }
  return 0;
}
