// tc_conv_px8.cu — first conv stage (3x3, Cin = 3, pad 1, stride 1 + bias + ReLU + MaxPool2) on tcgen05 with a TMA-built
// A operand (round 2).  Replaces stage 1 of vanilla_encoder (reference networks/fcn.py:25-31) and block1_conv1 of VGG16
// (vgg16.py:27-29); the input is what prediction.py:82-84 / data/generator.py:53-61 hand the model.
//
// Round 1's kernel (tc_conv_first.cu) built the im2col rows with per-thread shared-memory gathers: ~470 instructions per thread
// per 128-pixel tile, i.e. issue-bound at 0.105 ms per 256 faces against a TMEM-read floor of 0.059 ms.  Here the image is first
// widened to 8 bf16 channels per pixel (RGB + 5 zeros = 16 bytes, one pass, stays in L2) so that ONE K group of the UMMA
// no-swizzle K-major layout (8 elements = 16 bytes per row) is exactly one tap of one pixel.  ONE TMA box per tile brings the
// (16 + 2) x (8 + 2)-pixel halo patch (18 rows of 160 bytes; out-of-bounds fill = the convolution's zero padding); the nine
// taps are nine UMMA descriptors INTO that patch: start shifted by (ky * 10 + kx) pixels, stride between 8-row core matrices
// = one patch row (160 B), and — K = 16 per MMA = two taps — a leading-dimension offset equal to the distance between the two
// taps' starts (16 B, or 128 B across a patch row).  No thread touches an operand.  (A first version with nine [8 ch][8 px][16 px]
// boxes per tile was TMA-request bound — 1152 sixteen-byte rows per tile — and slower than round 1's kernel: 0.20 ms.)
//
// Roles (320 threads, one persistent CTA per SM): warp 0 TMA producer (ring of A stages), warp 1 MMA issuer (K = 16 per MMA =
// two taps; the tenth K group is a zero block), warps 2-9 epilogue (TMEM -> bias + ReLU + 2x2 pool -> bf16 / SPLIT store), two
// TMEM accumulators.  FLD_BF16X3 (uint8 input, exact in bf16): the weights are split hi / lo and the SAME A groups are
// multiplied a second time against the lo block (ten MMAs, no extra operand traffic).
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kAGroups = 10;                     // 9 taps + 1 zero group (K = 80)
constexpr int kPatchW = 10, kPatchH = 18;        // halo patch of an 8 x 16 tile, 16 B per pixel
constexpr int kRowBytes = kPatchW * 16;          // 160
constexpr int kPatchBytes = kPatchH * kRowBytes; // 2880
constexpr int kStageBytes = 3072;                // patch rounded up to the 128-byte TMA destination alignment
constexpr int kZeroBytes = 16 * kRowBytes;       // zero K group: 16 core matrices at the same 160-byte stride
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 6;

struct Px8Params {
  const __nv_bfloat16* w;   // [KGB][Cout/8][8 rows][8 k] core-matrix packed; KGB = 10 (bf16) or 20 (x3: hi block, lo block)
  const float* bias;
  void* out;
  int B, H, W, Cout;
  int act, pool, split, x3;
  int tiles_x, tiles_y, total_tiles;
  int stages;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_px8_kernel(const __grid_constant__ CUtensorMap tmA, const Px8Params p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int KGB = p.x3 ? 2 * kAGroups : kAGroups;
  const uint32_t smemA = smem_base;                                       // stages x halo patch [18][10 px][16 B]
  const uint32_t smemZ = smem_base + p.stages * kStageBytes;               // zero K group
  const uint32_t smemB = smemZ + kZeroBytes;                               // [KGB][Cout/8][8][16 B]
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  {  // weights (already in core-matrix order) and the all-zero tenth K group
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint8_t* gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    uint4* z = reinterpret_cast<uint4*>(gen + p.stages * kStageBytes);
    for (int i = tid; i < kZeroBytes / 16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    uint4* dstB = reinterpret_cast<uint4*>(gen + p.stages * kStageBytes + kZeroBytes);
    for (int i = tid; i < p.Cout * KGB; i += kThreads) dstB[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiWarps); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_async_smem();      // the generic-proxy writes above are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int txy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / txy;
        const int m = tile - b * txy;
        const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = full0 + 8 * stage;
        mbar_arrive_expect_tx(fb, (uint32_t)kPatchBytes);
        tma_load_3d(smemA + stage * kStageBytes, &tmA, fb, (tx * 8 - 1) * 8, ty * 16 - 1, b);   // inner coordinate in bf16 elements
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.Cout);
      const uint32_t b_lbo = (uint32_t)(p.Cout / 8) * 128;                 // distance between K groups of B
      const uint64_t bdesc0 = umma_desc(smemB, b_lbo, 128, 0);             // K-major, no swizzle: LBO = next K group, SBO = next 8 rows
      const uint64_t bstep = (uint64_t)((2 * b_lbo) >> 4);
      // tap t starts (t / 3) patch rows + (t % 3) pixels into the patch
      auto tap_off = [](int t) { return (uint32_t)(((t / 3) * kPatchW + (t % 3)) * 16); };
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        const uint32_t patch = smemA + stage * kStageBytes;
        uint64_t ad[kAGroups / 2];
#pragma unroll
        for (int m = 0; m < kAGroups / 2; ++m) {
          const uint32_t s0 = patch + tap_off(2 * m);
          // K = 16 = taps 2m and 2m + 1: the second K group starts LBO bytes after the first; the tenth group is the zero block
          const uint32_t lbo = (2 * m + 1 < 9) ? tap_off(2 * m + 1) - tap_off(2 * m) : smemZ - s0;
          ad[m] = umma_desc(s0, lbo, kRowBytes, 0);                        // SBO: next 8 tile pixels = next patch row
        }
#pragma unroll
        for (int m = 0; m < kAGroups / 2; ++m) umma_bf16(d, ad[m], bdesc0 + m * bstep, idesc, m ? 1u : 0u);
        if (p.x3) {   // the same activations against the lo halves of the weights
#pragma unroll
          for (int m = 0; m < kAGroups / 2; ++m) umma_bf16(d, ad[m], bdesc0 + (kAGroups / 2 + m) * bstep, idesc, 1u);
        }
        umma_commit(empty0 + 8 * stage);
        umma_commit(tfull0 + 8 * acc);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (TW = 8, TH = 16: pool partners lane^1, lane^8)
    const int ew = warp - 2, sub = warp & 3, half = ew >> 2;
    const int r = sub * 32 + lane;
    const int lx = r & 7, ly = r >> 3;
    const int PH = p.H >> 1, PW = p.W >> 1;
    const int pitch = p.split ? 2 * p.Cout : p.Cout;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / txy;
      const int m = tile - b * txy;
      const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
      const int ox = tx * 8 + lx, oy = ty * 16 + ly;
      EpiOut eo;
      eo.vec_ok = true;   // Cout % 16 == 0
      size_t pix;
      if (p.pool) {
        eo.valid = ((oy >> 1) < PH) && ((ox >> 1) < PW);
        pix = ((size_t)b * PH + (oy >> 1)) * PW + (ox >> 1);
      } else {
        eo.valid = (oy < p.H) && (ox < p.W);
        pix = ((size_t)b * p.H + oy) * p.W + ox;
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      for (int ch = half * 32; ch < p.Cout; ch += 64) {
        uint32_t regs[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + ch, regs);
        tmem_ld_wait();
        EpiOut e2 = eo;
        e2.c_left = p.Cout - ch;
        e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * pitch + ch;
        if (p.split) {
          if (p.pool) epilogue_chunk_split<true>(regs, p.bias + ch, p.act, lane, 8, e2, p.Cout);
          else epilogue_chunk_split<false>(regs, p.bias + ch, p.act, lane, 8, e2, p.Cout);
        } else {
          if (p.pool) epilogue_chunk<true, false>(regs, p.bias + ch, p.act, lane, 8, e2);
          else epilogue_chunk<false, false>(regs, p.bias + ch, p.act, lane, 8, e2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// [B][H][W][3] uint8 / float32  ->  [B][H][W][8] bf16 (channels 3..7 zero): 4 pixels per thread, 16-byte stores
template <typename TIn>
__global__ void widen_px8_kernel(const TIn* __restrict__ in, uint4* __restrict__ out, long long n_px, int aligned4) {
  const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (g >= n_px) return;
  const int n = (int)min((long long)4, n_px - g);
  float v[12];
  if (sizeof(TIn) == 1 && n == 4 && aligned4) {                      // 12 bytes = three aligned words
    const uint32_t* q = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(in) + g * 3);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (float)((w0 >> (8 * i)) & 255u); v[4 + i] = (float)((w1 >> (8 * i)) & 255u); v[8 + i] = (float)((w2 >> (8 * i)) & 255u); }
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = (i < 3 * n) ? (float)__ldg(in + g * 3 + i) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < n) out[g + i] = make_uint4(pack_bf16(v[3 * i], v[3 * i + 1]), pack_bf16(v[3 * i + 2], 0.f), 0u, 0u);
}

}  // namespace

struct TcPx8Plan {
  CUtensorMap tmA;
  Px8Params p;
  int grid;
  size_t smem;
  void* scratch;
  int in_dtype;
};

bool tc_conv_px8_supported(const ConvGeom& g) {
  static const bool off = getenv("FLD_C1_PX8_OFF") != nullptr;
  return !off && g.kh == 3 && g.kw == 3 && g.Cin == 3 && g.stride == 1 && g.pad_t == 1 && g.pad_l == 1 && g.OH == g.IH && g.OW == g.IW &&
         g.Cout % 16 == 0 && g.Cout >= 16 && g.Cout <= 256 && (g.pool == 0 || g.pool == 2);
}

size_t tc_conv_px8_scratch_bytes(const ConvGeom& g, int B) { return (size_t)B * g.IH * g.IW * 16; }

// w_host fp32 [27][Cout] (k = (kh*3+kw)*3 + c)  ->  bf16 [KGB][Cout/8][8][8]: K group t (< 9) = tap t, element e (< 3) = channel e;
// x3: groups 0..9 hold bf16(w), groups 10..19 the remainders bf16(w - bf16(w)).  The bias is added in the epilogue.
void tc_conv_px8_pack(const float* w_host, int Cout, uint16_t (*f2bf)(float), uint16_t* out, int x3) {
  const int KGB = x3 ? 2 * kAGroups : kAGroups;
  for (int kg = 0; kg < KGB; ++kg)
    for (int ng = 0; ng < Cout / 8; ++ng)
      for (int r = 0; r < 8; ++r)
        for (int e = 0; e < 8; ++e) {
          const int t = kg % kAGroups, o = ng * 8 + r;
          float v = 0.f;
          if (t < 9 && e < 3) {
            const float w = w_host[(size_t)(t * 3 + e) * Cout + o];
            if (kg < kAGroups) v = w;
            else {
              const uint32_t hu = (uint32_t)f2bf(w) << 16;
              float hf;
              memcpy(&hf, &hu, 4);
              v = w - hf;
            }
          }
          out[(((size_t)kg * (Cout / 8) + ng) * 8 + r) * 8 + e] = f2bf(v);
        }
}

int tc_conv_px8_plan_create(const fld_handle* h, void* scratch, int in_dtype, const ConvGeom& g, int B, int x3, int split_out,
                            TcPx8Plan** out) {
  if (!h->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  if (x3 && in_dtype != FLD_U8) { fld_set_error("tc_conv_px8: the FLD_BF16X3 variant takes a uint8 input"); return FLD_ERR_INVALID; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  TcPx8Plan* pl = new TcPx8Plan();
  Px8Params& p = pl->p;
  p.w = nullptr; p.bias = nullptr; p.out = nullptr;
  p.B = B; p.H = g.IH; p.W = g.IW; p.Cout = g.Cout; p.act = g.act; p.pool = g.pool; p.split = split_out; p.x3 = x3;
  p.tiles_x = fld_div_up(g.OW, 8); p.tiles_y = fld_div_up(g.OH, 16);
  p.total_tiles = B * p.tiles_x * p.tiles_y;
  const size_t bbytes = (size_t)g.Cout * 16 * (x3 ? 2 * kAGroups : kAGroups);
  p.stages = kMaxStages;
  pl->smem = (size_t)p.stages * kStageBytes + kZeroBytes + bbytes + 1024;
  pl->grid = std::min(p.total_tiles, h->sm_count);
  pl->scratch = scratch; pl->in_dtype = in_dtype;
  // the widened image as rows of W * 8 bf16 elements: a box row is 10 pixels = 160 contiguous bytes
  cuuint64_t dims[3] = {(cuuint64_t)g.IW * 8, (cuuint64_t)g.IH, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)g.IW * 16, (cuuint64_t)g.IH * g.IW * 16};
  cuuint32_t box[3] = {(cuuint32_t)kPatchW * 8, (cuuint32_t)kPatchH, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, scratch, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(px8 A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  *out = pl;
  return FLD_OK;
}

void tc_conv_px8_plan_destroy(TcPx8Plan* p) { delete p; }

int tc_conv_px8_run(const TcPx8Plan* pl, const void* in, const __nv_bfloat16* w_packed, const float* bias, void* out, cudaStream_t st) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  const long long n_px = (long long)pl->p.B * pl->p.H * pl->p.W;
  const long long blocks = (n_px / 4 + 255) / 256 + 1;
  if (blocks >= (1ll << 31)) { fld_set_error("tc_conv_px8: too many pixels"); return FLD_ERR_INVALID; }
  const int al = (reinterpret_cast<uintptr_t>(in) & 3) == 0;
  if (pl->in_dtype == FLD_U8) widen_px8_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)in, (uint4*)pl->scratch, n_px, al);
  else if (pl->in_dtype == FLD_F32) widen_px8_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, (uint4*)pl->scratch, n_px, al);
  else { fld_set_error("tc_conv_px8: input must be u8 or f32"); return FLD_ERR_INVALID; }
  FLD_LAUNCHED();
  Px8Params p = pl->p;
  p.w = w_packed; p.bias = bias; p.out = out;
  FLD_CUDA(cudaFuncSetAttribute(conv_px8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  conv_px8_kernel<<<pl->grid, kThreads, pl->smem, st>>>(pl->tmA, p);
  FLD_LAUNCHED();
  return FLD_OK;
}
