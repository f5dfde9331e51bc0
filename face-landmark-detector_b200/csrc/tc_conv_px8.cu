// tc_conv_px8.cu — first conv stage (3x3, Cin = 3, pad 1, stride 1 + bias + ReLU + MaxPool2) on tcgen05 with a TMA-built
// A operand (round 2).  Replaces stage 1 of vanilla_encoder (reference networks/fcn.py:25-31) and block1_conv1 of VGG16
// (vgg16.py:27-29); the input is what prediction.py:82-84 / data/generator.py:53-61 hand the model.
//
// STATUS: opt-in experiment (FLD_C1_PX8=1), correct (same parity tests as the default kernel) but not faster — see the end of
// this comment.
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
// Roles (576 threads, one persistent CTA per SM): warp 0 TMA producer (ring of A stages), warp 1 MMA issuer (K = 16 per MMA =
// two taps; the tenth K group is a zero block), warps 2-17 epilogue (TMEM -> bias + ReLU + 2x2 pool -> bf16 / SPLIT store) in
// four groups that take tiles round-robin over a ring of eight TMEM accumulators.  MEASURED: a tile is only 5 MMAs, so the
// kernel lives or dies by epilogue latency — with one group of 8 warps working on every tile in turn it took 0.18 ms per 256
// faces whatever the accumulator depth (each warp's TMEM load -> shuffle chain -> store is ~1500 dependent cycles per tile).
// Four groups: 0.127 ms; bias through the tenth K group (A = ones, B = bias split three ways) instead of 8 loads + 32 adds per
// chunk: 0.118 ms; stage-ring depth 6 vs 20, per-tile divisions removed, POOL / SPLIT as template parameters: no change.  ncu:
// 2000 warp instructions per tile (the packed-bf16 pooling epilogue: 16 F2FP + 24 SEL + 12 SHFL + 16 HMNMX2 per 32-column chunk),
// issue 53 %, and the epilogue warps still spend 40 % of their samples waiting for `tfull` — the same ~1900 instructions per tile
// as round 1's kernel, whose eight resident CTAs hide the latencies better.  With the 0.022 ms widening pass it loses: 0.141 ms.  FLD_BF16X3 (uint8 input, exact in bf16): the weights are split hi / lo and the SAME A groups are
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
constexpr int kZeroBytes = 16 * kRowBytes;       // zero K group: 16 core matrices at the same 160-byte stride
constexpr int kStageBytes = (kPatchBytes + kZeroBytes + 127) / 128 * 128;   // patch + its own zero block (constant descriptor offsets)
constexpr int kEpiWarps = 16;                    // four groups of four warps (one per TMEM lane quadrant): four tiles' epilogues in flight
constexpr int kEpiGroups = kEpiWarps / 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 20;                   // deep ring: a stage is 5.4 KB, an L2 / HBM round trip is ~1-2 us, a tile ~0.3 us
constexpr int kMaxAcc = 8;                       // TMEM accumulators in flight: 512 columns / Cout rounded up to a power of two

struct Px8Params {
  const __nv_bfloat16* w;   // [KGB][Cout/8][8 rows][8 k] core-matrix packed; KGB = 10 (bf16) or 20 (x3: hi block, lo block)
  const float* bias;
  void* out;
  int B, H, W, Cout;
  int act, pool, split, x3;
  int tiles_x, tiles_y, total_tiles;
  int stages;
  int nacc, acc_cols;   // accumulator ring: nacc buffers of acc_cols TMEM columns
};

template <bool POOL, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_px8_kernel(const __grid_constant__ CUtensorMap tmA, const Px8Params p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[kMaxAcc], tempty_bar[kMaxAcc];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int KGB = p.x3 ? 2 * kAGroups : kAGroups;
  const uint32_t smemA = smem_base;                                       // stages x halo patch [18][10 px][16 B]
  const uint32_t smemB = smem_base + p.stages * kStageBytes;               // [KGB][Cout/8][8][16 B]
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  {  // weights (already in core-matrix order) and the all-zero tenth K group
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint8_t* gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    // the tenth K group of every stage: A = (1, 1, 1, 0, ...) per row against B = the bias split three ways (bf16 hi / mid / lo:
    // exact to 2^-25), so the accumulator already holds conv + bias and the epilogue neither loads nor adds it
    for (int s = 0; s < p.stages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(gen + s * kStageBytes + kPatchBytes);
      for (int i = tid; i < kZeroBytes / 16; i += kThreads) z[i] = make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u);
    }
    uint4* dstB = reinterpret_cast<uint4*>(gen + p.stages * kStageBytes);
    for (int i = tid; i < p.Cout * KGB; i += kThreads) dstB[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < p.nacc; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
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
      const int step = (int)gridDim.x;
      const int step_b = step / txy, step_m = step - step_b * txy;
      const int step_ty = step_m / p.tiles_x, step_tx = step_m - step_ty * p.tiles_x;
      int b = (int)blockIdx.x / txy, ty = ((int)blockIdx.x - b * txy) / p.tiles_x, tx = (int)blockIdx.x - b * txy - ty * p.tiles_x;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += step) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = full0 + 8 * stage;
        mbar_arrive_expect_tx(fb, (uint32_t)kPatchBytes);
        tma_load_3d(smemA + stage * kStageBytes, &tmA, fb, (tx * 8 - 1) * 8, ty * 16 - 1, b);   // inner coordinate in bf16 elements
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        tx += step_tx; ty += step_ty; b += step_b;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
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
      // tap t starts (t / 3) patch rows + (t % 3) pixels into the patch.  K = 16 = taps 2m and 2m + 1: the second K group starts
      // LBO bytes after the first; the tenth group is the stage's zero block.  Descriptors of stage 0, built once.
      auto tap_off = [](int t) { return (uint32_t)(((t / 3) * kPatchW + (t % 3)) * 16); };
      uint64_t ad0[kAGroups / 2];
#pragma unroll
      for (int m = 0; m < kAGroups / 2; ++m) {
        const uint32_t lbo = (2 * m + 1 < 9) ? tap_off(2 * m + 1) - tap_off(2 * m) : (uint32_t)kPatchBytes - tap_off(2 * m);
        ad0[m] = umma_desc(smemA + tap_off(2 * m), lbo, kRowBytes, 0);      // SBO: next 8 tile pixels = next patch row
      }
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * p.acc_cols;
        const uint64_t soff = (uint64_t)((stage * kStageBytes) >> 4);     // only the 14-bit start-address field moves
        uint64_t ad[kAGroups / 2];
#pragma unroll
        for (int m = 0; m < kAGroups / 2; ++m) ad[m] = ad0[m] + soff;
#pragma unroll
        for (int m = 0; m < kAGroups / 2; ++m) umma_bf16(d, ad[m], bdesc0 + m * bstep, idesc, m ? 1u : 0u);
        if (p.x3) {   // the same activations against the lo halves of the weights
#pragma unroll
          for (int m = 0; m < kAGroups / 2; ++m) umma_bf16(d, ad[m], bdesc0 + (kAGroups / 2 + m) * bstep, idesc, 1u);
        }
        umma_commit(empty0 + 8 * stage);
        umma_commit(tfull0 + 8 * acc);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        if (++acc == (uint32_t)p.nacc) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (TW = 8, TH = 16: pool partners lane^1, lane^8)
    const int ew = warp - 2, sub = warp & 3, grp = ew >> 2;      // four consecutive warps cover the four TMEM lane quadrants
    const int r = sub * 32 + lane;
    const int lx = r & 7, ly = r >> 3;
    const int PH = p.H >> 1, PW = p.W >> 1;
    const int pitch = SPLIT ? 2 * p.Cout : p.Cout;
    // this group's tiles: local tile index i = grp, grp + kEpiGroups, ...; accumulator = i % nacc (nacc is a multiple of kEpiGroups).
    // (b, ty, tx) advance incrementally: no division in the loop.
    const int step = kEpiGroups * (int)gridDim.x;
    const int step_b = step / txy, step_m = step - step_b * txy;
    const int step_ty = step_m / p.tiles_x, step_tx = step_m - step_ty * p.tiles_x;
    int tile = blockIdx.x + grp * gridDim.x;
    int b = tile / txy, ty = (tile - b * txy) / p.tiles_x, tx = tile - b * txy - ty * p.tiles_x;
    uint32_t acc = (uint32_t)grp, acc_phase = 0;
    __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (; tile < p.total_tiles; tile += step) {
      const int ox = tx * 8 + lx, oy = ty * 16 + ly;
      EpiOut eo;
      eo.vec_ok = true;   // Cout % 16 == 0
      size_t pix;
      if (POOL) {
        eo.valid = ((oy >> 1) < PH) && ((ox >> 1) < PW);
        pix = ((size_t)b * PH + (oy >> 1)) * PW + (ox >> 1);
      } else {
        eo.valid = (oy < p.H) && (ox < p.W);
        pix = ((size_t)b * p.H + oy) * p.W + ox;
      }
      __nv_bfloat16* const opix = outp + pix * pitch;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * p.acc_cols;
      for (int ch = 0; ch < p.Cout; ch += 32) {
        uint32_t regs[32];
        tmem_ld32(tcol + ch, regs);
        tmem_ld_wait();
        if (ch + 32 >= p.Cout) {   // the last chunk is in registers: hand the accumulator back before the math and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        }
        EpiOut e2 = eo;
        e2.c_left = p.Cout - ch;
        e2.ptr = opix + ch;
        if (SPLIT) epilogue_chunk_split<POOL, false>(regs, nullptr, p.act, lane, 8, e2, p.Cout);
        else epilogue_chunk<POOL, false, false>(regs, nullptr, p.act, lane, 8, e2);
      }
      acc += kEpiGroups;
      if (acc >= (uint32_t)p.nacc) { acc -= p.nacc; acc_phase ^= 1; }
      tx += step_tx; ty += step_ty; b += step_b;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
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
  // MEASURED (B200, 256 faces): 0.119 ms + 0.022 ms for the widening pass against 0.105 ms of tc_conv_first.cu, so this path is
  // opt-in (FLD_C1_PX8=1; read per call so that tests can switch it).  See the header and DESIGN.md for what was tried.
  const char* e = getenv("FLD_C1_PX8");
  const bool on = e && atoi(e) != 0;
  return on && g.kh == 3 && g.kw == 3 && g.Cin == 3 && g.stride == 1 && g.pad_t == 1 && g.pad_l == 1 && g.OH == g.IH && g.OW == g.IW &&
         g.Cout % 16 == 0 && g.Cout >= 16 && g.Cout <= 128 && (g.pool == 0 || g.pool == 2);   // <= 128: four accumulators of 128 columns
}

size_t tc_conv_px8_scratch_bytes(const ConvGeom& g, int B) { return (size_t)B * g.IH * g.IW * 16; }

// w_host fp32 [27][Cout] (k = (kh*3+kw)*3 + c)  ->  bf16 [KGB][Cout/8][8][8]: K group t (< 9) = tap t, element e (< 3) = channel e;
// group 9, elements 0..2 = the bias split three ways (multiplied by the ones the kernel keeps in the tenth A group);
// x3: groups 0..9 hold bf16(w), groups 10..19 the remainders bf16(w - bf16(w)) (group 19 zero).
static float px8_bf(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
void tc_conv_px8_pack(const float* w_host, const float* bias_host, int Cout, uint16_t (*f2bf)(float), uint16_t* out, int x3) {
  const int KGB = x3 ? 2 * kAGroups : kAGroups;
  for (int kg = 0; kg < KGB; ++kg)
    for (int ng = 0; ng < Cout / 8; ++ng)
      for (int r = 0; r < 8; ++r)
        for (int e = 0; e < 8; ++e) {
          const int t = kg % kAGroups, o = ng * 8 + r;
          float v = 0.f;
          if (t < 9 && e < 3) {
            const float w = w_host[(size_t)(t * 3 + e) * Cout + o];
            v = kg < kAGroups ? w : w - px8_bf(f2bf(w));
          } else if (kg == kAGroups - 1 && e < 3 && bias_host) {
            const float b0 = px8_bf(f2bf(bias_host[o])), b1 = px8_bf(f2bf(bias_host[o] - b0));
            v = e == 0 ? b0 : e == 1 ? b1 : bias_host[o] - b0 - b1;
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
  { const char* e = getenv("FLD_PX8_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= kMaxStages) p.stages = atoi(e); }
  p.acc_cols = g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : g.Cout <= 128 ? 128 : 256;
  p.nacc = std::min(kMaxAcc, 512 / p.acc_cols);
  if (p.nacc < kEpiGroups) { delete pl; fld_set_error("tc_conv_px8: Cout too large for the accumulator ring"); return FLD_ERR_INVALID; }
  p.nacc -= p.nacc % kEpiGroups;
  pl->smem = (size_t)p.stages * kStageBytes + bbytes + 1024;
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
  auto launch = [&](auto kern) -> int {
    FLD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    kern<<<pl->grid, kThreads, pl->smem, st>>>(pl->tmA, p);
    return FLD_OK;
  };
  const int rc = p.pool ? (p.split ? launch(conv_px8_kernel<true, true>) : launch(conv_px8_kernel<true, false>))
                        : (p.split ? launch(conv_px8_kernel<false, true>) : launch(conv_px8_kernel<false, false>));
  if (rc) return rc;
  FLD_LAUNCHED();
  return FLD_OK;
}
