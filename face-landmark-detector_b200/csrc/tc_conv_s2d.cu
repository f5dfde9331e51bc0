// tc_conv_s2d.cu — first conv stage (3x3, Cin = 3, pad 1, stride 1 + bias + ReLU + MaxPool 2x2) on tcgen05 with the POOL
// WINDOW IN THE TMEM COLUMNS (round 2).  Replaces stage 1 of vanilla_encoder (reference networks/fcn.py:25-31); the input is what
// prediction.py:82-84 / data/generator.py:53-61 hand the model.
//
// The first layer has K = 27: its MMAs are nothing, its cost is the epilogue.  In round 1's kernel (tc_conv_first.cu) a TMEM lane
// is a conv pixel, so the 2x2 max-pool crosses lanes: 24 SEL + 12 SHFL + 16 HMNMX2 per 32-column chunk on top of the per-thread
// operand building — ~1900 warp instructions per 128-pixel tile, issue-bound at 0.105 ms per 256 faces against a TMEM-read floor of
// 0.059 ms.  (An intermediate round-2 kernel that only replaced the operand building by TMA — same lane = conv pixel epilogue —
// was measured at 0.118 ms + 0.022 ms for the widening pass: the shuffle epilogue alone is ~2000 warp instructions per tile and
// the epilogue warps starved on `tfull`; it was dropped.)  Here a TMEM lane is a POOLED pixel and the four conv pixels of
// its window are four column blocks of the same lane:
//   D[pooled px][pos * Cout + o],  pos = (dy, dx) in {0,1}^2,
// so the pool is three in-register FMNMX per output value, no shuffles, no selects, and a thread stores 64 contiguous bytes.
//
// Operand feed: the image is rewritten once as four space-to-depth planes S[y & 1][x & 1][B][H/2][W/2][8 bf16] (RGB + 5 zeros =
// 16 B per pixel).  The conv pixel of position (dy, dx) of pooled pixel (py, px) reads source (2 py + a, 2 px + b) with
// a = dy + ky - 1, b = dx + kx - 1 in {-1..2}: plane (a & 1, b & 1), pixel (py + (a >> 1), px + (b >> 1)) — contiguous in px, so a
// K group (8 elements = one tap of one pixel) is again a window of a TMA-loaded halo patch: per tile FOUR boxes (one per plane,
// 18 rows x 10 pixels x 16 B, out-of-bounds fill = the conv's zero padding) serve all 4 x 9 (position, tap) pairs through UMMA
// descriptors in the no-swizzle K-major layout (SBO = one patch row; a K = 16 MMA spans two taps whose starts differ by the
// descriptor's LBO — the host orders each pair by address and packs the weights in the same order).  The tenth K group of every
// position is a block of ones against the bias split three ways (bf16 hi / mid / lo), so the accumulators hold conv + bias.
//
// Roles (320 threads, one persistent CTA per SM): warp 0 TMA producer (stage ring), warp 1 MMA issuer (4 positions x 5 MMAs of
// M = 128, N = Cout, K = 16 per tile, two accumulators of 4 * Cout columns), warps 2-9 epilogue in two groups of four warps (one
// per TMEM lane quadrant) that take tiles alternately.  FLD_BF16X3 (uint8 input, exact in bf16): weights split hi / lo, the
// same activation groups multiplied a second time against the lo block; SPLIT output.
// MEASURED (B200, 256 faces @128x128, ncu): conv_s2d_kernel 58 us = the TMEM-read floor; the widening pass costs 21 us unless the
// producer of the crops writes the planes itself (fld_preprocess_faces_staged: then the layer is 0.058 ms, round 1: 0.105 ms).
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kPatchW = 10, kPatchH = 18;            // halo patch of an 8 x 16 pooled-pixel tile in one plane, 16 B per pixel
constexpr int kRowBytes = kPatchW * 16;              // 160
constexpr int kPatchBytes = kPatchH * kRowBytes;     // 2880 (what one TMA load delivers)
constexpr int kPatchPitch = (kPatchBytes + 127) / 128 * 128;   // 2944: TMA destinations are 128-byte aligned
constexpr int kOnesBytes = 16 * kRowBytes;           // ones K group: 16 core matrices at the 160-byte stride
constexpr int kStageBytes = (4 * kPatchPitch + kOnesBytes + 127) / 128 * 128;   // 14336
constexpr int kPairs = 5;                            // K = 16 MMAs per position: (t0,t1) (t2,t3) (t4,t5) (t6,t7) (t8, ones)
constexpr int kEpiWarps = 8, kEpiGroups = 2;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;

struct S2dParams {
  const __nv_bfloat16* w;   // [KGB = 40 (x3: 80)][Cout/8][8 rows][8 k] core-matrix packed, group = (pos * 5 + pair) * 2 + {first, second}
  void* out;
  int B, PH, PW, Cout;      // pooled output size
  int act, x3;
  int tiles_x, tiles_y, total_tiles;
  int stages;
  uint32_t a_off[4 * kPairs], a_lbo[4 * kPairs];   // per (pos, pair): byte offset of the first K group inside a stage, distance to the second
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_s2d_kernel(const __grid_constant__ CUtensorMap tmA, const S2dParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int KGB = (p.x3 ? 2 : 1) * 4 * kPairs * 2;
  const uint32_t smemA = smem_base;                                       // stages x (4 plane patches + ones block)
  const uint32_t smemB = smem_base + p.stages * kStageBytes;               // [KGB][Cout/8][8][16 B]
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
  const uint32_t acc_cols = 4u * p.Cout;                                   // 2 accumulators: 8 * Cout <= 512

  {  // weights (already in core-matrix order) and the ones K group of every stage: (1, 1, 1, 0, ...) per row
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint8_t* gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    for (int s = 0; s < p.stages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(gen + s * kStageBytes + 4 * kPatchPitch);
      for (int i = tid; i < kOnesBytes / 16; i += kThreads) z[i] = make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u);
    }
    uint4* dstB = reinterpret_cast<uint4*>(gen + p.stages * kStageBytes);
    for (int i = tid; i < p.Cout * KGB; i += kThreads) dstB[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
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
    // ------------------------------------------------------------------ TMA producer: four plane patches per tile
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const int step = (int)gridDim.x;
      const int step_b = step / txy, step_m = step - step_b * txy;
      const int step_ty = step_m / p.tiles_x, step_tx = step_m - step_ty * p.tiles_x;
      int b = (int)blockIdx.x / txy, ty = ((int)blockIdx.x - b * txy) / p.tiles_x, tx = (int)blockIdx.x - b * txy - ty * p.tiles_x;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += step) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = full0 + 8 * stage;
        mbar_arrive_expect_tx(fb, 4u * kPatchBytes);
        const uint32_t sa = smemA + stage * kStageBytes;
#pragma unroll
        for (int q = 0; q < 4; ++q)   // plane q = (y & 1) * 2 + (x & 1); inner coordinate in bf16 elements
          tma_load_4d(sa + q * kPatchPitch, &tmA, fb, (tx * 8 - 1) * 8, ty * 16 - 1, b, q);
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
      uint64_t ad0[4 * kPairs];                                            // A descriptors of stage 0
#pragma unroll
      for (int i = 0; i < 4 * kPairs; ++i) ad0[i] = umma_desc(smemA + p.a_off[i], p.a_lbo[i], kRowBytes, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint64_t soff = (uint64_t)((stage * kStageBytes) >> 4);     // only the 14-bit start-address field moves
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
          const uint32_t d = tmem_base + acc * acc_cols + pos * p.Cout;
#pragma unroll
          for (int m = 0; m < kPairs; ++m) umma_bf16(d, ad0[pos * kPairs + m] + soff, bdesc0 + (pos * kPairs + m) * bstep, idesc, m ? 1u : 0u);
          if (p.x3) {   // the same activations against the lo halves of the weights
#pragma unroll
            for (int m = 0; m < kPairs; ++m)
              umma_bf16(d, ad0[pos * kPairs + m] + soff, bdesc0 + (4 * kPairs + pos * kPairs + m) * bstep, idesc, 1u);
          }
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
    // ------------------------------------------------------------------ epilogue: lane = pooled pixel, pool over the four column blocks
    const int ew = warp - 2, sub = warp & 3, grp = ew >> 2;      // four consecutive warps cover the four TMEM lane quadrants
    const int r = sub * 32 + lane;
    const int lx = r & 7, ly = r >> 3;
    const int pitch = SPLIT ? 2 * p.Cout : p.Cout;
    const int step = kEpiGroups * (int)gridDim.x;
    const int step_b = step / txy, step_m = step - step_b * txy;
    const int step_ty = step_m / p.tiles_x, step_tx = step_m - step_ty * p.tiles_x;
    int tile = blockIdx.x + grp * gridDim.x;
    int b = tile / txy, ty = (tile - b * txy) / p.tiles_x, tx = tile - b * txy - ty * p.tiles_x;
    const uint32_t acc = (uint32_t)grp;                           // nacc == kEpiGroups: group g always uses accumulator g
    uint32_t acc_phase = 0;
    __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (; tile < p.total_tiles; tile += step) {
      const int px = tx * 8 + lx, py = ty * 16 + ly;
      const bool valid = (py < p.PH) && (px < p.PW);
      __nv_bfloat16* const opix = outp + (((size_t)b * p.PH + py) * p.PW + px) * pitch;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * acc_cols;
      for (int ch = 0; ch < p.Cout; ch += 16) {
        uint32_t q0[16], q1[16], q2[16], q3[16];
        tmem_ld16(tcol + ch, q0);
        tmem_ld16(tcol + p.Cout + ch, q1);
        tmem_ld16(tcol + 2 * p.Cout + ch, q2);
        tmem_ld16(tcol + 3 * p.Cout + ch, q3);
        tmem_ld_wait();
        if (ch + 16 >= p.Cout) {   // the last chunk is in registers: hand the accumulator back before the math and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = act_f(fmaxf(fmaxf(__uint_as_float(q0[j]), __uint_as_float(q1[j])), fmaxf(__uint_as_float(q2[j]), __uint_as_float(q3[j]))), p.act);
        if (valid) {
          if (SPLIT) {
            split_store8(opix + ch, p.Cout, v);
            split_store8(opix + ch + 8, p.Cout, v + 8);
          } else {
            uint4 u0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            uint4 u1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
            *reinterpret_cast<uint4*>(opix + ch) = u0;
            *reinterpret_cast<uint4*>(opix + ch + 8) = u1;
          }
        }
      }
      acc_phase ^= 1;
      tx += step_tx; ty += step_ty; b += step_b;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// [B][H][W][3] uint8 / float32  ->  four planes [(y & 1) * 2 + (x & 1)][B][H/2][W/2][8] bf16 (channels 3..7 zero).
// A thread takes 4 consecutive pixels of a row: two 32-byte stores, one per x-parity plane.
template <typename TIn>
__global__ void s2d_widen_kernel(const TIn* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int aligned4) {
  const int W4 = W >> 2;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, y, x / 4)
  if (g >= (long long)B * H * W4) return;
  const int x4 = (int)(g % W4);
  const long long by = g / W4;
  const int y = (int)(by % H), b = (int)(by / H);
  const TIn* src = in + ((size_t)by * W + (size_t)x4 * 4) * 3;
  float v[12];
  if (sizeof(TIn) == 1 && aligned4) {                                 // 12 bytes = three aligned words (W % 4 == 0)
    const uint32_t* q = reinterpret_cast<const uint32_t*>(src);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (float)((w0 >> (8 * i)) & 255u); v[4 + i] = (float)((w1 >> (8 * i)) & 255u); v[8 + i] = (float)((w2 >> (8 * i)) & 255u); }
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = (float)__ldg(src + i);
  }
  const int H2 = H >> 1, W2 = W >> 1;
  const size_t plane = (size_t)B * H2 * W2;
  uint4* o = out + (size_t)((y & 1) * 2) * plane + ((size_t)b * H2 + (y >> 1)) * W2 + (size_t)x4 * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i)       // pixel x4*4 + i: parity i & 1, plane column x4*2 + (i >> 1)
    o[(size_t)(i & 1) * plane + (i >> 1)] = make_uint4(pack_bf16(v[3 * i], v[3 * i + 1]), pack_bf16(v[3 * i + 2], 0.f), 0u, 0u);
}

// (a, b) = (dy + ky - 1, dx + kx - 1): byte offset of that K group inside a stage
int group_off(int a, int b) {
  const int q = (a & 1) * 2 + (b & 1);
  return q * kPatchPitch + ((a >> 1) + 1) * kRowBytes + ((b >> 1) + 1) * 16;    // >> on negatives: arithmetic (a = -1 -> -1)
}

}  // namespace

struct TcS2dPlan {
  CUtensorMap tmA;
  S2dParams p;
  int grid;
  size_t smem;
  void* scratch;
  int in_dtype, H, W;
};

bool tc_conv_s2d_supported(const ConvGeom& g) {
  const char* e = getenv("FLD_C1_S2D");     // read per call so that tests can switch it
  const bool on = !(e && atoi(e) == 0);
  return on && g.kh == 3 && g.kw == 3 && g.Cin == 3 && g.stride == 1 && g.pad_t == 1 && g.pad_l == 1 && g.OH == g.IH && g.OW == g.IW &&
         g.pool == 2 && g.IH % 2 == 0 && g.IW % 4 == 0 && g.Cout % 16 == 0 && g.Cout >= 16 && g.Cout <= 64;   // 8 * Cout TMEM columns
}

size_t tc_conv_s2d_scratch_bytes(const ConvGeom& g, int B) { return (size_t)B * g.IH * g.IW * 16; }

// Order of the K groups of position pos, pair m: the two taps (or tap 8 + the ones block) sorted by their address in a stage.
static void pair_groups(int pos, int m, int* first_tap, int* second_tap, uint32_t* off, uint32_t* lbo) {
  const int dy = pos >> 1, dx = pos & 1;
  auto tap_off = [&](int t) { return t < 9 ? group_off(dy + t / 3 - 1, dx + t % 3 - 1) : 4 * kPatchPitch; };   // t == 9: ones block
  int t0 = 2 * m, t1 = 2 * m + 1;
  if (tap_off(t1) < tap_off(t0)) { const int s = t0; t0 = t1; t1 = s; }
  *first_tap = t0; *second_tap = t1;
  *off = (uint32_t)tap_off(t0);
  *lbo = (uint32_t)(tap_off(t1) - tap_off(t0));
}

// w_host fp32 [27][Cout] (k = (kh*3+kw)*3 + c), bias [Cout]  ->  bf16 [KGB][Cout/8][8][8]: group (pos * 5 + m) * 2 + {0, 1} = the
// pair's first / second tap (element e < 3 = channel e); tap index 9 = the bias split three ways; x3: a second block of 40 groups
// with the remainders bf16(w - bf16(w)) (its bias slots zero).
static float s2d_bf(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
void tc_conv_s2d_pack(const float* w_host, const float* bias_host, int Cout, uint16_t (*f2bf)(float), uint16_t* out, int x3) {
  const int half = 4 * kPairs * 2;
  for (int blk = 0; blk < (x3 ? 2 : 1); ++blk)
    for (int pos = 0; pos < 4; ++pos)
      for (int m = 0; m < kPairs; ++m) {
        int taps[2];
        uint32_t off, lbo;
        pair_groups(pos, m, &taps[0], &taps[1], &off, &lbo);
        for (int s = 0; s < 2; ++s) {
          const int kg = blk * half + (pos * kPairs + m) * 2 + s, t = taps[s];
          for (int ng = 0; ng < Cout / 8; ++ng)
            for (int r = 0; r < 8; ++r)
              for (int e = 0; e < 8; ++e) {
                const int o = ng * 8 + r;
                float v = 0.f;
                if (t < 9 && e < 3) {
                  const float w = w_host[(size_t)(t * 3 + e) * Cout + o];
                  v = blk == 0 ? w : w - s2d_bf(f2bf(w));
                } else if (t == 9 && e < 3 && blk == 0 && bias_host) {
                  const float b0 = s2d_bf(f2bf(bias_host[o])), b1 = s2d_bf(f2bf(bias_host[o] - b0));
                  v = e == 0 ? b0 : e == 1 ? b1 : bias_host[o] - b0 - b1;
                }
                out[(((size_t)kg * (Cout / 8) + ng) * 8 + r) * 8 + e] = f2bf(v);
              }
        }
      }
}

int tc_conv_s2d_plan_create(const fld_handle* h, void* scratch, int in_dtype, const ConvGeom& g, int B, int x3, int split_out,
                            TcS2dPlan** out) {
  if (!h->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  if (x3 && in_dtype != FLD_U8) { fld_set_error("tc_conv_s2d: the FLD_BF16X3 variant takes a uint8 input"); return FLD_ERR_INVALID; }
  if ((x3 != 0) != (split_out != 0)) { fld_set_error("tc_conv_s2d: SPLIT output goes with FLD_BF16X3"); return FLD_ERR_INVALID; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  TcS2dPlan* pl = new TcS2dPlan();
  S2dParams& p = pl->p;
  p.w = nullptr; p.out = nullptr;
  p.B = B; p.PH = g.IH / 2; p.PW = g.IW / 2; p.Cout = g.Cout; p.act = g.act; p.x3 = x3;
  p.tiles_x = fld_div_up(p.PW, 8); p.tiles_y = fld_div_up(p.PH, 16);
  p.total_tiles = B * p.tiles_x * p.tiles_y;
  for (int pos = 0; pos < 4; ++pos)
    for (int m = 0; m < kPairs; ++m) {
      int t0, t1;
      pair_groups(pos, m, &t0, &t1, &p.a_off[pos * kPairs + m], &p.a_lbo[pos * kPairs + m]);
    }
  const size_t bbytes = (size_t)g.Cout * 16 * (x3 ? 2 : 1) * 4 * kPairs * 2;
  p.stages = (int)std::max<size_t>(2, std::min<size_t>(kMaxStages, (200 * 1024 - bbytes) / kStageBytes));
  pl->smem = (size_t)p.stages * kStageBytes + bbytes + 1024;
  pl->grid = std::min(p.total_tiles, h->sm_count);
  pl->scratch = scratch; pl->in_dtype = in_dtype; pl->H = g.IH; pl->W = g.IW;
  // four planes of [B][H/2][W/2] pixels x 8 bf16: a box row is 10 pixels = 160 contiguous bytes
  const cuuint64_t H2 = g.IH / 2, W2 = g.IW / 2;
  cuuint64_t dims[4] = {W2 * 8, H2, (cuuint64_t)B, 4};
  cuuint64_t strides[3] = {W2 * 16, H2 * W2 * 16, (cuuint64_t)B * H2 * W2 * 16};
  cuuint32_t box[4] = {(cuuint32_t)kPatchW * 8, (cuuint32_t)kPatchH, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, scratch, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(s2d A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  *out = pl;
  return FLD_OK;
}

void tc_conv_s2d_plan_destroy(TcS2dPlan* p) { delete p; }

int tc_conv_s2d_run(const TcS2dPlan* pl, const void* in, const __nv_bfloat16* w_packed, void* out, cudaStream_t st, int staged) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  const long long n = (long long)pl->p.B * pl->H * (pl->W / 4);
  const long long blocks = (n + 255) / 256;
  if (blocks >= (1ll << 31)) { fld_set_error("tc_conv_s2d: too many pixels"); return FLD_ERR_INVALID; }
  const int al = (reinterpret_cast<uintptr_t>(in) & 3) == 0;
  if (staged) {
    // the producer of `in` (fld_preprocess_faces_staged) has already written the space-to-depth planes
  } else if (pl->in_dtype == FLD_U8) s2d_widen_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)in, (uint4*)pl->scratch, pl->p.B, pl->H, pl->W, al);
  else if (pl->in_dtype == FLD_F32) s2d_widen_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, (uint4*)pl->scratch, pl->p.B, pl->H, pl->W, al);
  else { fld_set_error("tc_conv_s2d: input must be u8 or f32"); return FLD_ERR_INVALID; }
  if (!staged) FLD_LAUNCHED();
  S2dParams p = pl->p;
  p.w = w_packed; p.out = out;
  if (p.x3) {
    FLD_CUDA(cudaFuncSetAttribute(conv_s2d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_s2d_kernel<true><<<pl->grid, kThreads, pl->smem, st>>>(pl->tmA, p);
  } else {
    FLD_CUDA(cudaFuncSetAttribute(conv_s2d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_s2d_kernel<false><<<pl->grid, kThreads, pl->smem, st>>>(pl->tmA, p);
  }
  FLD_LAUNCHED();
  return FLD_OK;
}
