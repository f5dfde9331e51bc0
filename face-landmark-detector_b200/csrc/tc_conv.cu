// tc_conv.cu — generic tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (FLD_BF16 mode).
//
// Replaces the TensorFlow conv arithmetic behind reference networks/fcn.py:33-49 (ZeroPad -> Conv3x3 -> BN -> ReLU
// -> MaxPool2 stages 2..5), fcn.py:98-103,108,117 (7x7 / 1x1 head and score convs) and networks/vgg16.py:27-73.
// BN is folded into weights/bias on the host (net.cu); bias + ReLU + 2x2 max-pool are fused into the epilogue so
// a stage's pre-pool map never touches HBM.
//
// GEMM view: D[M = output pixels][N = Cout] = A[M][K = taps*Cin] * B[N][K]^T, bf16 operands, fp32 accumulation
// in TMEM.  An M tile is a TW x TH x NB pixel patch (128 pixels) so that
//   (1) one 4-D TMA box per (tap, 64-channel chunk) fetches the shifted input patch, with TMA's out-of-bounds
//       zero fill supplying the convolution's zero padding, and
//   (2) the four pixels of every 2x2 pool window sit in one warp's 32 TMEM lanes (lane^1, lane^TW): pooling is
//       two rounds of warp shuffles on packed bf16x2 (tc_common.cuh).
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      TMA producer        multi-stage mbarrier ring (A box + B box per k-block of 64 channels)
//   warp 1      MMA issuer          4 x tcgen05.mma (128 x BN x 16) per k-block, tcgen05.commit frees the stage
//   warps 2..9  epilogue            two warps per TMEM lane quadrant, alternating 32-column chunks; double-buffered
//                                   accumulators (2 x 256 TMEM columns) overlap tile i's epilogue with tile i+1's MMAs
// Each of the two issue roles runs its whole loop inside ONE `elect.sync`-elected lane (the compiler then knows
// the region is single-lane and emits plain UTMALDG / UTCHMMA / UTCBAR).  Measured with the clock64 trace below (FLD_TC_TRACE): the first version ran the loops
// inside `if (lane == 0)`, which made nvcc wrap every UTMALDG / UTCHMMA / UTCBAR in R2UR + ELECT waterfall loops —
// ~730 cycles per k-block in BOTH roles, i.e. 70 % tensor-active at N = 256 and 31 % at N = 128.

#include <stdlib.h>
#include <vector>
#include "tc_common.cuh"

#include <algorithm>
#include <utility>
#include <vector>

namespace {
using namespace tc;

struct TmaConvParams {
  const float* bias;   // [Cout_pad + 32]
  void* out;
  int B, OH, OW, Cout; // conv output (pre-pool) dims
  int Cin, kh, kw, pad_t, pad_l;
  int flat;            // 1x1 / stride 1 / no padding / no pool: the M tile is 128 CONSECUTIVE pixels of the flattened [B*H*W] list
  long long npx;       //   (2-D tensor map, no per-image tile padding: a 7x7 map wastes 23 % of an 8x8 tile otherwise)
  int IH, IW;          // input rows / columns (tap skipping)
  const int* sched;    // optional [n_iter]: CTA c runs tiles sched[c], sched[c + grid], ... (-1 = none).  Tap skipping makes tile
  int n_iter;          // costs unequal; the host balances them (longest-processing-time first).  Without it n_iter = total_tiles
  int st;              // conv stride (1, or 2 for 1x1 convs: the tensor map's element strides pick every second pixel)
  int TW, TH, NB;      // M tile = TW*TH*NB = 128 pixels
  int BN, cout_pad, n_ntiles;
  int tiles_x, tiles_y, tiles_b, total_tiles;
  int stages;
  int act, pool;
  int a_chunks;        // FLD_BF16X3: k-chunk kc reads activation chunk kc % a_chunks ([x_hi | x_lo] against [w_hi | w_hi | w_lo])
  int split;           // bf16 output stored as a SPLIT tensor ([hi | lo], pixel pitch 2 * Cout)
  int ksplit;          // split-K: tile = (k slice, N tile, M tile); slice s accumulates k-chunks [s*kchunks/ksplit, ...) into its own
                       // fp32 partial output [ksplit][npx][Cout] (flat mode only; a fixed-order reduce kernel sums them)
  unsigned long long* trace;  // FLD_TC_TRACE: clock64 event log of CTA 0, [3 roles][kTraceN]
  int dbg;             // FLD_TC_DBG bisect switches (results are garbage when set): 1 skip epilogue math/stores,
                       // 2 skip TMEM loads too, 4 skip the A-operand TMA, 8 skip the B-operand TMA, 16 skip the MMAs
};

constexpr int kEpiWarps = 8;
constexpr int kTmaThreads = 64 + 32 * kEpiWarps;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kMaxStages = 8;
constexpr int kTraceN = 2048;
#define TRACE(role, idx, tag)                                                                                         \
  do {                                                                                                                \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (idx) < kTraceN)                                                   \
      p.trace[(role) * kTraceN + (idx)++] = ((unsigned long long)clock64() << 4) | (tag);                             \
  } while (0)

// Kernel rows whose input rows fall into the zero padding for EVERY output row of the tile contribute nothing and are
// skipped; producer and MMA issuer derive the same range.  With one output row per tile (small maps under the 7x7 'same'
// head, fcn.py:98) this removes 12 of 49 taps on a 7x7 map.
__device__ __forceinline__ void ky_range(const TmaConvParams& p, int ty, int& lo, int& hi) {
  const int r0 = ty * p.TH * p.st, r1 = r0 + (p.TH - 1) * p.st;
  lo = max(0, p.pad_t - r1);
  hi = min(p.kh - 1, p.IH - 1 + p.pad_t - r0);
  if (hi < lo) { lo = 0; hi = p.kh - 1; }
}
__device__ __forceinline__ void kx_range(const TmaConvParams& p, int tx, int& lo, int& hi) {
  const int c0 = tx * p.TW * p.st, c1 = c0 + (p.TW - 1) * p.st;
  lo = max(0, p.pad_l - c1);
  hi = min(p.kw - 1, p.IW - 1 + p.pad_l - c0);
  if (hi < lo) { lo = 0; hi = p.kw - 1; }
}

template <bool OUT_F32>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TmaConvParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;  // BN multiple of 16 -> b_bytes multiple of 2048 -> 1024-aligned stages
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiWarps); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kchunks = p.Cin >> 6;
  const int kcs = kchunks / p.ksplit;      // k-chunks per slice
  const int mtiles = p.tiles_x * p.tiles_y * p.tiles_b;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: one elected lane runs the loop
    if (elect_one()) {
    uint32_t stage = 0, phase = 0;
    int ti = 0;
    for (int it = blockIdx.x; it < p.n_iter; it += gridDim.x) {
      const int tile = p.sched ? __ldg(p.sched + it) : it;
      if (tile < 0) continue;
      const int ks = tile / (mtiles * p.n_ntiles);   // k slice (0 unless split-K)
      const int tile_mn = tile - ks * (mtiles * p.n_ntiles);
      const int nt = tile_mn / mtiles;             // N tile is the slow index: concurrent CTAs share the weight tile
      int m = tile_mn - nt * mtiles;
      const int tb = m / (p.tiles_x * p.tiles_y);
      m -= tb * (p.tiles_x * p.tiles_y);
      const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
      const int x0 = tx * p.TW * p.st - p.pad_l, y0 = ty * p.TH * p.st - p.pad_t, b0 = tb * p.NB, n0 = nt * p.BN;
      int ky_lo, ky_hi, kx_lo, kx_hi;
      ky_range(p, ty, ky_lo, ky_hi);
      kx_range(p, tx, kx_lo, kx_hi);
      const int kc_lo = ks * kcs, kc_hi = kc_lo + kcs;
      for (int ky = ky_lo; ky <= ky_hi; ++ky) {
        for (int kx = kx_lo; kx <= kx_hi; ++kx) {
          const int wrow = n0 + (ky * p.kw + kx) * p.cout_pad;  // row of the [taps*cout_pad][Cin] weight matrix
          for (int kc = kc_lo; kc < kc_hi; ++kc) {
            const int ka = kc >= p.a_chunks ? kc - p.a_chunks : kc;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            TRACE(0, ti, 1);
            {
              const uint32_t fb = full0 + 8 * stage;
              const uint32_t sa = smem_base + stage * stage_bytes;
              const uint32_t bytes = ((p.dbg & 4) ? 0u : a_bytes) + ((p.dbg & 8) ? 0u : b_bytes);
              TRACE(0, ti, 5);
              mbar_arrive_expect_tx(fb, bytes);
              TRACE(0, ti, 6);
              if (!(p.dbg & 4)) {
                if (p.flat) tma_load_2d(sa, &tmA, fb, ka * 64, tb * 128);     // rows past npx are zero-filled
                else tma_load_4d(sa, &tmA, fb, ka * 64, x0 + kx, y0 + ky, b0);
              }
              TRACE(0, ti, 7);
              if (!(p.dbg & 8)) tma_load_2d(sa + a_bytes, &tmB, fb, kc * 64, wrow);
              TRACE(0, ti, 8);
            }
            TRACE(0, ti, 2);
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one elected lane runs the loop
    if (elect_one()) {
    // Descriptors are built once; per stage / per 16-element K step only the 14-bit start-address field moves
    // (units of 16 bytes).
    const uint32_t idesc = umma_idesc_bf16(128, p.BN);
    const uint64_t adesc0 = umma_desc(smem_base, 16, 1024, 2);
    const uint64_t bdesc0 = umma_desc(smem_base + a_bytes, 16, 1024, 2);
    const uint32_t stage_step = stage_bytes >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    int ti = 0;
    for (int it = blockIdx.x; it < p.n_iter; it += gridDim.x) {
      const int tile = p.sched ? __ldg(p.sched + it) : it;
      if (tile < 0) continue;
      TRACE(1, ti, 0);
      int kblocks_t;
      {
        const int m = tile % mtiles;               // tile = (ks * n_ntiles + nt) * mtiles + m
        const int mm = m % (p.tiles_x * p.tiles_y);
        const int ty = mm / p.tiles_x, tx = mm - ty * p.tiles_x;
        int ky_lo, ky_hi, kx_lo, kx_hi;
        ky_range(p, ty, ky_lo, ky_hi);
        kx_range(p, tx, kx_lo, kx_hi);
        kblocks_t = (ky_hi - ky_lo + 1) * (kx_hi - kx_lo + 1) * kcs;
      }
      mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
      TRACE(1, ti, 3);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256;
      uint32_t accum = 0;
      for (int kb = 0; kb < kblocks_t; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        TRACE(1, ti, 1);
        tc_fence_after();
        TRACE(1, ti, 5);
        {
          TRACE(1, ti, 6);
          const uint64_t ad = adesc0 + (uint64_t)(stage * stage_step);
          const uint64_t bd = bdesc0 + (uint64_t)(stage * stage_step);
          if (!(p.dbg & 16)) {
            umma_bf16(d, ad, bd, idesc, accum);
            umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
            umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
            umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
          }
          TRACE(1, ti, 7);
          umma_commit(empty0 + 8 * stage);                       // stage reusable once these MMAs have read it
          if (kb == kblocks_t - 1) umma_commit(tfull0 + 8 * acc);  // accumulator complete
          TRACE(1, ti, 8);
        }
        accum = 1u;
        TRACE(1, ti, 2);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    const int ew = warp - 2;
    const int sub = warp & 3;            // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int half = ew >> 2;            // the quadrant's two warps take alternating 32-column chunks
    const int r = sub * 32 + lane;       // tile row = pixel index inside the M tile
    const int lx = r % p.TW, ly = (r / p.TW) % p.TH, nb = r / (p.TW * p.TH);
    const int PH = p.OH >> 1, PW = p.OW >> 1;
    uint32_t acc = 0, acc_phase = 0;
    int ti = 0;
    const bool tr = (warp == 2);
    for (int it = blockIdx.x; it < p.n_iter; it += gridDim.x) {
      const int tile = p.sched ? __ldg(p.sched + it) : it;
      if (tile < 0) continue;
      const int ks = tile / (mtiles * p.n_ntiles);
      const int tile_mn = tile - ks * (mtiles * p.n_ntiles);
      const int nt = tile_mn / mtiles;
      int m = tile_mn - nt * mtiles;
      const int tb = m / (p.tiles_x * p.tiles_y);
      m -= tb * (p.tiles_x * p.tiles_y);
      const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
      const int ox = tx * p.TW + lx, oy = ty * p.TH + ly, b = tb * p.NB + nb, n0 = nt * p.BN;
      EpiOut eo;
      eo.vec_ok = OUT_F32 ? (p.Cout % 4 == 0) : (p.Cout % 8 == 0);
      size_t pix;
      if (p.pool) {
        eo.valid = (b < p.B) && ((oy >> 1) < PH) && ((ox >> 1) < PW);
        pix = ((size_t)b * PH + (oy >> 1)) * PW + (ox >> 1);
      } else if (p.flat) {
        pix = (size_t)tb * 128 + r;
        eo.valid = (long long)pix < p.npx;
      } else {
        eo.valid = (b < p.B) && (oy < p.OH) && (ox < p.OW);
        pix = ((size_t)b * p.OH + oy) * p.OW + ox;
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      if (tr) TRACE(2, ti, 1);
      tc_fence_after();
      for (int ch = half * 32; ch < p.BN; ch += 64) {
        if (p.dbg & 2) break;
        uint32_t regs[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + ch, regs);
        tmem_ld_wait();
        if (p.dbg & 1) { if ((regs[0] ^ regs[31]) == 0x7fc12345u) eo.valid = false; continue; }
        EpiOut e2 = eo;
        e2.c_left = p.Cout - n0 - ch;
        if (!OUT_F32 && p.split) {
          e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * (2 * p.Cout) + n0 + ch;
          if (p.pool) epilogue_chunk_split<true>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2, p.Cout);
          else epilogue_chunk_split<false>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2, p.Cout);
          continue;
        }
        if (OUT_F32) e2.ptr = reinterpret_cast<float*>(p.out) + ((size_t)ks * p.npx + pix) * p.Cout + n0 + ch;
        else e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + n0 + ch;
        if (p.pool) epilogue_chunk<true, false>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2);
        else epilogue_chunk<false, OUT_F32>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (tr) TRACE(2, ti, 2);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


}  // namespace

struct TcConvPlan {
  int* d_perm = nullptr;
  CUtensorMap tmA, tmB;
  TmaConvParams p;
  int grid;
  size_t smem;
};

bool tc_conv_supported(const ConvGeom& g) {
  if (g.Cin % 64 != 0 || g.Cin < 64) return false;
  // stride 2 only for 1x1 / no padding (ResNet50's down-sampling convs, resnet50.py:98,110): the TMA box itself strides
  if (g.stride != 1 && !(g.stride == 2 && g.kh == 1 && g.kw == 1 && g.pad_t == 0 && g.pad_l == 0 && g.pool == 0)) return false;
  if (g.pool != 0 && g.pool != 2) return false;
  if (g.pool == 2 && (g.Cout % 8 != 0)) return false;
  if (g.kh * g.kw > 64) return false;
  return true;
}

int tc_conv_plan_create(const fld_handle* h, const void* in, const __nv_bfloat16* w_packed, int cout_pad, const ConvGeom& g, int B,
                        TcConvPlan** out, int x3, int split_out, int ksplit) {
  if (!h->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  TcConvPlan* pl = new TcConvPlan();
  TmaConvParams& p = pl->p;
  p.bias = nullptr; p.out = nullptr;
  // x3: the K dimension is [x_hi w_hi | x_lo w_hi | x_hi w_lo] = 3 Cin channels over a 2 Cin-channel activation tensor
  const int Kc = x3 ? 3 * g.Cin : g.Cin, Ac = x3 ? 2 * g.Cin : g.Cin;
  p.B = B; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout;
  p.Cin = Kc; p.kh = g.kh; p.kw = g.kw; p.pad_t = g.pad_t; p.pad_l = g.pad_l;
  p.a_chunks = Ac / 64; p.split = split_out; p.ksplit = 1;
  p.st = g.stride; p.IH = g.IH; p.IW = g.IW;
  p.act = g.act; p.pool = g.pool;
  { const char* e = getenv("FLD_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.trace = nullptr;
  // M-tile geometry: TW*TH*NB = 128 pixels, TW in {4, 8}: the pool partners are lane^1 and lane^TW
  int TW = g.OW > 4 ? 8 : 4;
  // one output row per tile under a tall padded kernel on a small map: most kernel rows of the border rows multiply padding
  const bool row_tiles = g.pool == 0 && g.kh >= 5 && g.pad_t > 0 && g.OH <= 16 && !getenv("FLD_TC_ROWTILES_OFF");
  int TH = row_tiles ? 1 : 2;
  while (!row_tiles && TH * 2 <= 128 / TW && TH < g.OH) TH *= 2;
  // with >= 96 images a tile can be ONE output pixel of 128 images: kernel columns that only see padding are skipped too
  // (7x7 'same' on a 7x7 map: 57 % of the taps remain instead of 76 %)
  if (row_tiles && g.kw >= 5 && g.pad_l > 0 && B >= 96 && !getenv("FLD_TC_PIXTILES_OFF")) TW = 1;
  const int NB = 128 / (TW * TH);
  p.TW = TW; p.TH = TH; p.NB = NB;
  // N tile
  const int BN = cout_pad <= 256 ? cout_pad : ((cout_pad % 256 == 0) ? 256 : 128);
  if (cout_pad % BN != 0 || BN % 16 != 0) { delete pl; fld_set_error("tc_conv: bad cout_pad %d", cout_pad); return FLD_ERR_INVALID; }
  p.BN = BN; p.cout_pad = cout_pad; p.n_ntiles = cout_pad / BN;
  p.flat = (g.kh == 1 && g.kw == 1 && g.stride == 1 && g.pad_t == 0 && g.pad_l == 0 && g.pool == 0 && g.OH == g.IH && g.OW == g.IW &&
            !getenv("FLD_TC_FLAT_OFF")) ? 1 : 0;
  p.npx = (long long)B * g.OH * g.OW;
  p.tiles_x = fld_div_up(g.OW, TW); p.tiles_y = fld_div_up(g.OH, TH); p.tiles_b = fld_div_up(B, NB);
  if (p.flat) { p.tiles_x = 1; p.tiles_y = 1; p.tiles_b = (int)((p.npx + 127) / 128); }
  if (ksplit > 1) {
    if (!p.flat || (Kc / 64) % ksplit != 0) { delete pl; fld_set_error("tc_conv: split-K needs a flat 1x1 conv and ksplit | K/64"); return FLD_ERR_INVALID; }
    p.ksplit = ksplit;
  }
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.n_ntiles * p.ksplit;
  p.sched = nullptr;
  p.n_iter = p.total_tiles;

  const size_t stage_bytes = 128 * 128 + (size_t)BN * 128;
  int stages = (int)((200 * 1024) / stage_bytes);
  stages = std::max(2, std::min(stages, kMaxStages));
  p.stages = stages;
  pl->smem = stages * stage_bytes + 1024;
  pl->grid = std::min(p.total_tiles, h->sm_count);
  if (row_tiles && !p.flat && p.total_tiles > pl->grid && p.total_tiles <= (1 << 20)) {
    // tile cost = k-blocks that touch the map (same formulas as ky_range / kx_range) + a fixed epilogue / pipeline-fill share
    const int mt = p.tiles_x * p.tiles_y * p.tiles_b, G = pl->grid, kch = Kc / 64;
    std::vector<std::pair<long long, int>> cost(p.total_tiles);
    for (int t = 0; t < p.total_tiles; ++t) {
      const int mm = (t % mt) % (p.tiles_x * p.tiles_y), ty = mm / p.tiles_x, tx = mm % p.tiles_x;
      const int r0 = ty * TH * g.stride, r1 = r0 + (TH - 1) * g.stride, c0 = tx * TW * g.stride, c1 = c0 + (TW - 1) * g.stride;
      const int nky = std::min(g.kh - 1, g.IH - 1 + g.pad_t - r0) - std::max(0, g.pad_t - r1) + 1;
      const int nkx = std::min(g.kw - 1, g.IW - 1 + g.pad_l - c0) - std::max(0, g.pad_l - c1) + 1;
      cost[t] = {-((long long)std::max(nky, 1) * std::max(nkx, 1) * kch * 2 * BN + 3000), t};   // ~cycles
    }
    std::stable_sort(cost.begin(), cost.end());
    std::vector<std::vector<int>> lists(G);
    std::vector<std::pair<long long, int>> heap(G);   // (load, cta) min-heap
    for (int c = 0; c < G; ++c) heap[c] = {0, c};
    auto cmp = [](const std::pair<long long, int>& a, const std::pair<long long, int>& b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    for (auto& ct : cost) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      auto& top = heap.back();
      lists[top.second].push_back(ct.second);
      top.first += -ct.first;
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
    size_t rounds = 0;
    for (auto& l : lists) { std::sort(l.begin(), l.end()); rounds = std::max(rounds, l.size()); }   // ascending id: N tile slow per CTA
    std::vector<int> sched(rounds * G, -1);
    for (int c = 0; c < G; ++c)
      for (size_t i = 0; i < lists[c].size(); ++i) sched[i * G + c] = lists[c][i];
    if (cudaMalloc(&pl->d_perm, sched.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(pl->d_perm, sched.data(), sched.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete pl; fld_set_error("tc_conv: tile schedule upload failed"); return FLD_ERR_CUDA;
    }
    p.sched = pl->d_perm;
    p.n_iter = (int)sched.size();
  }

  // activations: bf16 NHWC [B][IH][IW][Cin]
  if (p.flat) {
    cuuint64_t dims[2] = {(cuuint64_t)Ac, (cuuint64_t)p.npx};
    cuuint64_t strides[1] = {(cuuint64_t)Ac * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(A, flat) failed: %d", (int)r); return FLD_ERR_CUDA; }
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)Ac, (cuuint64_t)g.IW, (cuuint64_t)g.IH, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Ac * 2, (cuuint64_t)g.IW * Ac * 2, (cuuint64_t)g.IH * g.IW * Ac * 2};
    // boxDim counts TRAVERSED elements: with element stride s the box loads ceil(boxDim / s) pixels per spatial dimension
    const cuuint32_t s = (cuuint32_t)g.stride;
    cuuint32_t box[4] = {64, (cuuint32_t)TW * s, (cuuint32_t)TH * s, (cuuint32_t)NB};
    cuuint32_t es[4] = {1, s, s, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  // weights: bf16 [taps*cout_pad][Cin]
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kc, (cuuint64_t)g.kh * g.kw * cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Kc * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  *out = pl;
  return FLD_OK;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (p && p->d_perm) cudaFree(p->d_perm);
  delete p;
}

int tc_conv_run(const TcConvPlan* pl, const float* bias, void* out, int out_dtype, cudaStream_t st) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  TmaConvParams p = pl->p;
  p.bias = bias; p.out = out;
  if (getenv("FLD_TC_TRACE")) {
    static unsigned long long* tbuf = nullptr;
    if (!tbuf) FLD_CUDA(cudaMalloc(&tbuf, 3 * kTraceN * 8));
    FLD_CUDA(cudaMemsetAsync(tbuf, 0, 3 * kTraceN * 8, st));
    p.trace = tbuf;
  }
  if (p.pool && out_dtype != FLD_BF16) { fld_set_error("tc_conv: fused pool needs bf16 output"); return FLD_ERR_INVALID; }
  if (out_dtype == FLD_F32) {
    FLD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_tma_kernel<true><<<pl->grid, kTmaThreads, pl->smem, st>>>(pl->tmA, pl->tmB, p);
  } else {
    FLD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_tma_kernel<false><<<pl->grid, kTmaThreads, pl->smem, st>>>(pl->tmA, pl->tmB, p);
  }
  FLD_LAUNCHED();
  if (p.trace) {  // dump CTA 0's event log: "<role> <tag> <clock>" per line
    static int n_dump = 0;
    std::vector<unsigned long long> hbuf(3 * kTraceN);
    FLD_CUDA(cudaStreamSynchronize(st));
    FLD_CUDA(cudaMemcpy(hbuf.data(), p.trace, hbuf.size() * 8, cudaMemcpyDeviceToHost));
    char name[256];
    snprintf(name, sizeof(name), "%s/trace_%02d_cin%d_cout%d_oh%d.txt", getenv("FLD_TC_TRACE"), n_dump++, p.Cin, p.Cout, p.OH);
    if (FILE* f = fopen(name, "w")) {
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < kTraceN && hbuf[r * kTraceN + i]; ++i)
          fprintf(f, "%d %llu %llu\n", r, hbuf[r * kTraceN + i] & 15, hbuf[r * kTraceN + i] >> 4);
      fclose(f);
    }
  }
  return FLD_OK;
}
