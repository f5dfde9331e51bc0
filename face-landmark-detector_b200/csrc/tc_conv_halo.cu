// tc_conv_halo.cu — 3x3 (pad 1, stride 1) tcgen05 convolution with HALO REUSE and optional STATIONARY WEIGHTS.
//
// Same math and epilogue as tc_conv.cu (reference networks/fcn.py:33-49 stages; vgg16.py:27-73), different
// operand feed.  Measured on B200 (FLD_TC_TRACE + ncu, DESIGN.md "conv kernels"): the per-tap TMA version is bound
// by how many operand bytes must be (re)filled into shared memory per MMA cycle — 4 stages x 48 KB with a
// ~2800-cycle refill loop give ~700 cycles per k-block against 512 cycles of MMA (N = 256), and 31 % tensor-active
// at N = 128.  This kernel cuts the fill:
//   * A operand: ONE 4-D TMA box per 64-channel chunk brings the (TH+2) x 16 pixel halo patch of the 16 x 8 output
//     tile (36 KB) instead of nine shifted 16 KB boxes.  The nine taps are nine UMMA descriptors into that patch:
//     start address shifted by (ky*16 + kx) pixels (128 B each), stride between 8-row groups = 16 pixels = 2048 B.
//     MEASURED on B200: the tensor core derives the 128B-swizzle phase from the absolute shared-memory address
//     bits [7,10) exactly like TMA does when it writes, so a start address that is only 128-byte aligned needs NO
//     descriptor base_offset (base_offset = kx gives wrong results; 0 is bit-correct — tests/test_gpu_parity.py).
//   * B operand: a ring of (tap, chunk) weight tiles as before, or — when all taps*chunks*BN*128 B fit beside two
//     halo buffers (conv2 of the vanilla trunk: 147 KB) — loaded ONCE per CTA and kept stationary.
// Tile = 8 wide x 16 high output pixels of one image (TW = 8, TH = 16, NB = 1), so the layer needs OH >= 16.
#include <stdlib.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

struct HaloParams {
  const float* bias;
  void* out;
  int B, OH, OW, Cout, Cin;
  int BN, cout_pad;
  int tiles_x, tiles_y, total_tiles;
  int SA, SB;          // halo ring depth, weight ring depth (SB = 0: stationary weights)
  int act, pool;
  int a_chunks;        // distinct 64-channel chunks of the activation tensor; k-chunk kc reads activation chunk kc % a_chunks.
                       // FLD_BF16X3: activations are [x_hi | x_lo] (a_chunks = 2 Cin/64), weights [w_hi | w_hi | w_lo] (3 Cin/64 chunks)
  int split;           // store the output as a SPLIT tensor ([hi | lo] bf16, pixel pitch 2 * Cout)
  int baseoff;         // bring-up switch FLD_TC_HALO_BASEOFF=1: set descriptor base_offset = kx (WRONG on B200; default 0)
};

constexpr int kHaloW = 16, kHaloH = 18;                 // halo box: 16 x 18 pixels x 64 channels
constexpr uint32_t kHaloBytes = kHaloW * kHaloH * 128;  // 36864 = 36 * 1024
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxRing = 6;

__device__ __forceinline__ uint64_t with_base_offset(uint64_t desc, uint32_t off) { return desc | ((uint64_t)(off & 7) << 49); }

__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t fullA[2], emptyA[2], fullB[kMaxRing], emptyB[kMaxRing], wfull, tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.BN * 128;
  const uint32_t smemA = smem_base;                           // SA halo buffers
  const uint32_t smemB = smem_base + p.SA * kHaloBytes;        // weight ring, or all 9*kchunks weight tiles
  const int kchunks = p.Cin >> 6;
  const bool stationary = p.SB == 0;
  const uint32_t fullA0 = smem_u32(&fullA[0]), emptyA0 = smem_u32(&emptyA[0]);
  const uint32_t fullB0 = smem_u32(&fullB[0]), emptyB0 = smem_u32(&emptyB[0]);
  const uint32_t wfull0 = smem_u32(&wfull), tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(fullA0 + 8 * s, 1); mbar_init(emptyA0 + 8 * s, 1); }
    for (int s = 0; s < kMaxRing; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
    mbar_init(wfull0, 1);
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

  // chunk-iterations of this CTA: g = (local tile index) * kchunks + kc
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) ++my_tiles;
  const int G = my_tiles * kchunks;
  const int txy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: ONE elected lane runs the whole
    // loop (the compiler then knows the region is single-lane: plain UTMALDG / SYNCS, no per-instruction waterfall)
    if (elect_one()) {
      auto issue_halo = [&](int g) {
        const int lt = g / kchunks, kc = g - lt * kchunks;
        const int tile = blockIdx.x + lt * gridDim.x;
        const int b = tile / txy;
        const int m = tile - b * txy;
        const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
        const int slot = g & 1;
        const int ka = kc >= p.a_chunks ? kc - p.a_chunks : kc;
        mbar_wait(emptyA0 + 8 * slot, ((g >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(fullA0 + 8 * slot, kHaloBytes);
        tma_load_4d(smemA + slot * kHaloBytes, &tmA, fullA0 + 8 * slot, ka * 64, tx * 8 - 1, ty * 16 - 1, b);
      };
      if (stationary) {
        mbar_arrive_expect_tx(wfull0, 9u * kchunks * b_bytes);
        for (int t = 0; t < 9; ++t)
          for (int kc = 0; kc < kchunks; ++kc)
            tma_load_2d(smemB + (t * kchunks + kc) * b_bytes, &tmB, wfull0, kc * 64, t * p.cout_pad);
        for (int g = 0; g < G; ++g) issue_halo(g);
      } else {
        uint32_t sb = 0, phb = 0;
        if (G > 0) issue_halo(0);
        for (int g = 0; g < G; ++g) {
          const int kc = g % kchunks;
          for (int t = 0; t < 9; ++t) {
            if (t == 2 && g + 1 < G) issue_halo(g + 1);  // prefetch the next halo a few taps ahead
            mbar_wait(emptyB0 + 8 * sb, phb ^ 1);
            mbar_arrive_expect_tx(fullB0 + 8 * sb, b_bytes);
            tma_load_2d(smemB + sb * b_bytes, &tmB, fullB0 + 8 * sb, kc * 64, t * p.cout_pad);
            if (++sb == (uint32_t)p.SB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one elected lane, taps unrolled
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      // A: K-major SW128, rows of 128 B, 8-row groups 2048 B apart (one halo row of 16 pixels)
      const uint64_t adesc0 = umma_desc(smemA, 16, kHaloW * 128, 2);
      const uint64_t bdesc0 = umma_desc(smemB, 16, 1024, 2);
      const uint32_t b_step = b_bytes >> 4;
      uint32_t sb = 0, phb = 0, acc = 0, acc_phase = 0;
      if (stationary) { mbar_wait(wfull0, 0); tc_fence_after(); }
      for (int g = 0; g < G; ++g) {
        const int kc = g % kchunks;
        if (kc == 0) {
          mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
        }
        const uint32_t d = tmem_base + acc * 256;
        const int slot = g & 1;
        mbar_wait(fullA0 + 8 * slot, (g >> 1) & 1);
        tc_fence_after();
        const uint64_t a_slot = adesc0 + (uint64_t)((slot * kHaloBytes) >> 4);
        uint64_t bd = bdesc0 + (uint64_t)(kc * b_step);   // stationary: tile (t*kchunks + kc)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t - 3 * ky;
          if (!stationary) {
            mbar_wait(fullB0 + 8 * sb, phb);
            tc_fence_after();
            bd = bdesc0 + (uint64_t)(sb * b_step);
          }
          // shifted window: + (ky*16 + kx) pixels of 128 B (8 units of 16 B each); no base_offset (see header)
          uint64_t ad = a_slot + (uint64_t)((ky * kHaloW + kx) * 8);
          if (p.baseoff) ad = with_base_offset(ad, (uint32_t)kx);
          umma_bf16(d, ad, bd, idesc, (kc | t) ? 1u : 0u);
          umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
          umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
          umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
          if (stationary) {
            bd += (uint64_t)(kchunks * b_step);
          } else {
            umma_commit(emptyB0 + 8 * sb);
            if (++sb == (uint32_t)p.SB) { sb = 0; phb ^= 1; }
          }
        }
        umma_commit(emptyA0 + 8 * slot);
        if (kc == kchunks - 1) {
          umma_commit(tfull0 + 8 * acc);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (same mapping as tc_conv.cu, TW = 8)
    const int ew = warp - 2, sub = warp & 3, half = ew >> 2;
    const int r = sub * 32 + lane;
    const int lx = r & 7, ly = r >> 3;
    const int PH = p.OH >> 1, PW = p.OW >> 1;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / txy;
      const int m = tile - b * txy;
      const int ty = m / p.tiles_x, tx = m - ty * p.tiles_x;
      const int ox = tx * 8 + lx, oy = ty * 16 + ly;
      EpiOut eo;
      eo.vec_ok = (p.Cout % 8 == 0);
      size_t pix;
      if (p.pool) {
        eo.valid = ((oy >> 1) < PH) && ((ox >> 1) < PW);
        pix = ((size_t)b * PH + (oy >> 1)) * PW + (ox >> 1);
      } else {
        eo.valid = (oy < p.OH) && (ox < p.OW);
        pix = ((size_t)b * p.OH + oy) * p.OW + ox;
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      for (int ch = half * 32; ch < p.BN; ch += 64) {
        uint32_t regs[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + ch, regs);
        tmem_ld_wait();
        EpiOut e2 = eo;
        e2.c_left = p.Cout - ch;
        if (p.split) {
          e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * (2 * p.Cout) + ch;
          if (p.pool) epilogue_chunk_split<true>(regs, p.bias + ch, p.act, lane, 8, e2, p.Cout);
          else epilogue_chunk_split<false>(regs, p.bias + ch, p.act, lane, 8, e2, p.Cout);
          continue;
        }
        e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + ch;
        if (p.pool) epilogue_chunk<true, false>(regs, p.bias + ch, p.act, lane, 8, e2);
        else epilogue_chunk<false, false>(regs, p.bias + ch, p.act, lane, 8, e2);
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

}  // namespace

struct TcHaloPlan {
  CUtensorMap tmA, tmB;
  HaloParams p;
  int grid;
  size_t smem;
};

bool tc_halo_supported(const ConvGeom& g, int cout_pad) {
  static const bool enabled = !(getenv("FLD_TC_HALO") && atoi(getenv("FLD_TC_HALO")) == 0);
  if (!enabled) return false;
  if (g.kh != 3 || g.kw != 3 || g.stride != 1 || g.pad_t != 1 || g.pad_l != 1) return false;
  if (g.Cin % 64 != 0 || g.OH != g.IH || g.OW != g.IW) return false;
  if (g.OH < 16 || g.OW < 8) return false;
  if (cout_pad > 256 || cout_pad % 16 != 0) return false;  // one N tile
  if (g.pool != 0 && g.pool != 2) return false;
  return true;
}

int tc_halo_plan_create(const fld_handle* h, const void* in, const __nv_bfloat16* w_packed, int cout_pad, const ConvGeom& g, int B,
                        TcHaloPlan** out, int x3, int split_out) {
  if (!h->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  TcHaloPlan* pl = new TcHaloPlan();
  HaloParams& p = pl->p;
  p.bias = nullptr; p.out = nullptr;
  // x3: the K dimension is [x_hi w_hi | x_lo w_hi | x_hi w_lo] = 3 Cin channels over a 2 Cin-channel activation tensor
  const int Kc = x3 ? 3 * g.Cin : g.Cin, Ac = x3 ? 2 * g.Cin : g.Cin;
  p.B = B; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout; p.Cin = Kc;
  p.a_chunks = Ac / 64; p.split = split_out;
  p.BN = cout_pad; p.cout_pad = cout_pad;
  p.tiles_x = fld_div_up(g.OW, 8); p.tiles_y = fld_div_up(g.OH, 16);
  p.total_tiles = p.tiles_x * p.tiles_y * B;
  p.act = g.act; p.pool = g.pool;
  { const char* e = getenv("FLD_TC_HALO_BASEOFF"); p.baseoff = e ? atoi(e) : 0; }
  const size_t b_bytes = (size_t)cout_pad * 128;
  const size_t budget = 226 * 1024;
  const size_t w_all = 9 * (size_t)(Kc / 64) * b_bytes;
  p.SA = 2;
  if (w_all + 2 * kHaloBytes + 1024 <= budget) {
    p.SB = 0;  // stationary weights
    pl->smem = 2 * kHaloBytes + w_all + 1024;
  } else {
    int sb = (int)((budget - 1024 - 2 * kHaloBytes) / b_bytes);
    p.SB = std::max(2, std::min(sb, kMaxRing));
    pl->smem = 2 * kHaloBytes + p.SB * b_bytes + 1024;
  }
  pl->grid = std::min(p.total_tiles, h->sm_count);
  {
    cuuint64_t dims[4] = {(cuuint64_t)Ac, (cuuint64_t)g.IW, (cuuint64_t)g.IH, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Ac * 2, (cuuint64_t)g.IW * Ac * 2, (cuuint64_t)g.IH * g.IW * Ac * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(halo A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kc, (cuuint64_t)9 * cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Kc * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)cout_pad};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(halo B) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  *out = pl;
  return FLD_OK;
}

void tc_halo_plan_destroy(TcHaloPlan* p) { delete p; }

int tc_halo_run(const TcHaloPlan* pl, const float* bias, void* out, cudaStream_t st) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  HaloParams p = pl->p;
  p.bias = bias; p.out = out;
  FLD_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  conv_halo_kernel<<<pl->grid, kThreads, pl->smem, st>>>(pl->tmA, pl->tmB, p);
  FLD_LAUNCHED();
  return FLD_OK;
}
