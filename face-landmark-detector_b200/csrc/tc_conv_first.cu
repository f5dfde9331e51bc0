// tc_conv_first.cu — first conv stage (3x3, Cin = 3, pad 1, stride 1 + bias + ReLU + MaxPool2) on tcgen05,
// straight from the uint8 (regression path, reference prediction.py:82-84) or float32 (FCN path, reference
// data/generator.py:53-61) image.  Replaces stage 1 of vanilla_encoder (reference networks/fcn.py:25-31).
//
// Cin = 3 makes TMA im2col useless (K = 27), so the CTA builds the A tile itself:
//   1. every input pixel of the (16+2) x (8+2) halo patch is converted ONCE to 4 x bf16 (RGB0, 8 bytes) and
//      parked in shared memory (uint8 0..255 is exact in bf16);
//   2. thread r = output pixel r of the 8 x 16 tile gathers its 3 x 3 pixels with nine 8-byte shared loads and
//      writes its 48-element K row (k = kh*12 + kw*4 + c; 36 real taps, 12 zero) with five 16-byte stores in the
//      UMMA no-swizzle K-major core-matrix layout;
//   3. one thread issues 3 MMAs (128 x Cout x 16) into TMEM; the fused epilogue of tc_common.cuh pools and stores.
//      The folded bias rides in two of the spare K slots (A = 1.0, B = bf16 hi / lo halves of the bias, fp32
//      accumulation: error <= 2^-17 |bias|), which removes 64 FADD + 16 bias loads per thread per tile.
// The next tile's pixels are prefetched into registers before the current tile's MMA/epilogue, so the global
// load latency (the dominant stall of the first version: ~31 % of samples on the staging store) is hidden.
// TMEM-read bound in the limit: 128 x Cout fp32 accumulators per 128 pixels at 64 B/clk/SM.
#include <string.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

struct FirstParams {
  const void* in;
  const __nv_bfloat16* w;  // [6 kgroups][Cout/8][8 rows][8 k] core-matrix packed, k = kh*12 + kw*4 + c
  const float* bias;
  __nv_bfloat16* out;
  int B, H, W;             // input == conv output size
  int Cout;                // 16..256, multiple of 16
  int act, pool;
  int tiles_x, tiles_y, n_tiles;
  unsigned long long mul_x, mul_y;  // ceil(2^40 / tiles_x), ceil(2^40 / tiles_y): exact division for n < 2^21
};

constexpr int TWc = 8, THc = 16, PW_ = TWc + 2, PH_ = THc + 2, NPIX = PW_ * PH_;  // 10 x 18 = 180 halo pixels

template <typename TIn> struct Raw3 { TIn c[3]; };

template <typename TIn>
__device__ __forceinline__ void load_pixel(const TIn* __restrict__ img, int H, int W, int y, int x, Raw3<TIn>& v) {
  if (y >= 0 && y < H && x >= 0 && x < W) {
    const TIn* p = img + ((size_t)y * W + x) * 3;
    v.c[0] = __ldg(p); v.c[1] = __ldg(p + 1); v.c[2] = __ldg(p + 2);
  } else {
    v.c[0] = v.c[1] = v.c[2] = (TIn)0;  // ZeroPadding2D(1)
  }
}

template <typename TIn>
__global__ void __launch_bounds__(128, 8)
conv_first_kernel(const FirstParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // A [6 kgroups][16 rowgroups][8 rows][16 B] = 12 KB | B [6 kgroups][Cout/8][8][16 B] = Cout*96 B | patch [18][10] x 8 B
  uint8_t* sA = smem_raw;
  uint8_t* sB = smem_raw + 12288;
  uint2* patch = reinterpret_cast<uint2*>(sB + p.Cout * 96);
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = p.Cout <= 32 ? 32 : p.Cout <= 64 ? 64 : p.Cout <= 128 ? 128 : 256;

  {  // weights (already in core-matrix order) and the all-zero sixth K group of A
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < p.Cout * 6; i += 128) dst[i] = src[i];
    *reinterpret_cast<uint4*>(sA + (5 * 16 + (tid >> 3)) * 128 + (tid & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mma_bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_s), ncols);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16(128, p.Cout);
  const uint32_t a_lbo = 16 * 128, b_lbo = (uint32_t)(p.Cout / 8) * 128;
  const uint64_t adesc0 = umma_desc(smem_u32(sA), a_lbo, 128, 0);
  const uint64_t bdesc0 = umma_desc(smem_u32(sB), b_lbo, 128, 0);

  // halo pixels owned by this thread: e0 = tid, e1 = tid + 128 (< 180 for tid < 52)
  const int r0 = tid / PW_, c0 = tid - r0 * PW_;
  const int e1 = tid + 128;
  const bool has1 = e1 < NPIX;
  const int r1 = e1 / PW_, c1 = e1 - r1 * PW_;

  auto tile_origin = [&](int tile, int& b, int& x0, int& y0) {
    const int t2 = (int)(((unsigned long long)(unsigned)tile * p.mul_x) >> 40);
    const int tx = tile - t2 * p.tiles_x;
    b = (int)(((unsigned long long)(unsigned)t2 * p.mul_y) >> 40);
    const int ty = t2 - b * p.tiles_y;
    x0 = tx * TWc; y0 = ty * THc;
  };
  auto fetch = [&](int tile, Raw3<TIn>& v0, Raw3<TIn>& v1) {
    int b, x0, y0;
    tile_origin(tile, b, x0, y0);
    const TIn* img = reinterpret_cast<const TIn*>(p.in) + (size_t)b * p.H * p.W * 3;
    load_pixel<TIn>(img, p.H, p.W, y0 - 1 + r0, x0 - 1 + c0, v0);
    if (has1) load_pixel<TIn>(img, p.H, p.W, y0 - 1 + r1, x0 - 1 + c1, v1);
  };

  Raw3<TIn> v0, v1;
  v1.c[0] = v1.c[1] = v1.c[2] = (TIn)0;
  if ((int)blockIdx.x < p.n_tiles) fetch(blockIdx.x, v0, v1);

  uint32_t phase = 0;
  const int PH = p.H >> 1, PW = p.W >> 1;
  const int ly = tid >> 3, lx = tid & 7;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int b, x0, y0;
    tile_origin(tile, b, x0, y0);
    // ---- 1. park this tile's pixels as bf16 RGB0
    patch[tid] = make_uint2(pack_bf16((float)v0.c[0], (float)v0.c[1]), pack_bf16((float)v0.c[2], 0.f));
    if (has1) patch[e1] = make_uint2(pack_bf16((float)v1.c[0], (float)v1.c[1]), pack_bf16((float)v1.c[2], 0.f));
    __syncthreads();
    // ---- prefetch the next tile's pixels (in flight during im2col + MMA + epilogue)
    const int next = tile + gridDim.x;
    if (next < p.n_tiles) fetch(next, v0, v1);
    // ---- 2. im2col row of output pixel (ly, lx)
    {
      uint2 q[3][3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) q[kh][kw] = patch[(ly + kh) * PW_ + lx + kw];
      uint8_t* row = sA + (tid >> 3) * 128 + (tid & 7) * 16;  // + kgroup * 16 * 128
      *reinterpret_cast<uint4*>(row + 0 * 2048) = make_uint4(q[0][0].x, q[0][0].y, q[0][1].x, q[0][1].y);
      *reinterpret_cast<uint4*>(row + 1 * 2048) = make_uint4(q[0][2].x, q[0][2].y, q[1][0].x, q[1][0].y);
      *reinterpret_cast<uint4*>(row + 2 * 2048) = make_uint4(q[1][1].x, q[1][1].y, q[1][2].x, q[1][2].y);
      *reinterpret_cast<uint4*>(row + 3 * 2048) = make_uint4(q[2][0].x, q[2][0].y, q[2][1].x, q[2][1].y);
      *reinterpret_cast<uint4*>(row + 4 * 2048) = make_uint4(q[2][2].x, q[2][2].y, 0x3f803f80u, 0u);  // k 36,37 = 1.0 (bias)
    }
    fence_async_smem();
    __syncthreads();
    // ---- 3. three K = 16 MMAs (k groups 0-1, 2-3, 4-5); LBO field moves by 2 groups per step
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {  // single-lane region known to the compiler: plain UTCHMMA / UTCBAR, no waterfall loops
        const uint64_t astep = (uint64_t)((2 * a_lbo) >> 4), bstep = (uint64_t)((2 * b_lbo) >> 4);
        umma_bf16(tmem_base, adesc0, bdesc0, idesc, 0u);
        umma_bf16(tmem_base, adesc0 + astep, bdesc0 + bstep, idesc, 1u);
        umma_bf16(tmem_base, adesc0 + 2 * astep, bdesc0 + 2 * bstep, idesc, 1u);
        umma_commit(smem_u32(&mma_bar));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&mma_bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- 4. epilogue: thread = TMEM lane = tile pixel
    EpiOut eo;
    eo.vec_ok = true;  // Cout % 16 == 0
    if (p.pool) {
      const int py = (y0 + ly) >> 1, px = (x0 + lx) >> 1;
      eo.valid = (py < PH) && (px < PW);
      eo.ptr = p.out + (((size_t)b * PH + py) * PW + px) * p.Cout;
    } else {
      eo.valid = (y0 + ly < p.H) && (x0 + lx < p.W);
      eo.ptr = p.out + (((size_t)b * p.H + (y0 + ly)) * p.W + (x0 + lx)) * p.Cout;
    }
    for (int ch = 0; ch < p.Cout; ch += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch, acc);
      tmem_ld_wait();
      EpiOut e2 = eo;
      e2.ptr = reinterpret_cast<__nv_bfloat16*>(eo.ptr) + ch;
      e2.c_left = p.Cout - ch;
      if (p.pool) epilogue_chunk<true, false, false>(acc, p.bias + ch, p.act, lane, TWc, e2);
      else epilogue_chunk<false, false, false>(acc, p.bias + ch, p.act, lane, TWc, e2);
    }
    tc_fence_before();
    __syncthreads();  // TMEM, A tile and patch are free for the next tile
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace

bool tc_conv_first_supported(const ConvGeom& g) {
  return g.kh == 3 && g.kw == 3 && g.Cin == 3 && g.stride == 1 && g.pad_t == 1 && g.pad_l == 1 && g.OH == g.IH && g.OW == g.IW &&
         g.Cout % 16 == 0 && g.Cout >= 16 && g.Cout <= 256 && (g.pool == 0 || g.pool == 2);
}

// host-side weight packing for conv_first_kernel: w_host fp32 [27][Cout] (k = (kh*3+kw)*3 + c) -> bf16 bits
// [6][Cout/8][8][8] with k' = kh*12 + kw*4 + c; k' = 36 / 37 carry the bias split into bf16 hi / lo
void tc_conv_first_pack(const float* w_host, const float* bias_host, int Cout, uint16_t (*f2bf)(float), uint16_t* out) {
  for (int kg = 0; kg < 6; ++kg)
    for (int ng = 0; ng < Cout / 8; ++ng)
      for (int r = 0; r < 8; ++r)
        for (int e = 0; e < 8; ++e) {
          const int kp = kg * 8 + e, o = ng * 8 + r;
          const int kh = kp / 12, kw = (kp % 12) / 4, c = kp % 4;
          float v = 0.f;
          if (kp < 36 && c < 3) v = w_host[(size_t)((kh * 3 + kw) * 3 + c) * Cout + o];
          if (bias_host && (kp == 36 || kp == 37)) {  // bias = hi + lo (both bf16), multiplied by A[k] = 1.0
            uint16_t hb = f2bf(bias_host[o]);
            uint32_t hu = (uint32_t)hb << 16;
            float hi;
            memcpy(&hi, &hu, 4);
            v = (kp == 36) ? hi : bias_host[o] - hi;
          }
          out[(((size_t)kg * (Cout / 8) + ng) * 8 + r) * 8 + e] = f2bf(v);
        }
}

int tc_conv_first(const fld_handle* h, const void* in, int in_dtype, const __nv_bfloat16* w_packed, const float* bias,
                  __nv_bfloat16* out, const ConvGeom& g, int B, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  FirstParams p;
  p.in = in; p.w = w_packed; p.bias = bias; p.out = out;
  p.B = B; p.H = g.IH; p.W = g.IW; p.Cout = g.Cout; p.act = g.act; p.pool = g.pool;
  p.tiles_x = fld_div_up(g.OW, TWc); p.tiles_y = fld_div_up(g.OH, THc);
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  p.mul_x = ((1ull << 40) + p.tiles_x - 1) / p.tiles_x;
  p.mul_y = ((1ull << 40) + p.tiles_y - 1) / p.tiles_y;
  if (p.n_tiles >= (1 << 21)) { fld_set_error("tc_conv_first: too many tiles (%d)", p.n_tiles); return FLD_ERR_INVALID; }
  const size_t smem = 12288 + (size_t)g.Cout * 96 + NPIX * 8 + 64;
  const int ncols = g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : g.Cout <= 128 ? 128 : 256;
  const int cta_per_sm = std::max(1, std::min(512 / ncols, 8));
  const int grid = std::min(p.n_tiles, h->sm_count * cta_per_sm);
  FLD_CUDA(cudaFuncSetAttribute(conv_first_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  FLD_CUDA(cudaFuncSetAttribute(conv_first_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  if (in_dtype == FLD_U8) {
    if (smem > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(conv_first_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_first_kernel<uint8_t><<<grid, 128, smem, st>>>(p);
  } else if (in_dtype == FLD_F32) {
    if (smem > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(conv_first_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_first_kernel<float><<<grid, 128, smem, st>>>(p);
  } else {
    fld_set_error("tc_conv_first: input must be u8 or f32");
    return FLD_ERR_INVALID;
  }
  FLD_LAUNCHED();
  return FLD_OK;
}
