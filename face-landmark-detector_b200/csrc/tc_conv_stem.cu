// tc_conv_stem.cu — strided first conv of the MobileNet / ResNet50 encoders on tcgen05, straight from the float32 (or uint8)
// image:  ZeroPadding2D + Conv2D(k x k, stride 2, Cin = 3) + folded BN + ReLU / ReLU6
//   reference networks/mobilenet.py:16-28 (3x3, stride 2, 32 filters), networks/resnet50.py:143-148 (7x7, stride 2, 64 filters).
//
// Same construction as tc_conv_first.cu (the 3x3 / stride-1 / pooled stage of the vanilla encoder), generalised over the
// kernel size KS and stride ST at compile time:
//   1. the ((TW-1)*ST + KS) x ((TH-1)*ST + KS) input patch of an 8 x 16 output tile is converted once to bf16 RGB0 (8 bytes
//      per pixel) and parked in shared memory;
//   2. thread r builds the K row of output pixel r: tap t = kh*KS + kw occupies k = 4t .. 4t+3, two taps per 16-byte chunk,
//      written in the UMMA no-swizzle K-major core-matrix layout; the slot after the last tap carries A = 1.0 for the bias
//      (B holds the bias split into bf16 hi / lo);
//   3. one thread issues K/16 MMAs (128 x Cout x 16) into TMEM; the shared epilogue applies the activation and stores bf16.
// K = 4*KS*KS + 4 padded to 16: 48 for 3x3, 208 for 7x7 (A tile 52 KB: two CTAs per SM).
#include <string.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

struct StemParams {
  const void* in;
  const __nv_bfloat16* w;  // [KG][Cout/8][8 rows][8 k] core-matrix packed
  __nv_bfloat16* out;
  int B, H, W;             // input size
  int OH, OW;              // output size
  int Cout, act;
  int pad_t, pad_l;
  int tiles_x, tiles_y, n_tiles;
};

constexpr int TW_ = 8, TH_ = 16;

template <typename TIn> struct Px3 { TIn c[3]; };

template <typename TIn>
__device__ __forceinline__ void load_px(const TIn* __restrict__ img, int H, int W, int y, int x, Px3<TIn>& v) {
  if (y >= 0 && y < H && x >= 0 && x < W) {
    const TIn* p = img + ((size_t)y * W + x) * 3;
    v.c[0] = __ldg(p); v.c[1] = __ldg(p + 1); v.c[2] = __ldg(p + 2);
  } else {
    v.c[0] = v.c[1] = v.c[2] = (TIn)0;  // ZeroPadding2D
  }
}

template <int KS> struct StemGeom {
  static constexpr int NTAP = KS * KS;
  static constexpr int NST = (NTAP + 2) / 2;        // 16-byte chunks: two taps each, the bias slot follows the last tap
  static constexpr int KG = (NST + 1) & ~1;         // k groups of 8, even (K16 steps)
};

template <typename TIn, int KS, int ST>
__global__ void __launch_bounds__(128)
conv_stem_kernel(const StemParams p) {
  using G = StemGeom<KS>;
  constexpr int PWd = (TW_ - 1) * ST + KS, PHd = (TH_ - 1) * ST + KS, NPIX = PWd * PHd, NF = (NPIX + 127) / 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;                                        // [KG][16 rowgroups][8 rows][16 B]
  uint8_t* sB = smem_raw + G::KG * 2048;                         // [KG][Cout/8][8][16 B]
  uint2* patch = reinterpret_cast<uint2*>(sB + (size_t)G::KG * p.Cout * 16);
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = p.Cout <= 32 ? 32 : p.Cout <= 64 ? 64 : p.Cout <= 128 ? 128 : 256;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < G::KG * p.Cout; i += 128) dst[i] = src[i];
    if (G::KG > G::NST)   // trailing all-zero k group of A
      *reinterpret_cast<uint4*>(sA + ((G::KG - 1) * 16 + (tid >> 3)) * 128 + (tid & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
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

  auto tile_origin = [&](int tile, int& b, int& ox0, int& oy0) {
    const int tx = tile % p.tiles_x;
    const int t2 = tile / p.tiles_x;
    const int ty = t2 % p.tiles_y;
    b = t2 / p.tiles_y;
    ox0 = tx * TW_; oy0 = ty * TH_;
  };
  auto fetch = [&](int tile, Px3<TIn> (&v)[NF]) {
    int b, ox0, oy0;
    tile_origin(tile, b, ox0, oy0);
    const TIn* img = reinterpret_cast<const TIn*>(p.in) + (size_t)b * p.H * p.W * 3;
    const int iy0 = oy0 * ST - p.pad_t, ix0 = ox0 * ST - p.pad_l;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int e = tid + 128 * j;
      if (e < NPIX) {
        const int r = e / PWd, c = e - r * PWd;
        load_px<TIn>(img, p.H, p.W, iy0 + r, ix0 + c, v[j]);
      }
    }
  };

  Px3<TIn> v[NF];
#pragma unroll
  for (int j = 0; j < NF; ++j) v[j].c[0] = v[j].c[1] = v[j].c[2] = (TIn)0;
  if ((int)blockIdx.x < p.n_tiles) fetch(blockIdx.x, v);

  uint32_t phase = 0;
  const int ly = tid >> 3, lx = tid & 7;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int b, ox0, oy0;
    tile_origin(tile, b, ox0, oy0);
    // ---- 1. park this tile's pixels as bf16 RGB0
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int e = tid + 128 * j;
      if (e < NPIX) patch[e] = make_uint2(pack_bf16((float)v[j].c[0], (float)v[j].c[1]), pack_bf16((float)v[j].c[2], 0.f));
    }
    __syncthreads();
    const int next = tile + gridDim.x;
    if (next < p.n_tiles) fetch(next, v);
    // ---- 2. K row of output pixel (ly, lx): two taps per 16-byte chunk
    {
      const uint2* pp = patch + (ly * ST) * PWd + lx * ST;
      uint8_t* row = sA + (tid >> 3) * 128 + (tid & 7) * 16;
#pragma unroll
      for (int s = 0; s < G::NST; ++s) {
        const int ta = 2 * s, tb = 2 * s + 1;
        uint2 qa, qb;
        if (ta < G::NTAP) qa = pp[(ta / KS) * PWd + (ta % KS)];
        else qa = (ta == G::NTAP) ? make_uint2(0x3f803f80u, 0u) : make_uint2(0u, 0u);   // bias slot: k = 4*NTAP, +1 carry 1.0
        if (tb < G::NTAP) qb = pp[(tb / KS) * PWd + (tb % KS)];
        else qb = (tb == G::NTAP) ? make_uint2(0x3f803f80u, 0u) : make_uint2(0u, 0u);
        *reinterpret_cast<uint4*>(row + s * 2048) = make_uint4(qa.x, qa.y, qb.x, qb.y);
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- 3. KG/2 MMAs of K = 16 (two k groups per step)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t astep = (uint64_t)((2 * a_lbo) >> 4), bstep = (uint64_t)((2 * b_lbo) >> 4);
#pragma unroll
        for (int j = 0; j < G::KG / 2; ++j) umma_bf16(tmem_base, adesc0 + j * astep, bdesc0 + j * bstep, idesc, j ? 1u : 0u);
        umma_commit(smem_u32(&mma_bar));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&mma_bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- 4. epilogue: thread = TMEM lane = output pixel
    EpiOut eo;
    eo.vec_ok = true;  // Cout % 16 == 0
    eo.valid = (oy0 + ly < p.OH) && (ox0 + lx < p.OW);
    eo.ptr = p.out + (((size_t)b * p.OH + (oy0 + ly)) * p.OW + (ox0 + lx)) * p.Cout;
    for (int ch = 0; ch < p.Cout; ch += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch, acc);
      tmem_ld_wait();
      EpiOut e2 = eo;
      e2.ptr = reinterpret_cast<__nv_bfloat16*>(eo.ptr) + ch;
      e2.c_left = p.Cout - ch;
      epilogue_chunk<false, false, false>(acc, nullptr, p.act, lane, TW_, e2);
    }
    tc_fence_before();
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

template <int KS>
size_t stem_smem(int Cout, int ST) {
  using G = StemGeom<KS>;
  const int PWd = (TW_ - 1) * ST + KS, PHd = (TH_ - 1) * ST + KS;
  return (size_t)G::KG * 2048 + (size_t)G::KG * Cout * 16 + (size_t)PWd * PHd * 8 + 64;
}

}  // namespace

bool tc_conv_stem_supported(const ConvGeom& g) {
  return g.Cin == 3 && g.kh == g.kw && (g.kh == 3 || g.kh == 7) && g.stride == 2 && g.pool == 0 && g.Cout % 16 == 0 && g.Cout >= 16 &&
         g.Cout <= 256 && g.pad_t >= 0 && g.pad_l >= 0 && !getenv("FLD_TC_STEM_OFF");
}

int tc_conv_stem_kgroups(int ks) { return ks == 7 ? StemGeom<7>::KG : StemGeom<3>::KG; }

// w_host fp32 [(kh*KS+kw)*3 + c][Cout] -> bf16 bits [KG][Cout/8][8][8], k' = 4*tap + c; k' = 4*NTAP / +1 = bias hi / lo
void tc_conv_stem_pack(const float* w_host, const float* bias_host, int ks, int Cout, uint16_t (*f2bf)(float), uint16_t* out) {
  const int ntap = ks * ks, KG = tc_conv_stem_kgroups(ks);
  for (int kg = 0; kg < KG; ++kg)
    for (int ng = 0; ng < Cout / 8; ++ng)
      for (int r = 0; r < 8; ++r)
        for (int e = 0; e < 8; ++e) {
          const int kp = kg * 8 + e, o = ng * 8 + r;
          const int tap = kp / 4, c = kp % 4;
          float v = 0.f;
          if (tap < ntap && c < 3) v = w_host[(size_t)(tap * 3 + c) * Cout + o];
          if (bias_host && tap == ntap && c < 2) {
            uint16_t hb = f2bf(bias_host[o]);
            uint32_t hu = (uint32_t)hb << 16;
            float hi;
            memcpy(&hi, &hu, 4);
            v = (c == 0) ? hi : bias_host[o] - hi;
          }
          out[(((size_t)kg * (Cout / 8) + ng) * 8 + r) * 8 + e] = f2bf(v);
        }
}

int tc_conv_stem(const fld_handle* h, const void* in, int in_dtype, const __nv_bfloat16* w_packed, __nv_bfloat16* out, const ConvGeom& g,
                 int B, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  StemParams p;
  p.in = in; p.w = w_packed; p.out = out;
  p.B = B; p.H = g.IH; p.W = g.IW; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout; p.act = g.act;
  p.pad_t = g.pad_t; p.pad_l = g.pad_l;
  p.tiles_x = fld_div_up(g.OW, TW_); p.tiles_y = fld_div_up(g.OH, TH_);
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  const size_t smem = g.kh == 7 ? stem_smem<7>(g.Cout, 2) : stem_smem<3>(g.Cout, 2);
  const int ncols = g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : g.Cout <= 128 ? 128 : 256;
  const int by_smem = (int)std::max<size_t>(1, (size_t)(220 * 1024) / (smem + 1024));
  const int cta_per_sm = std::max(1, std::min(std::min(512 / ncols, 8), by_smem));
  const int grid = std::min(p.n_tiles, h->sm_count * cta_per_sm);
  auto launch = [&](auto kern) -> int {
    FLD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (smem > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, st>>>(p);
    return FLD_OK;
  };
  int rc;
  if (in_dtype == FLD_U8) rc = g.kh == 7 ? launch(conv_stem_kernel<uint8_t, 7, 2>) : launch(conv_stem_kernel<uint8_t, 3, 2>);
  else if (in_dtype == FLD_F32) rc = g.kh == 7 ? launch(conv_stem_kernel<float, 7, 2>) : launch(conv_stem_kernel<float, 3, 2>);
  else {
    fld_set_error("tc_conv_stem: input must be u8 or f32");
    return FLD_ERR_INVALID;
  }
  if (rc) return rc;
  FLD_LAUNCHED();
  return FLD_OK;
}
