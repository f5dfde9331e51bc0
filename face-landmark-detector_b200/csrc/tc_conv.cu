// tc_conv.cu — tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (FLD_BF16 mode).
//
// Replaces the TensorFlow conv arithmetic behind reference networks/fcn.py:25-49 (ZeroPad -> Conv3x3 ->
// BN -> ReLU -> MaxPool2 stages), fcn.py:98-103,108,117 (7x7 / 1x1 head and score convs).  BN is folded
// into weights/bias on the host (net.cu); bias + ReLU + 2x2 max-pool are fused into the epilogue so a
// stage's pre-pool map never touches HBM.
//
// GEMM view: D[M = output pixels][N = Cout] = A[M][K = taps*Cin] * B[N][K]^T, bf16 operands, fp32
// accumulation in TMEM.  An M tile is a TW x TH x NB pixel patch (128 pixels) so that
//   (1) one 4-D TMA box per (tap, 64-channel chunk) fetches the shifted input patch, with TMA's
//       out-of-bounds zero fill supplying the convolution's zero padding, and
//   (2) the four pixels of every 2x2 pool window sit in one warp's 32 TMEM lanes (lane^1, lane^TW):
//       pooling is two rounds of warp shuffles on packed bf16x2.
//
// Two kernels:
//   conv_first_kernel  3x3xCin=3 first layer straight from the u8 / f32 image: the CTA builds the
//                      im2col tile (K = 27 -> 32) in shared memory itself (no padded copy of the input),
//                      2 MMAs of 128xCoutx16 per tile.
//   conv_tma_kernel    generic kh x kw, Cin % 64 == 0, stride 1: warp-specialised (TMA producer / MMA
//                      issuer / 4 epilogue warps), multi-stage mbarrier ring, double-buffered TMEM
//                      accumulators, persistent over tiles.
#include <cuda.h>
#include "ops.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < 20000000u; ++it)
    if (mbar_try_wait(bar, parity)) return;
  printf("fld: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (0 none, 2 = 128B swizzle).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1, K-major A/B,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float act_f(float v, int act) {
  if (act == FLD_ACT_RELU) return fmaxf(v, 0.f);
  if (act == FLD_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// ------------------------------------------------------------------------------------------------
// Epilogue for one 32-column chunk held by one thread (= one output pixel / TMEM lane).
// geometry of the M tile: lane bits select the pixel inside the warp's 32-pixel slab; pool partners are
// lane^1 (x) and lane^TW (y).  Each of the 4 lanes of a pool window ends up storing a different 8-channel
// quarter of the pooled 32 channels (one 16-byte store each).
// ------------------------------------------------------------------------------------------------
struct EpiOut {
  void* ptr;       // pointer to channel (n0 + chunk*32) of this thread's output pixel (pooled pixel when pooling)
  bool valid;      // pixel inside the output
  int c_left;      // channels left from the chunk start (Cout - n0 - chunk*32), may be <= 0
  bool vec_ok;     // pixel pitch keeps 16-byte vector stores aligned (Cout % 8 == 0 for bf16, % 4 for f32)
};

template <bool POOL, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&acc)[32], const float* __restrict__ bias, int act, int lane, int TW,
                                               const EpiOut& o) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = act_f(__uint_as_float(acc[j]) + bias[j], act);
  if (POOL) {
    // bf16 output only.  round first (monotone, so max commutes), then max on packed pairs
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    const bool bx = lane & 1;
    uint32_t k1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = (i < 4) ? i : i + 4;
      const uint32_t send = bx ? w[a] : w[a + 4];
      const uint32_t keep = bx ? w[a + 4] : w[a];
      k1[i] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 1));
    }
    const bool by = (lane & TW) != 0;
    uint32_t k2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t send = by ? k1[i] : k1[i + 4];
      const uint32_t keep = by ? k1[i + 4] : k1[i];
      k2[i] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, TW));
    }
    const int cb = (bx ? 8 : 0) + (by ? 16 : 0);
    if (o.valid && cb + 8 <= o.c_left) {
      uint4 q = make_uint4(k2[0], k2[1], k2[2], k2[3]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(o.ptr) + cb) = q;
    }
  } else if (!OUT_F32) {
    if (!o.valid) return;
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(o.ptr);
    if (o.c_left >= 32 && o.vec_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                             pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        *reinterpret_cast<uint4*>(p + 8 * q) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < o.c_left) p[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    if (!o.valid) return;
    float* p = reinterpret_cast<float*>(o.ptr);
    if (o.c_left >= 32 && o.vec_ok) {
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < o.c_left) p[j] = v[j];
    }
  }
}

// ================================================================================================
// First layer: 3x3, Cin = 3, pad 1, stride 1.  Tile = 8 wide x 16 high output pixels of one image.
// ================================================================================================
struct FirstParams {
  const void* in;
  const __nv_bfloat16* w;  // [Cout/8][4][8][8]: core-matrix packed (n-group, k-group, row, 8 k-elems) ... see pack_first_weights
  const float* bias;
  __nv_bfloat16* out;
  int B, H, W;     // input == conv output size
  int Cout;        // 16..256, multiple of 16
  int act, pool;
  int tiles_x, tiles_y, n_tiles;
  float in_scale_unused;
};

template <typename TIn>
__global__ void __launch_bounds__(128)
conv_first_kernel(const FirstParams p) {
  constexpr int TWc = 8, THc = 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // layout: A [4 kgroups][16 rowgroups][8 rows][16 B] = 8 KB | B [4 kgroups][Cout/8][8][16 B] = Cout*64 B | patch | bars
  uint8_t* sA = smem_raw;
  uint8_t* sB = smem_raw + 8192;
  TIn* patch = reinterpret_cast<TIn*>(sB + p.Cout * 64);  // [18][32] elements (30 used per row)
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = p.Cout <= 32 ? 32 : p.Cout <= 64 ? 64 : p.Cout <= 128 ? 128 : 256;

  // weights -> smem (already in core-matrix order, so a straight copy)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < p.Cout * 4; i += 128) dst[i] = src[i];
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
  const uint32_t a_lbo = 16 * 128, a_sbo = 128;
  const uint32_t b_lbo = (uint32_t)(p.Cout / 8) * 128, b_sbo = 128;

  uint32_t phase = 0;
  const int PH = p.H >> 1, PW = p.W >> 1;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int x0 = tx * TWc, y0 = ty * THc;
    const TIn* img = reinterpret_cast<const TIn*>(p.in) + (size_t)b * p.H * p.W * 3;
    // ---- stage the (TH+2) x (TW+2) x 3 halo patch, zero outside the image (ZeroPadding2D(1))
    for (int e = tid; e < 18 * 30; e += 128) {
      const int r = e / 30, c = e - r * 30;
      const int iy = y0 - 1 + r, ixc = (x0 - 1) * 3 + c;
      TIn v = (TIn)0;
      if (iy >= 0 && iy < p.H && ixc >= 0 && ixc < p.W * 3) v = img[(size_t)iy * p.W * 3 + ixc];
      patch[r * 32 + c] = v;
    }
    __syncthreads();
    // ---- im2col row of this thread: pixel (ly, lx) = (tid / 8, tid % 8); k = (kh*3 + kw)*3 + c
    {
      const int ly = tid >> 3, lx = tid & 7;
      float f[32];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < 9; ++j) f[kh * 9 + j] = (float)patch[(ly + kh) * 32 + lx * 3 + j];
#pragma unroll
      for (int j = 27; j < 32; ++j) f[j] = 0.f;
      const int rg = tid >> 3, rr = tid & 7;
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) {
        uint4 u = make_uint4(pack_bf16(f[8 * kg], f[8 * kg + 1]), pack_bf16(f[8 * kg + 2], f[8 * kg + 3]),
                             pack_bf16(f[8 * kg + 4], f[8 * kg + 5]), pack_bf16(f[8 * kg + 6], f[8 * kg + 7]));
        *reinterpret_cast<uint4*>(sA + (kg * 16 + rg) * 128 + rr * 16) = u;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t ad = umma_desc(smem_u32(sA) + ks * 2 * a_lbo, a_lbo, a_sbo, 0);
        const uint64_t bd = umma_desc(smem_u32(sB) + ks * 2 * b_lbo, b_lbo, b_sbo, 0);
        umma_bf16(tmem_base, ad, bd, idesc, ks);
      }
      umma_commit(smem_u32(&mma_bar));
    }
    mbar_wait(smem_u32(&mma_bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread = TMEM lane = tile pixel tid
    const int ly = tid >> 3, lx = tid & 7;
    EpiOut eo;
    eo.vec_ok = true;  // Cout % 16 == 0 here
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
      if (p.pool) epilogue_chunk<true, false>(acc, p.bias + ch, p.act, lane, TWc, e2);
      else epilogue_chunk<false, false>(acc, p.bias + ch, p.act, lane, TWc, e2);
    }
    tc_fence_before();
    __syncthreads();  // TMEM + smem A/patch free for the next tile
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ================================================================================================
// Generic conv: warp-specialised TMA -> tcgen05.mma -> epilogue, persistent over tiles.
// ================================================================================================
struct TmaConvParams {
  const float* bias;   // [Cout_pad]
  void* out;
  int B, OH, OW, Cout; // conv output (pre-pool) dims
  int Cin, kh, kw, pad_t, pad_l;
  int TW, TH, NB;      // M tile = TW*TH*NB = 128 pixels
  int BN, cout_pad, n_ntiles;
  int tiles_x, tiles_y, tiles_b, total_tiles;
  int stages;
  int act, pool;
};

constexpr int kTmaThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kMaxStages = 8;

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

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&tfull_bar[a]), 1); mbar_init(smem_u32(&tempty_bar[a]), 4); }
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
  const int kblocks = p.kh * p.kw * kchunks;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_ntiles;
        int m = tile / p.n_ntiles;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int tb = m / p.tiles_y;
        const int x0 = tx * p.TW - p.pad_l, y0 = ty * p.TH - p.pad_t, b0 = tb * p.NB, n0 = nt * p.BN;
        for (int tap = 0; tap < p.kh * p.kw; ++tap) {
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_arrive_expect_tx(fb, stage_bytes);
            const uint32_t sa = smem_base + stage * stage_bytes;
            tma_load_4d(sa, &tmA, fb, kc * 64, x0 + kx, y0 + ky, b0);
            tma_load_2d(sa + a_bytes, &tmB, fb, kc * 64, tap * p.cout_pad + n0);
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t bd = umma_desc(sa + a_bytes + k * 32, 16, 1024, 2);
            umma_bf16(d, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int sub = warp & 3;            // TMEM sub-partition this warp may access
    const int r = sub * 32 + lane;       // tile row = pixel index inside the M tile
    const int lx = r % p.TW, ly = (r / p.TW) % p.TH, nb = r / (p.TW * p.TH);
    const int PH = p.OH >> 1, PW = p.OW >> 1;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_ntiles;
      int m = tile / p.n_ntiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int tb = m / p.tiles_y;
      const int ox = tx * p.TW + lx, oy = ty * p.TH + ly, b = tb * p.NB + nb, n0 = nt * p.BN;
      EpiOut eo;
      eo.vec_ok = OUT_F32 ? (p.Cout % 4 == 0) : (p.Cout % 8 == 0);
      size_t pix;
      if (p.pool) {
        eo.valid = (b < p.B) && ((oy >> 1) < PH) && ((ox >> 1) < PW);
        pix = ((size_t)b * PH + (oy >> 1)) * PW + (ox >> 1);
      } else {
        eo.valid = (b < p.B) && (oy < p.OH) && (ox < p.OW);
        pix = ((size_t)b * p.OH + oy) * p.OW + ox;
      }
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      for (int ch = 0; ch < p.BN; ch += 32) {
        uint32_t regs[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + ch, regs);
        tmem_ld_wait();
        EpiOut e2 = eo;
        e2.c_left = p.Cout - n0 - ch;
        if (OUT_F32) e2.ptr = reinterpret_cast<float*>(p.out) + pix * p.Cout + n0 + ch;
        else e2.ptr = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + n0 + ch;
        if (p.pool) epilogue_chunk<true, false>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2);
        else epilogue_chunk<false, OUT_F32>(regs, p.bias + n0 + ch, p.act, lane, p.TW, e2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct TcConvPlan {
  CUtensorMap tmA, tmB;
  TmaConvParams p;
  int grid;
  size_t smem;
};

bool tc_conv_first_supported(const ConvGeom& g) {
  return g.kh == 3 && g.kw == 3 && g.Cin == 3 && g.stride == 1 && g.pad_t == 1 && g.pad_l == 1 && g.OH == g.IH && g.OW == g.IW &&
         g.Cout % 16 == 0 && g.Cout >= 16 && g.Cout <= 256 && (g.pool == 0 || g.pool == 2);
}

int tc_conv_first(const fld_handle* h, const void* in, int in_dtype, const __nv_bfloat16* w_packed, const float* bias,
                  __nv_bfloat16* out, const ConvGeom& g, int B, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  FirstParams p;
  p.in = in; p.w = w_packed; p.bias = bias; p.out = out;
  p.B = B; p.H = g.IH; p.W = g.IW; p.Cout = g.Cout; p.act = g.act; p.pool = g.pool;
  p.tiles_x = fld_div_up(g.OW, 8); p.tiles_y = fld_div_up(g.OH, 16);
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  p.in_scale_unused = 0.f;
  const int esz = in_dtype == FLD_U8 ? 1 : 4;
  const size_t smem = 8192 + (size_t)g.Cout * 64 + 18 * 32 * esz + 1024;
  const int ncols = g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : g.Cout <= 128 ? 128 : 256;
  const int cta_per_sm = std::max(1, std::min(512 / ncols, 8));
  const int grid = std::min(p.n_tiles, h->sm_count * cta_per_sm);
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

bool tc_conv_supported(const ConvGeom& g) {
  if (g.stride != 1 || g.Cin % 64 != 0 || g.Cin < 64) return false;
  if (g.pool != 0 && g.pool != 2) return false;
  if (g.pool == 2 && (g.Cout % 8 != 0)) return false;
  if (g.kh * g.kw > 64) return false;
  return true;
}

int tc_conv_plan_create(const fld_handle* h, const void* in, const __nv_bfloat16* w_packed, int cout_pad, const ConvGeom& g, int B,
                        TcConvPlan** out) {
  if (!h->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  TcConvPlan* pl = new TcConvPlan();
  TmaConvParams& p = pl->p;
  p.bias = nullptr; p.out = nullptr;
  p.B = B; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout;
  p.Cin = g.Cin; p.kh = g.kh; p.kw = g.kw; p.pad_t = g.pad_t; p.pad_l = g.pad_l;
  p.act = g.act; p.pool = g.pool;
  // M-tile geometry: TW*TH*NB = 128, TW in {4, 8, 16}; prefer 8x16 (pool partner lane^8)
  int TW, TH, NB;
  if (g.OW >= 16 && g.OW % 16 == 0 && g.OW % 8 != 0) { TW = 16; }
  else if (g.OW > 4) TW = 8;
  else TW = 4;
  const int rows_left = 128 / TW;
  TH = 1;
  while (TH * 2 <= rows_left && TH < g.OH) TH *= 2;
  if (TH < 2) TH = 2;
  NB = 128 / (TW * TH);
  p.TW = TW; p.TH = TH; p.NB = NB;
  // N tile
  int BN = cout_pad <= 256 ? cout_pad : ((cout_pad % 256 == 0) ? 256 : 128);
  if (cout_pad % BN != 0 || BN % 16 != 0) { delete pl; fld_set_error("tc_conv: bad cout_pad %d", cout_pad); return FLD_ERR_INVALID; }
  p.BN = BN; p.cout_pad = cout_pad; p.n_ntiles = cout_pad / BN;
  p.tiles_x = fld_div_up(g.OW, TW); p.tiles_y = fld_div_up(g.OH, TH); p.tiles_b = fld_div_up(B, NB);
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.n_ntiles;
  const size_t stage_bytes = 128 * 128 + (size_t)BN * 128;
  int stages = (int)((200 * 1024) / stage_bytes);
  stages = std::max(2, std::min(stages, kMaxStages));
  p.stages = stages;
  pl->smem = stages * stage_bytes + 1024;
  pl->grid = std::min(p.total_tiles, h->sm_count);

  // activations: bf16 NHWC [B][IH][IW][Cin]
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.IW, (cuuint64_t)g.IH, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.IW * g.Cin * 2, (cuuint64_t)g.IH * g.IW * g.Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)NB};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  // weights: bf16 [taps*cout_pad][Cin]
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.Cin, (cuuint64_t)g.kh * g.kw * cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)g.Cin * 2};
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

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

int tc_conv_run(const TcConvPlan* pl, const float* bias, void* out, int out_dtype, cudaStream_t st) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  TmaConvParams p = pl->p;
  p.bias = bias; p.out = out;
  if (p.pool && out_dtype != FLD_BF16) { fld_set_error("tc_conv: fused pool needs bf16 output"); return FLD_ERR_INVALID; }
  if (out_dtype == FLD_F32) {
    FLD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_tma_kernel<true><<<pl->grid, kTmaThreads, pl->smem, st>>>(pl->tmA, pl->tmB, p);
  } else {
    FLD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    conv_tma_kernel<false><<<pl->grid, kTmaThreads, pl->smem, st>>>(pl->tmA, pl->tmB, p);
  }
  FLD_LAUNCHED();
  return FLD_OK;
}
