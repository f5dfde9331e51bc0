// tc_common.cuh — PTX wrappers (mbarrier / TMA / tcgen05 / TMEM), UMMA descriptors and the fused
// bias + activation + 2x2 max-pool epilogue shared by the tensor-core conv kernels.
#pragma once
#include <cuda.h>
#include "ops.cuh"

namespace tc {

// cuTensorMapEncodeTiled, resolved once per handle through cudaGetDriverEntryPoint (api.cu: fld_handle::encode_tiled)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
__device__ __forceinline__ uint32_t min_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
// one lane of the (converged) warp; lets the compiler keep the enclosing loop on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

template <bool POOL, bool OUT_F32, bool BIAS = true>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&acc)[32], const float* __restrict__ bias, int act, int lane, int TW,
                                               const EpiOut& o) {
  // bias: 8 x 16-byte loads (the pointer is 64-byte aligned: n0 and the chunk offset are multiples of 16 floats)
  float v[32];
  if (BIAS) {
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(b4 + q);
      v[4 * q + 0] = __uint_as_float(acc[4 * q + 0]) + b.x;
      v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b.y;
      v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b.z;
      v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b.w;
    }
  } else {  // bias already accumulated by the tensor core (first layer: spare K slots)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (!POOL) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_f(v[j], act);
  }
  if (POOL) {
    // bf16 output only.  Round first, then max on packed pairs, then the activation on the 4 pooled words: rounding,
    // ReLU / ReLU6 and max are all monotone, so act(max(round(x))) == round(max(act(x))) — 4 ops instead of 32.
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
    if (act != FLD_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        k2[i] = max_bf16x2(k2[i], 0u);
        if (act == FLD_ACT_RELU6) k2[i] = min_bf16x2(k2[i], 0x40c040c0u);  // bf16(6.0) = 0x40c0
      }
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

// ------------------------------------------------------------------------------------------------
// FLD_BF16X3 mode: activations travel as SPLIT tensors — per pixel 2*C bf16 values [hi(C) | lo(C)] with
// hi = bf16(v), lo = bf16(v - hi) — so that the next tensor-core conv can form x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// (fp32 accumulation in TMEM; relative error ~2^-17 per product instead of bf16's 2^-9).
// Epilogue for one 32-column chunk with a split store.  o.ptr points at channel (n0 + chunk*32) of the hi half of this
// thread's (pooled) output pixel; the lo half lives `cout` elements further.  Pooling runs on the fp32 values (two rounds
// of half-exchanges with the pool partners lane^1 / lane^TW: 16 + 8 shuffles), the activation after it (monotone).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store8(__nv_bfloat16* hi_ptr, int cout, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const float r0 = v[2 * i] - __bfloat162float(h0), r1 = v[2 * i + 1] - __bfloat162float(h1);
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = pack_bf16(r0, r1);
  }
  *reinterpret_cast<uint4*>(hi_ptr) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(hi_ptr + cout) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool POOL, bool BIAS = true>
__device__ __forceinline__ void epilogue_chunk_split(uint32_t (&acc)[32], const float* __restrict__ bias, int act, int lane, int TW,
                                                     const EpiOut& o, int cout) {
  float v[32];
  if (BIAS) {
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(b4 + q);
      v[4 * q + 0] = __uint_as_float(acc[4 * q + 0]) + b.x;
      v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b.y;
      v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b.z;
      v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(o.ptr);
  if (POOL) {
    const bool bx = lane & 1;
    float k1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {    // x partner: this lane keeps channels [16*bx, 16*bx + 16)
      const float send = bx ? v[i] : v[16 + i];
      const float keep = bx ? v[16 + i] : v[i];
      k1[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
    }
    const bool by = (lane & TW) != 0;
    float k2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {     // y partner: keeps [8*by, 8*by + 8) of those
      const float send = by ? k1[i] : k1[8 + i];
      const float keep = by ? k1[8 + i] : k1[i];
      k2[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, TW));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) k2[i] = act_f(k2[i], act);
    const int cb = (bx ? 16 : 0) + (by ? 8 : 0);
    if (o.valid && cb + 8 <= o.c_left) split_store8(p + cb, cout, k2);
  } else {
    if (!o.valid) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_f(v[j], act);
    if (o.c_left >= 32 && o.vec_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) split_store8(p + 8 * q, cout, v + 8 * q);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < o.c_left) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
          p[j] = h;
          p[cout + j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
        }
    }
  }
}

}  // namespace tc
