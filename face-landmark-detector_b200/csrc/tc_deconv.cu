// tc_deconv.cu — Conv2DTranspose(k = 2*stride, 'valid', no bias) on the tcgen05 tensor cores, with the per-pixel decode
// of the segmentation head fused into the epilogue.
//
//   reference: networks/fcn.py:104,114,121 (4x4 s2, 4x4 s2, 16x16 s8 transposed convs), networks/utils.py:28-30 (softmax
//   over the classes), prediction.py:209 (argmax class map).
//
// A transposed conv with k = 2s is s*s independent 2x2 "phase" convolutions over the (h+1) x (w+1) grid of input
// neighbourhoods:
//   out[b, oy*s + a, ox*s + bq, o] = sum_{u,v in {0,1}} sum_c in[b, oy-1+u, ox-1+v, c] * W[a + s(1-u), bq + s(1-v), o, c]
// so the whole layer is ONE GEMM   D[M, N] = A[M, K] * Bw[N, K]^T   with
//   M = B*(h+1)*(w+1) neighbourhoods (flat, no per-image tile padding), K = 4*C (taps packed densely, padded to 64),
//   N = s*s phases x cpp (Cout padded to 8).
// A is materialised once by a small im2col kernel (bf16, <= 0.6 MB per image for up8 — 4 % of the output bytes).
//
// Kernel structure (one persistent CTA per SM, 320 threads):
//   warp 0   TMA producer: the M tile's whole A block (all K, 80 KB for up8) is loaded ONCE and stays resident while the CTA
//            walks a run of N tiles (phase pairs) whose weight slabs stream through a ring (L2-resident, 92 KB per tile).
//            Operand fill, not the tensor pipe, bounded the earlier tap-by-tap version (320 KB of smem fill per tile).
//            N is the fast index on purpose: all s*s phases of a pixel block are written within microseconds of each
//            other, so L2 merges the interleaved Cout*4-byte runs into whole output rows before they reach HBM (with N
//            slow, HBM saw isolated 544-byte runs at a 2176-byte pitch and delivered 1.4 TB/s).
//   warp 1   MMA issuer: tcgen05.mma M=128, N=2*cpp, fp32 accumulators in TMEM, two accumulator buffers.
//   warps 2-9 epilogue: each half (4 warps = 128 TMEM lanes) owns one phase of the N tile, so a thread holds ALL Cout
//            logits of one output pixel: softmax / argmax need no cross-thread traffic.  Values leave through a shared
//            memory transpose so that global stores are contiguous Cout*4-byte runs (272 B for 68 classes).
#include "tc_common.cuh"

#include <algorithm>
#include <vector>

namespace {

using namespace tc;

struct DeconvParams {
  void* out;
  long long M;          // rows = B * GH * GW
  int GH, GW;           // neighbourhood grid = (h+1, w+1)
  int s, Cout, cpp;     // stride, real / padded channels per phase
  int BN;               // 2 * cpp
  int kblocks;          // weight k-blocks per tile (Kp / 64; FLD_BF16X3: the three product terms' blocks, see tc_deconv_plan_create)
  int a_blocks;         // 64-column slabs of the stationary A block; weight block kb multiplies A slab kb (kb < a_blocks) or kb - a_blocks
  int n_ntiles, mtiles, total_tiles;
  int nsplit, cn, units;  // work unit = (M tile, run of cn consecutive N tiles); units = mtiles * nsplit
  int stages;
  int mode;             // 0 logits, 1 softmax probabilities, 2 int64 argmax class map, 3 per-class soft-centroid sums
  const float* addend;  // fp32 [B][(h+1)s][(w+1)s][Cout] added to the accumulators before the epilogue's decode (FLD_BF16X3 second pass), or null
  float* acc;           // mode 3: [B][Cout][3] fp32 (sum p, sum p*col, sum p*row), accumulated with atomics
  int walk;             // mode 3: 1 = round 1's per-class walk over the staged block (FLD_TC_DECONV_WALK=1), 0 = tf32 tensor-core reduction
  unsigned long long* trace;  // FLD_TC_TRACE: clock64 event log of CTA 0, [3 roles][kTraceN]
};

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

// ---- fused soft centroid on the tensor cores (mode 3).  The per-class sums over a tile's 128 pixels are a tiny GEMM
//   D[class][j] = sum_px P[px][class] * W[px][j],   W = (1, col, row) of the tile's first image | the same for its second image,
// evaluated with tcgen05.mma kind::tf32 (fp32 storage, 10-bit mantissa inputs, fp32 accumulation: the column / row indices are
// exact, the probabilities carry 2^-11 relative error that largely cancels in the ratio).  It replaces a per-class walk of 68
// threads down the staged block (128 x (2 LDS + 3 FMA) per thread per tile: ~60 % of the epilogue's instructions).
// A = P^T, K-major no-swizzle: element (class c, pixel k) at (k / 4) * kRedLbo + (c / 8) * 128 + (c % 8) * 16 + (k % 4) * 4.
constexpr int kRedCols = 160;                                     // TMEM columns [160, 192): two 16-column reduction accumulators
__host__ __device__ constexpr int red_lbo(int cout) { return ((cout + 7) / 8) * 128 + 16; }   // +16: skews the K groups over the banks
__host__ __device__ constexpr int red_a_bytes(int cout) { return 32 * red_lbo(cout); }        // 32 K groups of 4 pixels
// B: ten 16-byte rows j of [4 px] fp32 per K group.  The N = 16 operand's second 8-row block starts 32 B after the first (SBO = 32),
// so operand rows 8..15 alias storage rows 2..9: D columns 0..7 are rows 0..7, D columns 14, 15 are rows 8, 9.
constexpr int kRedBLbo = 176;                                     // 160 B of rows + 16 B bank skew
constexpr int kRedBBytes = 32 * kRedBLbo;
// 16-warp variant: B = six rows (1, col, row | 1, col, row) of centred integers; operand rows 8..15 alias rows 0..7 (SBO = 0)
constexpr int kRedB16Lbo = 144;                                   // 128 B of rows + 16 B bank skew
constexpr int kRedB16Bytes = 32 * kRedB16Lbo;
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8v(uint32_t taddr, uint32_t* r) { tmem_ld8(taddr, r); }
constexpr int kMaxStages = 8;
constexpr int kTraceN = 2048;

#define DTRACE(role, idx, tag)                                                                                        \
  do {                                                                                                                \
    if (p.trace && blockIdx.x == 0 && (idx) < kTraceN) p.trace[(role) * kTraceN + (idx)++] = ((unsigned long long)clock64() << 4) | (tag); \
  } while (0)

template <int COUT, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
deconv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DeconvParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2], afull_bar, afree_bar, red_bar[2];
  __shared__ long long goff[2][128];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * 128;
  const uint32_t smem_a = smem_base;                                   // stationary A block: kblocks slabs [128][64] (SW128)
  const uint32_t smem_b = smem_a + (uint32_t)p.a_blocks * a_bytes;     // weight ring: slabs [BN][64] (b_bytes multiple of 1024)
  const uint32_t smem_stg = smem_b + (uint32_t)p.stages * b_bytes;     // 2 x [128][Cout] fp32 staging
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
  const uint32_t afull = smem_u32(&afull_bar), afree = smem_u32(&afree_bar);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, EW); }
    mbar_init(afull, 1);
    mbar_init(afree, 1);
    mbar_init(smem_u32(&red_bar[0]), 1);
    mbar_init(smem_u32(&red_bar[1]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, afree_phase = 0;
      bool first = true;
      int ti = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int mt = u / p.nsplit, nc = u - mt * p.nsplit;
        if (!first) { mbar_wait(afree, afree_phase); afree_phase ^= 1; }          // MMAs reading the previous A block are done
        first = false;
        DTRACE(0, ti, 3);
        mbar_arrive_expect_tx(afull, (uint32_t)p.a_blocks * a_bytes);
        for (int kb = 0; kb < p.a_blocks; ++kb) tma_load_2d(smem_a + kb * a_bytes, &tmA, afull, kb * 64, mt * 128);  // rows past M: zero fill
        for (int nt = nc * p.cn; nt < (nc + 1) * p.cn; ++nt)
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            DTRACE(0, ti, 1);
            const uint32_t fb = full0 + 8 * stage;
            mbar_arrive_expect_tx(fb, b_bytes);
            tma_load_2d(smem_b + stage * b_bytes, &tmB, fb, kb * 64, nt * p.BN);
            DTRACE(0, ti, 2);
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      const uint64_t adesc0 = umma_desc(smem_a, 16, 1024, 2);
      const uint64_t bdesc0 = umma_desc(smem_b, 16, 1024, 2);
      const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, aphase = 0;
      int ti = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(afull, aphase);
        DTRACE(1, ti, 4);
        aphase ^= 1;
        for (int t = 0; t < p.cn; ++t) {
          mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
          DTRACE(1, ti, 3);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 256;
          uint32_t accum = 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(full0 + 8 * stage, phase);
            DTRACE(1, ti, 1);
            tc_fence_after();
            const uint64_t ad = adesc0 + (uint64_t)((kb < p.a_blocks ? kb : kb - p.a_blocks) * a_step);
            const uint64_t bd = bdesc0 + (uint64_t)(stage * b_step);
            umma_bf16(d, ad, bd, idesc, accum);
            umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
            umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
            umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
            umma_commit(empty0 + 8 * stage);
            if (kb == p.kblocks - 1) {
              umma_commit(tfull0 + 8 * acc);
              if (t == p.cn - 1) umma_commit(afree);
            }
            DTRACE(1, ti, 2);
            accum = 1u;
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    if constexpr (EW == 16) {
      // ---- soft-centroid epilogue with SIXTEEN warps (mode 3, 68 classes).  Two threads share a pixel: thread ch = 0 holds
      // classes [0, 32), ch = 1 classes [32, 68), so a thread carries 32..36 logits instead of 68 and every scheduler has four
      // epilogue warps to hide the MUFU / TMEM / shared-memory latencies behind (with eight warps the epilogue ran at 0.27
      // instructions per clock per scheduler and bounded the kernel: 3600 clk per tile-phase against 840 clk of MMA issue).
      // Each thread exponentiates against ITS OWN maximum; the pair exchanges (max, sum) through four unused rows of the
      // staging block and rescales, p = e * 2^(m_own - m) / (s_own 2^(m_own - m) + s_other 2^(m_other - m)), as it stages
      // P^T (A operand).  B = (1, col, row) of the tile's first | second image, centred integers: exact in tf32.
      static_assert(COUT == 68, "the 16-warp epilogue splits 68 classes at 32");
      const int ew = warp - 2;
      const int sub = warp & 3;              // TMEM lane quadrant (hardware rule: warp id % 4)
      const int half = (ew >> 2) & 1;        // phase of the N tile
      const int ch = ew >> 3;                // class half
      const int r = sub * 32 + lane;         // tile row = pixel
      const int OWs = p.GW * p.s, OHs = p.GH * p.s;
      const uint32_t lbo = (uint32_t)red_lbo(COUT);
      const uint32_t sA = smem_stg + (uint32_t)half * (uint32_t)red_a_bytes(COUT);
      const uint32_t sB = smem_stg + 2u * (uint32_t)red_a_bytes(COUT) + (uint32_t)half * (uint32_t)kRedB16Bytes;
      const uint32_t a_thr = sA + (uint32_t)(r >> 2) * lbo + (uint32_t)(r & 3) * 4u;
      const uint32_t st_own = a_thr + 8u * 128u + (uint32_t)(4 + 2 * ch) * 16u;          // "classes" 68 + 2 ch, 69 + 2 ch: never read back
      const uint32_t st_oth = a_thr + 8u * 128u + (uint32_t)(4 + 2 * (1 - ch)) * 16u;
      const uint32_t b_thr = sB + (uint32_t)(r >> 2) * (uint32_t)kRedB16Lbo + (uint32_t)(r & 3) * 4u;
      const uint32_t rbar = smem_u32(&red_bar[half]);
      const uint32_t dred = tmem_base + kRedCols + (uint32_t)half * 16u;
      const bool issuer = (ch == 0 && sub == 0 && lane == 0);
      uint32_t acc = 0, acc_phase = 0, red_phase = 0;
      bool red_pending = false, red_unit = false;
      int red_mt = 0;
      auto red_read = [&]() {
        if (ch == 0 && sub < 3) {
          tc_fence_after();
          uint32_t q[8];
          tmem_ld8(dred + ((uint32_t)(sub * 32) << 16), q);
          tmem_ld_wait();
          tc_fence_before();
          if (r < COUT) {                                 // TMEM lane r of the reduction accumulator = class r
            const long long g0 = (long long)red_mt * 128;
            const int per = p.GH * p.GW;
            const long long b0 = g0 / per;
            const int nvalid = (int)min((long long)128, p.M - g0);
            const int split = (int)min((long long)nvalid, (b0 + 1) * per - g0);
            float* d0 = p.acc + ((size_t)b0 * COUT + r) * 3;
            atomicAdd(d0, __uint_as_float(q[0])); atomicAdd(d0 + 1, __uint_as_float(q[1])); atomicAdd(d0 + 2, __uint_as_float(q[2]));
            if (split < nvalid) {
              float* d1 = d0 + (size_t)COUT * 3;
              atomicAdd(d1, __uint_as_float(q[3])); atomicAdd(d1 + 1, __uint_as_float(q[4])); atomicAdd(d1 + 2, __uint_as_float(q[5]));
            }
          }
        }
        red_unit = false;
      };
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int mt = u / p.nsplit, nt0 = (u - mt * p.nsplit) * p.cn;
        const long long g = (long long)mt * 128 + r;
        const bool valid = g < p.M;
        const long long b = g / (p.GH * p.GW);
        const int rem = (int)(g - b * (p.GH * p.GW));
        const int oy = rem / p.GW, ox = rem - oy * p.GW;
        const long long b0 = ((long long)mt * 128) / (p.GH * p.GW);   // a tile of 128 flat rows spans at most two images
        const float w0 = (valid && b == b0) ? 1.f : 0.f, w1 = (valid && b != b0) ? 1.f : 0.f;
        int a = (nt0 * 2 + half) / p.s, bq = (nt0 * 2 + half) - a * p.s;
        for (int t = 0; t < p.cn; ++t, bq += 2) {
          if (bq >= p.s) { bq -= p.s; ++a; }
          mbar_wait(tfull0 + 8 * acc, acc_phase);
          tc_fence_after();
          uint32_t rg[40];
          const uint32_t tbase = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + half * p.cpp + (uint32_t)ch * 32u;
          tmem_ld32(tbase, *reinterpret_cast<uint32_t(*)[32]>(&rg[0]));
          if (ch) tmem_ld8(tbase + 32, &rg[32]);          // classes 64..67 (+ 4 padding columns)
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);  // accumulator is in registers: the MMAs of the tile after next may start
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;

          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(rg[j]));
          if (ch) {
#pragma unroll
            for (int j = 32; j < 36; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(rg[j]));
          }
          const float m = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          const float nm = -m * 1.4426950408889634f;
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rg[j]), 1.4426950408889634f, nm)));
            rg[j] = __float_as_uint(e);
            sum[j & 3] += e;
          }
          if (ch) {
#pragma unroll
            for (int j = 32; j < 36; ++j) {
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rg[j]), 1.4426950408889634f, nm)));
              rg[j] = __float_as_uint(e);
              sum[j & 3] += e;
            }
          }
          const float sm = (sum[0] + sum[1]) + (sum[2] + sum[3]);

          if (red_pending) {                              // previous tile's MMAs are done with the staging block
            mbar_wait(rbar, red_phase);
            red_phase ^= 1u;
            red_pending = false;
          }
          if (t == 0 && red_unit) red_read();             // previous unit complete: its sums leave TMEM before they are overwritten
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_own), "f"(m) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_own + 16u), "f"(sm) : "memory");
          if (ch == 0) {
            // grids centred on the map (utils/metrics.py:57-64 grids, shifted back in centroid_finish_kernel): small exact integers
            const float fx = (float)(ox * p.s + bq - OWs / 2), fy = (float)(oy * p.s + a - OHs / 2);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr), "f"(w0) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + 16u), "f"(w0 * fx) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + 32u), "f"(w0 * fy) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + 48u), "f"(w1) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + 64u), "f"(w1 * fx) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + 80u), "f"(w1 * fy) : "memory");
          }
          named_bar_sync(1 + half, 256);                  // the pair's (max, sum) are posted
          float mo, so;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mo) : "r"(st_oth) : "memory");
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(so) : "r"(st_oth + 16u) : "memory");
          const float mm = fmaxf(m, mo);
          float f_own, f_oth;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_own) : "f"((m - mm) * 1.4426950408889634f));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_oth) : "f"((mo - mm) * 1.4426950408889634f));
          const float scale = __fdividef(f_own, fmaf(sm, f_own, so * f_oth));
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = j;                              // ch 0: class j; ch 1: class 32 + j (same offsets, 4 core matrices on)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_thr + (uint32_t)ch * 512u + (uint32_t)((c >> 3) * 128 + (c & 7) * 16)),
                         "f"(__uint_as_float(rg[j]) * scale) : "memory");
          }
          if (ch) {
#pragma unroll
            for (int j = 32; j < 36; ++j)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_thr + 512u + (uint32_t)((j >> 3) * 128 + (j & 7) * 16)), "f"(__uint_as_float(rg[j]) * scale) : "memory");
          }
          fence_async_smem();                             // generic-proxy stores -> tensor-core (async proxy) reads
          named_bar_sync(1 + half, 256);                  // A and B complete; at t == 0 every reader has also consumed the previous D
          if (issuer) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_tf32(128, 16);
            const uint64_t ad = umma_desc(sA, lbo, 128, 0);          // K group = 4 pixels (16 B); 8-class core matrices 128 B apart
            const uint64_t bd = umma_desc(sB, kRedB16Lbo, 0, 0);     // SBO 0: rows 8..15 of the N = 16 operand alias rows 0..7
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2)
              umma_tf32(dred, ad + (uint64_t)(m2 * ((2 * lbo) >> 4)), bd + (uint64_t)(m2 * ((2 * kRedB16Lbo) >> 4)), idesc, (t | m2) ? 1u : 0u);
            umma_commit(rbar);
          }
          red_pending = true;
          red_unit = true;
          red_mt = mt;
        }
      }
      if (red_pending) mbar_wait(rbar, red_phase);
      if (red_unit) red_read();
    } else {
    // COUT > 0: the channel count is a compile-time constant (68 = the reference's n_classes), so none of the unrolled
    // per-channel loops carries a predicate; COUT == 0 is the generic runtime-count variant.
    const int Cout = COUT ? COUT : p.Cout;
    const int ew = warp - 2;
    const int sub = warp & 3;              // TMEM lane quadrant (hardware rule: warp id % 4)
    const int half = ew >> 2;              // phase of the N tile this warp group owns
    const int r = sub * 32 + lane;         // tile row
    const int gt = (ew & 3) * 32 + lane;   // thread index inside the 128-thread group (copy-out role)
    const int n32 = Cout >> 5;             // full 32-column chunks
    const int tail8 = ((Cout & 31) + 7) >> 3;
    const bool vec = (Cout & 3) == 0;
    const uint32_t stg = smem_stg + (uint32_t)half * (uint32_t)(128 * Cout * 4);
    const uint32_t row = stg + (uint32_t)r * (uint32_t)(Cout * 4);
    const int OWs = p.GW * p.s;
    const long long OHs = (long long)p.GH * p.s;
    // copy-out role: piece i = gt + 128*j of the [128 px][nper pieces] staging block; (px, q) advance incrementally
    const int nper = vec ? (Cout >> 2) : Cout;
    const int px0 = gt / nper, q0 = gt - px0 * nper, px_step = 128 / nper, q_step = 128 - px_step * nper;
    float* const outp = reinterpret_cast<float*>(p.out);
    uint32_t acc = 0, acc_phase = 0, red_phase = 0;
    int ti = 0;
    const bool tr = (warp == 2 && lane == 0);
    // Tensor-core soft-centroid reduction (mode 3): the sums of a unit (one M tile = the same 128 pixels under every phase pair)
    // accumulate in TMEM across the unit's tiles; they are read back and added to acc[b][c][:] once per unit, at the first tile
    // of the next unit (a flush per tile cost 88 M same-address atomics per 1024 images).  The MMAs of tile i run while this
    // half loads and exponentiates tile i+1; red_bar only guards the reuse of the staging block.
    const uint32_t rbar = smem_u32(&red_bar[half]);
    const uint32_t dred = tmem_base + kRedCols + (uint32_t)half * 16u;
    bool red_pending = false, red_unit = false;
    int red_mt = 0;
    auto red_read = [&]() {
      tc_fence_after();
      uint32_t q[16];
      tmem_ld16(dred + ((uint32_t)(sub * 32) << 16), q);
      tmem_ld_wait();
      tc_fence_before();
      if (r < Cout) {                                   // TMEM lane r of the reduction accumulator = class r
        const long long g0 = (long long)red_mt * 128;
        const int per = p.GH * p.GW;
        const long long b0 = g0 / per;
        const int nvalid = (int)min((long long)128, p.M - g0);
        const int split = (int)min((long long)nvalid, (b0 + 1) * per - g0);
        float* d0 = p.acc + ((size_t)b0 * Cout + r) * 3;
        atomicAdd(d0, __uint_as_float(q[0]));
        atomicAdd(d0 + 1, __uint_as_float(q[1]) + __uint_as_float(q[2]));
        atomicAdd(d0 + 2, __uint_as_float(q[3]) + __uint_as_float(q[4]));
        if (split < nvalid) {
          float* d1 = d0 + (size_t)Cout * 3;
          atomicAdd(d1, __uint_as_float(q[5]));
          atomicAdd(d1 + 1, __uint_as_float(q[6]) + __uint_as_float(q[7]));
          atomicAdd(d1 + 2, __uint_as_float(q[14]) + __uint_as_float(q[15]));
        }
      }
      red_unit = false;
    };
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
    // the pixel this thread owns is fixed over the unit's run of phase pairs: the divisions are done once per unit
    const int mt = u / p.nsplit, nt0 = (u - mt * p.nsplit) * p.cn;
    const long long g = (long long)mt * 128 + r;
    const bool valid = g < p.M;
    const long long b = g / (p.GH * p.GW);
    const int rem = (int)(g - b * (p.GH * p.GW));
    const int oy = rem / p.GW, ox = rem - oy * p.GW;
    const long long b0 = ((long long)mt * 128) / (p.GH * p.GW);   // a tile of 128 flat rows spans at most two images
    const float w0 = (valid && b == b0) ? 1.f : 0.f, w1 = (valid && b != b0) ? 1.f : 0.f;
    int a = (nt0 * 2 + half) / p.s, bq = (nt0 * 2 + half) - a * p.s;
    for (int t = 0; t < p.cn; ++t, bq += 2) {
      if (bq >= p.s) { bq -= p.s; ++a; }               // phase index nt*2 + half advances by 2 per tile (s >= 2)
      const long long opix = (b * OHs + (long long)oy * p.s + a) * OWs + (long long)ox * p.s + bq;

      if (tr) DTRACE(2, ti, 0);
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      if (tr) DTRACE(2, ti, 1);
      tc_fence_after();
      uint32_t rg[3][32];
      const uint32_t tbase = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 256 + half * p.cpp;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < n32) tmem_ld32(tbase + k * 32, rg[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k == n32) {
#pragma unroll
          for (int t8 = 0; t8 < 4; ++t8)
            if (t8 < tail8) tmem_ld8(tbase + k * 32 + t8 * 8, &rg[k][t8 * 8]);
        }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);   // accumulator is in registers: the MMAs of the tile after next may start
      if (tr) DTRACE(2, ti, 2);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;

      if (p.mode == 2) {
        float mx = -INFINITY;
        int amax = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k * 32 + j < Cout) {
              const float v = __uint_as_float(rg[k][j]);
              if (v > mx) { mx = v; amax = k * 32 + j; }   // first maximum wins (numpy argmax, prediction.py:209)
            }
        if (valid) reinterpret_cast<long long*>(p.out)[opix] = amax;
        continue;
      }

      float inv = 1.0f;
      if (p.addend != nullptr && valid) {
        // FLD_BF16X3 second pass: the first pass's logits (x_hi w_hi + x_hi w_lo) join this pass's x_lo w_hi before the decode
        const float* ad = p.addend + opix * Cout;
        if (vec) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (k * 32 + 4 * q + 4 <= Cout) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(ad) + k * 8 + q);
                rg[k][4 * q] = __float_as_uint(__uint_as_float(rg[k][4 * q]) + a4.x);
                rg[k][4 * q + 1] = __float_as_uint(__uint_as_float(rg[k][4 * q + 1]) + a4.y);
                rg[k][4 * q + 2] = __float_as_uint(__uint_as_float(rg[k][4 * q + 2]) + a4.z);
                rg[k][4 * q + 3] = __float_as_uint(__uint_as_float(rg[k][4 * q + 3]) + a4.w);
              }
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k * 32 + j < Cout) rg[k][j] = __float_as_uint(__uint_as_float(rg[k][j]) + __ldg(ad + k * 32 + j));
        }
      }
      if (p.mode == 1 || p.mode == 3) {
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four chains: the warps of a half hide little latency
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k * 32 + j < Cout) mx4[(j >> 1) & 3] = fmaxf(mx4[(j >> 1) & 3], __uint_as_float(rg[k][j]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // exp(v - mx) = 2^(v*log2e - mx*log2e): one FFMA + one MUFU.EX2 per class (flush-to-zero: terms below 2^-126 add nothing)
        const float nmx = -mx * 1.4426950408889634f;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains instead of one 68-deep dependent FADD chain
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k * 32 + j < Cout) {
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rg[k][j]), 1.4426950408889634f, nmx)));
              rg[k][j] = __float_as_uint(e);
              sum[j & 3] += e;
            }
        inv = __fdividef(1.0f, (sum[0] + sum[1]) + (sum[2] + sum[3]));
      }
      if (p.mode == 3 && !p.walk) {
        // ---- fused soft centroid on the tensor cores (see the top of the file): stage E^T (A: the un-normalised exponentials)
        // and the pixel weights (B: 1/sum folded in) of this tile; one thread issues 16 tf32 MMAs (K = 8 pixels each).
        if (red_pending) {                               // previous tile's MMAs are done with the staging block
          mbar_wait(rbar, red_phase);
          red_phase ^= 1u;
          red_pending = false;
        }
        if (t == 0 && red_unit) red_read();              // previous unit complete: its sums leave TMEM before they are overwritten
        const uint32_t lbo = (uint32_t)red_lbo(Cout);
        const uint32_t sA = smem_stg + (uint32_t)half * (uint32_t)red_a_bytes(Cout);
        const uint32_t sB = smem_stg + 2u * (uint32_t)red_a_bytes(Cout) + (uint32_t)half * (uint32_t)kRedBBytes;
        const uint32_t a_thr = sA + (uint32_t)(r >> 2) * lbo + (uint32_t)(r & 3) * 4u;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k * 32 + j < Cout) {
              const int c = k * 32 + j;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_thr + (uint32_t)((c >> 3) * 128 + (c & 7) * 16)), "r"(rg[k][j]) : "memory");
            }
        // B rows (1/sum, col hi, col lo, row hi, row lo) x (first, second image of the tile).  Everything the tensor core sees is
        // exact in tf32: 1/sum is cut to 11 bits, its product with an (integer < 2^13) coordinate splits into an 11-bit head and
        // an exact tail, so the numerators and the denominator carry the SAME per-pixel weight (utils/metrics.py:57-64 grids).
        const float it = valid ? __uint_as_float(__float_as_uint(inv) & 0xffffe000u) : 0.f;
        // grids centred on the map (signed 8..12-bit integers): the running sums stay small, which keeps the tensor core's fp32
        // accumulation (products are aligned to the accumulator and cut, not rounded) well below the probabilities' own error
        const float px_ = it * (float)(ox * p.s + bq - OWs / 2), py_ = it * (float)(oy * p.s + a - (int)(OHs / 2));
        const float hx = __uint_as_float(__float_as_uint(px_) & 0xffffe000u), hy = __uint_as_float(__float_as_uint(py_) & 0xffffe000u);
        const uint32_t b_thr = sB + (uint32_t)(r >> 2) * (uint32_t)kRedBLbo + (uint32_t)(r & 3) * 4u;
        const uint32_t own = (b != b0) ? 5u : 0u, oth = 5u - own;      // first row of this pixel's image / of the other image
        const float vals[5] = {it, hx, px_ - hx, hy, py_ - hy};
#pragma unroll
        for (uint32_t j = 0; j < 5; ++j) {
          const uint32_t ro = own + j, rz = oth + j;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + ro * 16u), "f"(vals[j]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_thr + rz * 16u), "f"(0.f) : "memory");
        }
        fence_async_smem();                              // generic-proxy stores -> tensor-core (async proxy) reads
        named_bar_sync(1 + half, 128);                   // A and B complete; at t == 0 every thread has also read the previous D
        if (gt == 0) {
          tc_fence_after();
          const uint32_t idesc = umma_idesc_tf32(128, 16);
          const uint64_t ad = umma_desc(sA, lbo, 128, 0);          // K group = 4 pixels (16 B); 8-class core matrices 128 B apart
          const uint64_t bd = umma_desc(sB, kRedBLbo, 32, 0);      // overlapping 8-row blocks, see kRedBLbo
#pragma unroll
          for (int m2 = 0; m2 < 16; ++m2)
            umma_tf32(dred, ad + (uint64_t)(m2 * ((2 * lbo) >> 4)), bd + (uint64_t)(m2 * ((2 * kRedBLbo) >> 4)), idesc, (t | m2) ? 1u : 0u);
          umma_commit(rbar);
        }
        red_pending = true;
        red_unit = true;
        red_mt = mt;
        continue;
      }
      named_bar_sync(1 + half, 128);                   // previous tile's copy-out has drained the staging buffer
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = k * 32 + 4 * q;
          if (c + 4 <= Cout && vec) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + c * 4), "f"(__uint_as_float(rg[k][4 * q]) * inv),
                         "f"(__uint_as_float(rg[k][4 * q + 1]) * inv), "f"(__uint_as_float(rg[k][4 * q + 2]) * inv),
                         "f"(__uint_as_float(rg[k][4 * q + 3]) * inv)
                         : "memory");
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c + e < Cout) asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + (c + e) * 4), "f"(__uint_as_float(rg[k][4 * q + e]) * inv) : "memory");
          }
        }
      if (p.mode == 3) {
        // pixel meta for the reduction: (float column, float row) of this thread's output pixel (utils/metrics.py:57-64 grids)
        const float2 xyf = make_float2((float)(ox * p.s + bq), (float)(oy * p.s + a));
        goff[half][r] = *reinterpret_cast<const long long*>(&xyf);
      } else {
        goff[half][r] = valid ? opix * Cout : -1;
      }
      named_bar_sync(1 + half, 128);
      if (p.mode == 3) {
        // Fused soft centroid (get_average_xy with n_points < 1 on the softmax output): thread c sums class c over the tile's
        // pixels (conflict-free column walk through the staging block) and adds the three sums to acc[b][c][:].  A tile of 128
        // flat rows spans at most two images (runs [0, split) and [split, nvalid)).  The probabilities never leave the SM.
        if (gt < Cout) {
          const float* sp = reinterpret_cast<const float*>(__cvta_shared_to_generic((size_t)stg)) + gt;   // column gt of [128][Cout]
          const long long g0 = (long long)mt * 128;
          const int per = p.GH * p.GW;
          const long long b0 = g0 / per;
          const int nvalid = (int)min((long long)128, p.M - g0);
          const int split = (int)min((long long)nvalid, (b0 + 1) * per - g0);
          int lo = 0, hi = split;
#pragma unroll 1
          for (int run = 0; run < 2; ++run) {
            if (hi > lo) {
              float s0[2] = {0.f, 0.f}, sx[2] = {0.f, 0.f}, sy[2] = {0.f, 0.f};
              int px = lo;
              for (; px + 8 <= hi; px += 8) {
                float v[8];
                float2 m[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  v[e] = sp[(px + e) * Cout];
                  m[e] = *reinterpret_cast<const float2*>(&goff[half][px + e]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  s0[e & 1] += v[e];
                  sx[e & 1] = fmaf(v[e], m[e].x, sx[e & 1]);
                  sy[e & 1] = fmaf(v[e], m[e].y, sy[e & 1]);
                }
              }
              for (; px < hi; ++px) {
                const float v = sp[px * Cout];
                const float2 m = *reinterpret_cast<const float2*>(&goff[half][px]);
                s0[0] += v; sx[0] = fmaf(v, m.x, sx[0]); sy[0] = fmaf(v, m.y, sy[0]);
              }
              float* d = p.acc + ((size_t)(b0 + run) * Cout + gt) * 3;
              atomicAdd(d, s0[0] + s0[1]); atomicAdd(d + 1, sx[0] + sx[1]); atomicAdd(d + 2, sy[0] + sy[1]);
            }
            lo = split; hi = nvalid;
          }
        }
        continue;
      }
      // copy-out: consecutive threads write consecutive pieces of each pixel's contiguous Cout*4-byte run
      int px = px0, q = q0;
      if (COUT && vec) {
        constexpr int NPER = COUT ? COUT / 4 : 1;
        float4 v[NPER];
        long long off[NPER];
#pragma unroll
        for (int j = 0; j < NPER; ++j) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                       : "r"(stg + (uint32_t)(gt + 128 * j) * 16u));
          const long long o = goff[half][px];
          off[j] = o < 0 ? -1 : o + 4 * q;
          q += q_step; px += px_step;
          if (q >= nper) { q -= nper; ++px; }
        }
#pragma unroll
        for (int j = 0; j < NPER; ++j)
          if (off[j] >= 0) *reinterpret_cast<float4*>(outp + off[j]) = v[j];
      } else if (vec) {
        for (int j = 0; j < nper; ++j) {
          const long long o = goff[half][px];
          if (o >= 0) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(stg + (uint32_t)(gt + 128 * j) * 16u));
            *reinterpret_cast<float4*>(outp + o + 4 * q) = v;
          }
          q += q_step; px += px_step;
          if (q >= nper) { q -= nper; ++px; }
        }
      } else {
        for (int j = 0; j < nper; ++j) {
          const long long o = goff[half][px];
          if (o >= 0) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stg + (uint32_t)(gt + 128 * j) * 4u));
            outp[o + q] = v;
          }
          q += q_step; px += px_step;
          if (q >= nper) { q -= nper; ++px; }
        }
      }
    }
    }
    if (red_pending) mbar_wait(rbar, red_phase);
    if (red_unit) red_read();
    }  // EW == 8
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// A[g][(u*2+v)*C + c] = in[b][oy-1+u][ox-1+v][c] (zero outside the map / past 4*C), bf16; one thread per two K entries
__global__ void deconv_im2col_kernel(const float* __restrict__ in, __nv_bfloat162* __restrict__ A, long long M, int h, int w, int C, int Kp) {
  const int kp2 = Kp >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * kp2) return;
  const long long g = i / kp2;
  const int k0 = (int)(i - g * kp2) * 2;
  const int GW = w + 1, GH = h + 1;
  const long long b = g / (GH * GW);
  const int rem = (int)(g - b * (GH * GW));
  const int oy = rem / GW, ox = rem - oy * GW;
  float v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = k0 + e;
    const int t = k / C, c = k - t * C;
    const int iy = oy - 1 + (t >> 1), ix = ox - 1 + (t & 1);
    v[e] = (t < 4 && iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(in + ((b * h + iy) * w + ix) * C + c) : 0.f;
  }
  A[i] = __floats2bfloat162_rn(v[0], v[1]);
}

// C % 4 == 0: one warp per row of A; lane j handles the 8-byte chunk j (4 channels of one tap): a 16-byte load, an 8-byte store,
// the row's index arithmetic once per warp.  (The element-wise kernel above ran at 0.5 TB/s: 0.26 ms per 256 images for up8.)
__global__ void deconv_im2col_vec4_kernel(const float4* __restrict__ in, uint2* __restrict__ A, long long M, int h, int w, int C4, int Kp4) {
  const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= M) return;
  const int lane = threadIdx.x & 31;
  const int GW = w + 1, GH = h + 1;
  const long long b = g / (GH * GW);
  const int rem = (int)(g - b * (GH * GW));
  const int oy = rem / GW, ox = rem - oy * GW;
  uint2* row = A + g * Kp4;
  for (int j = lane; j < Kp4; j += 32) {
    const int t = j / C4, q = j - t * C4;
    const int iy = oy - 1 + (t >> 1), ix = ox - 1 + (t & 1);
    uint2 o = make_uint2(0u, 0u);
    if (t < 4 && iy >= 0 && iy < h && ix >= 0 && ix < w) {
      const float4 v = __ldg(in + ((b * h + iy) * w + ix) * C4 + q);
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
    }
    row[j] = o;
  }
}

// FLD_BF16X3: A row = [hi(4C) | lo(4C) | 0 ...] with hi = bf16(v), lo = bf16(v - hi); one thread per K entry pair of a row
// part 0: [hi | lo] (class-map variant); part 1: the lo parts alone in the standard K layout (second pass of the two-pass variant)
__global__ void deconv_im2col_x3_kernel(const float* __restrict__ in, __nv_bfloat162* __restrict__ A, long long M, int h, int w, int C, int KpA,
                                        int part) {
  const int kp2 = KpA >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * kp2) return;
  const long long g = i / kp2;
  const int k0 = (int)(i - g * kp2) * 2;
  const int GW = w + 1, GH = h + 1;
  const long long b = g / (GH * GW);
  const int rem = (int)(g - b * (GH * GW));
  const int oy = rem / GW, ox = rem - oy * GW;
  float v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    int k = k0 + e;
    const bool lo = part == 1 || k >= 4 * C;
    if (part == 0 && lo) k -= 4 * C;
    const int t = k / C, c = k - t * C;
    const int iy = oy - 1 + (t >> 1), ix = ox - 1 + (t & 1);
    const float x = (t < 4 && iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(in + ((b * h + iy) * w + ix) * C + c) : 0.f;
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    v[e] = lo ? x - hi : hi;
  }
  A[i] = __floats2bfloat162_rn(v[0], v[1]);
}

// (x, y) = (sum p*col / sum p, sum p*row / sum p); (-1, -1) when mean(p) <= thresh (utils/metrics.py:78-80)
// (cx, cy): the origin the sums were taken about (the tensor-core reduction centres the grids, see the epilogue)
__global__ void centroid_finish_kernel(const float* __restrict__ acc, long long n, double hw, double thresh, double cx, double cy,
                                       double* __restrict__ xy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s0 = acc[i * 3], sx = acc[i * 3 + 1], sy = acc[i * 3 + 2];
  double x = cx + sx / s0, y = cy + sy / s0;
  if (s0 / hw <= thresh) { x = -1.0; y = -1.0; }
  xy[i * 2] = x;
  xy[i * 2 + 1] = y;
}


constexpr size_t kSmemMax = 232448 - 4096;  // 227 KB per CTA minus the static part (barriers, goff table)

size_t fixed_smem(int Cin, int Cout) {
  const int Kp = tc_deconv_kp(Cin), cpp = tc_deconv_cpp(Cout);
  (void)cpp;
  // stationary A block + staging (modes 0 / 1: 2 x [128][Cout] fp32; mode 3: 2 x E^T + 2 x pixel weights; the MMA over-reads of A land inside the block) + alignment slack
  const size_t stg = std::max((size_t)2 * 128 * Cout * 4, (size_t)2 * red_a_bytes(Cout) + 2 * kRedBBytes);
  return (size_t)(Kp / 64) * 16384 + stg + 1024;
}

}  // namespace

int tc_deconv_kp(int Cin) { return (4 * Cin + 63) / 64 * 64; }
int tc_deconv_cpp(int Cout) {
  const char* e = getenv("FLD_TC_DECONV_CPP");   // bring-up switch: 96 reproduces the 32-column-chunk layout
  const int want = e ? atoi(e) : 0;
  const int cpp = (Cout + 7) / 8 * 8;
  return want >= cpp && want % 8 == 0 ? want : cpp;
}

bool tc_deconv_supported(int k, int s, int Cin, int Cout) {
  if (k != 2 * s || s < 2 || (s * s) % 2 != 0 || Cout < 1 || Cout > 96 || Cin < 1) return false;
  if (2 * tc_deconv_cpp(Cout) > 256) return false;
  return fixed_smem(Cin, Cout) + (size_t)2 * (2 * tc_deconv_cpp(Cout)) * 128 <= kSmemMax;
}

size_t tc_deconv_scratch_bytes(int B, int IH, int IW, int Cin) { return (size_t)B * (IH + 1) * (IW + 1) * tc_deconv_kp(Cin) * 2; }

// FLD_BF16X3 (class-map mode only: no staging fits beside the larger A block).  The three product terms x_hi w_hi + x_lo w_hi +
// x_hi w_lo run in ONE accumulation: A row = [x_hi(4C) | x_lo(4C) | 0] (KpA columns, stationary), weight row = [w_hi | w_hi | 0] over
// the same KpA columns followed by [w_lo | 0] over Kp1 = pad64(4C) columns that multiply the A block's FIRST slabs again (their columns
// past 4C hold x_lo where the weights are zero).  up8 of fcn_8: 9 + 5 weight blocks instead of 5, the epilogue unchanged.
static int x3_kpa(int Cin) { return (8 * Cin + 63) / 64 * 64; }
size_t tc_deconv_x3_scratch_bytes(int B, int IH, int IW, int Cin) { return (size_t)B * (IH + 1) * (IW + 1) * x3_kpa(Cin) * 2; }
bool tc_deconv_x3_supported(int k, int s, int Cin, int Cout) {
  if (!tc_deconv_supported(k, s, Cin, Cout)) return false;
  const size_t fixed = (size_t)(x3_kpa(Cin) / 64) * 16384 + 1024, b_bytes = (size_t)2 * tc_deconv_cpp(Cout) * 128;
  return fixed + 2 * b_bytes <= kSmemMax;
}
// two-pass variant (probabilities / soft centroid, where the staging leaves no room for the [hi | lo] A block): pass 1 multiplies
// A = x_hi with [w_hi | w_lo] (the second half of the weight blocks wraps onto the same A slabs) into fp32 logits, pass 2 multiplies
// A = x_lo with w_hi, adds pass 1's logits in the epilogue (DeconvParams::addend) and decodes.
void tc_deconv_x3_pack_weights_hilo(const float* w_phase, int s, int Cin, int Cout, uint16_t (*f2bf)(float), float (*bf2f)(uint16_t),
                                    std::vector<uint16_t>& out) {
  const int Kp = tc_deconv_kp(Cin), KpW = 2 * Kp, cpp = tc_deconv_cpp(Cout), nph = s * s;
  out.assign((size_t)nph * cpp * KpW, 0);
  for (int ph = 0; ph < nph; ++ph)
    for (int t = 0; t < 4; ++t)
      for (int c = 0; c < Cin; ++c)
        for (int o = 0; o < Cout; ++o) {
          const float wv = w_phase[(((size_t)ph * 4 + t) * Cin + c) * Cout + o];
          const uint16_t hi = f2bf(wv);
          uint16_t* row = &out[((size_t)ph * cpp + o) * KpW];
          row[(size_t)t * Cin + c] = hi;
          row[Kp + (size_t)t * Cin + c] = f2bf(wv - bf2f(hi));
        }
}
void tc_deconv_x3_pack_weights(const float* w_phase, int s, int Cin, int Cout, uint16_t (*f2bf)(float), float (*bf2f)(uint16_t),
                               std::vector<uint16_t>& out) {
  const int KpA = x3_kpa(Cin), Kp1 = tc_deconv_kp(Cin), KpW = KpA + Kp1, cpp = tc_deconv_cpp(Cout), nph = s * s;
  out.assign((size_t)nph * cpp * KpW, 0);
  for (int ph = 0; ph < nph; ++ph)
    for (int t = 0; t < 4; ++t)
      for (int c = 0; c < Cin; ++c)
        for (int o = 0; o < Cout; ++o) {
          const float wv = w_phase[(((size_t)ph * 4 + t) * Cin + c) * Cout + o];
          const uint16_t hi = f2bf(wv), lo = f2bf(wv - bf2f(hi));
          uint16_t* row = &out[((size_t)ph * cpp + o) * KpW];
          const size_t k = (size_t)t * Cin + c;
          row[k] = hi; row[4 * Cin + k] = hi; row[KpA + k] = lo;
        }
}

// w_phase: [s*s][2][2][Cin][Cout] (net.cu set_weights)  ->  out bf16 [s*s*cpp][Kp], k = (u*2+v)*Cin + c
void tc_deconv_pack_weights(const float* w_phase, int s, int Cin, int Cout, uint16_t (*f2bf)(float), std::vector<uint16_t>& out) {
  const int Kp = tc_deconv_kp(Cin), cpp = tc_deconv_cpp(Cout), nph = s * s;
  out.assign((size_t)nph * cpp * Kp, 0);
  for (int ph = 0; ph < nph; ++ph)
    for (int t = 0; t < 4; ++t)
      for (int c = 0; c < Cin; ++c)
        for (int o = 0; o < Cout; ++o)
          out[((size_t)ph * cpp + o) * Kp + (size_t)t * Cin + c] = f2bf(w_phase[(((size_t)ph * 4 + t) * Cin + c) * Cout + o]);
}

struct TcDeconvPlan {
  CUtensorMap tmA, tmB;
  DeconvParams p;
  int grid;
  size_t smem;
  const void* scratch;
  int B, h, w, C, Kp, x3;
};

int tc_deconv_plan_create(const fld_handle* hd, void* scratch, const __nv_bfloat16* w_packed, int B, int IH, int IW, int Cin, int Cout,
                          int s, TcDeconvPlan** out, int x3) {
  if (!hd->encode_tiled) { fld_set_error("cuTensorMapEncodeTiled entry point not available"); return FLD_ERR_CUDA; }
  if (!tc_deconv_supported(2 * s, s, Cin, Cout)) { fld_set_error("tc_deconv: unsupported geometry"); return FLD_ERR_INVALID; }
  EncodeTiledFn enc = (EncodeTiledFn)hd->encode_tiled;
  TcDeconvPlan* pl = new TcDeconvPlan();
  DeconvParams& p = pl->p;
  // x3: 0 plain bf16; 1 FLD_BF16X3 class maps (A = [hi | lo]); 2 / 3 the two passes of the FLD_BF16X3 probability / centroid variant
  const int Kp = x3 == 1 ? x3_kpa(Cin) : tc_deconv_kp(Cin), cpp = tc_deconv_cpp(Cout);   // columns of A
  const int KpW = x3 == 1 ? Kp + tc_deconv_kp(Cin) : x3 == 2 ? 2 * Kp : Kp;             // columns of the weight matrix
  p.out = nullptr;
  p.GH = IH + 1; p.GW = IW + 1;
  p.M = (long long)B * p.GH * p.GW;
  p.s = s; p.Cout = Cout; p.cpp = cpp; p.BN = 2 * cpp; p.kblocks = KpW / 64; p.a_blocks = Kp / 64;
  p.n_ntiles = s * s / 2;
  p.mtiles = (int)((p.M + 127) / 128);
  p.total_tiles = p.mtiles * p.n_ntiles;
  const size_t fixed = x3 == 1 ? (size_t)(Kp / 64) * 16384 + 1024 : fixed_smem(Cin, Cout);
  const size_t b_bytes = (size_t)p.BN * 128;
  p.stages = (int)std::min<size_t>(kMaxStages, (kSmemMax - fixed) / b_bytes);
  p.mode = 0; p.trace = nullptr; p.acc = nullptr; p.walk = 0; p.addend = nullptr;
  pl->smem = fixed + (size_t)p.stages * b_bytes;
  // work units: the smallest split of the N range that still gives every SM a few units
  p.nsplit = 1;
  while (p.nsplit < p.n_ntiles && (long long)p.mtiles * p.nsplit < 4LL * hd->sm_count && p.n_ntiles % (p.nsplit * 2) == 0) p.nsplit *= 2;
  p.cn = p.n_ntiles / p.nsplit;
  p.units = p.mtiles * p.nsplit;
  pl->grid = std::min(p.units, hd->sm_count);
  pl->scratch = scratch; pl->B = B; pl->h = IH; pl->w = IW; pl->C = Cin; pl->Kp = Kp; pl->x3 = x3;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)p.M};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scratch, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(deconv A) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)KpW, (cuuint64_t)s * s * cpp};
    cuuint64_t strides[1] = {(cuuint64_t)KpW * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; fld_set_error("cuTensorMapEncodeTiled(deconv B) failed: %d", (int)r); return FLD_ERR_CUDA; }
  }
  *out = pl;
  return FLD_OK;
}

void tc_deconv_plan_destroy(TcDeconvPlan* p) { delete p; }

// in: fp32 NHWC [B][h][w][C]; out: mode 0/1 fp32 [B][(h+1)s][(w+1)s][Cout], mode 2 int64 [B][(h+1)s][(w+1)s]
// scratch of the fused soft-centroid decode (mode 3).  A fused top-n variant (per-tile sorted candidate lists + merge kernel) was
// built and measured 2.9x SLOWER than materialising the probabilities and running the stand-alone decode: a tile holds only 128
// pixels per class, so every warp stays on the divergent insertion path for the whole scan; it was removed.
size_t tc_deconv_acc_bytes(int B, int Cout) { return (size_t)B * Cout * 3 * sizeof(float); }

int tc_deconv_run(const TcDeconvPlan* pl, const float* in, void* out, int mode, cudaStream_t st, float* acc, double thresh, const float* addend) {
  if (pl->p.total_tiles == 0) return FLD_OK;
  if (pl->x3 == 1 || pl->x3 == 3) {
    if (pl->x3 == 1 && mode != 2) { fld_set_error("tc_deconv: the [hi | lo] FLD_BF16X3 variant produces class maps only"); return FLD_ERR_INVALID; }
    const long long n = pl->p.M * (pl->Kp / 2);
    deconv_im2col_x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, (__nv_bfloat162*)pl->scratch, pl->p.M, pl->h, pl->w, pl->C, pl->Kp,
                                                                      pl->x3 == 3 ? 1 : 0);
    FLD_LAUNCHED();
  } else if (pl->C % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    deconv_im2col_vec4_kernel<<<(unsigned)((pl->p.M + 7) / 8), 256, 0, st>>>((const float4*)in, (uint2*)pl->scratch, pl->p.M, pl->h, pl->w, pl->C / 4,
                                                                          pl->Kp / 4);
    FLD_LAUNCHED();
  } else {
    const long long n = pl->p.M * (pl->Kp / 2);
    deconv_im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, (__nv_bfloat162*)pl->scratch, pl->p.M, pl->h, pl->w, pl->C, pl->Kp);
    FLD_LAUNCHED();
  }
  DeconvParams p = pl->p;
  p.out = out; p.mode = mode; p.trace = nullptr; p.acc = acc; p.addend = addend;
  { const char* e = getenv("FLD_TC_DECONV_WALK"); p.walk = (e && atoi(e) != 0) ? 1 : 0; }
  if (mode == 3) {
    if (!acc) { fld_set_error("tc_deconv: mode 3 needs an accumulator buffer"); return FLD_ERR_INVALID; }
    if (pl->p.GH * pl->p.s > 65535 || pl->p.GW * pl->p.s > 65535) { fld_set_error("tc_deconv: map too large for the fused centroid"); return FLD_ERR_INVALID; }
    FLD_CUDA(cudaMemsetAsync(acc, 0, (size_t)pl->B * pl->p.Cout * 3 * sizeof(float), st));
  }
  if (getenv("FLD_TC_TRACE")) {
    static unsigned long long* tbuf = nullptr;
    if (!tbuf) FLD_CUDA(cudaMalloc(&tbuf, 3 * kTraceN * 8));
    FLD_CUDA(cudaMemsetAsync(tbuf, 0, 3 * kTraceN * 8, st));
    p.trace = tbuf;
  }
  size_t smem = pl->smem;
  if (mode == 2) {   // class-map mode stages nothing: the staging bytes deepen the weight ring instead
    const size_t b_bytes = (size_t)p.BN * 128, fixed = (size_t)p.a_blocks * 16384 + 1024;
    p.stages = (int)std::min<size_t>(kMaxStages, (kSmemMax - fixed) / b_bytes);
    smem = fixed + (size_t)p.stages * b_bytes;
  }
  // soft centroid over the reference's 68 classes: the 16-warp epilogue (FLD_TC_DECONV_EW8=1 keeps the 8-warp one)
  const bool ew16 = mode == 3 && p.Cout == 68 && !p.walk && !addend && !getenv("FLD_TC_DECONV_EW8");   // (the addend lives in the 8-warp epilogue)
  if (ew16) {
    const size_t b_bytes = (size_t)p.BN * 128;
    const size_t fixed = (size_t)p.a_blocks * 16384 + 1024 + (size_t)2 * red_a_bytes(68) + 2 * kRedB16Bytes;
    p.stages = (int)std::min<size_t>(kMaxStages, (kSmemMax - fixed) / b_bytes);
    smem = fixed + (size_t)p.stages * b_bytes;
    FLD_CUDA(cudaFuncSetAttribute(deconv_gemm_kernel<68, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    deconv_gemm_kernel<68, 16><<<pl->grid, 64 + 32 * 16, smem, st>>>(pl->tmA, pl->tmB, p);
  } else if (p.Cout == 68) {   // the reference's n_classes (scripts/cli.py:39, training.py:110)
    FLD_CUDA(cudaFuncSetAttribute(deconv_gemm_kernel<68, kEpiWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    deconv_gemm_kernel<68, kEpiWarps><<<pl->grid, kThreads, smem, st>>>(pl->tmA, pl->tmB, p);
  } else {
    FLD_CUDA(cudaFuncSetAttribute(deconv_gemm_kernel<0, kEpiWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    deconv_gemm_kernel<0, kEpiWarps><<<pl->grid, kThreads, smem, st>>>(pl->tmA, pl->tmB, p);
  }
  FLD_LAUNCHED();
  if (mode == 3) {
    const long long n = (long long)pl->B * p.Cout;
    const double hw = (double)(p.GH * p.s) * (double)(p.GW * p.s);
    const double cx = p.walk ? 0.0 : (double)(p.GW * p.s / 2), cy = p.walk ? 0.0 : (double)(p.GH * p.s / 2);
    centroid_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, n, hw, thresh, cx, cy, (double*)out);
    FLD_LAUNCHED();
  }
  if (p.trace) {  // dump CTA 0's event log: "<role> <tag> <clock>" per line
    static int n_dump = 0;
    std::vector<unsigned long long> hbuf(3 * kTraceN);
    FLD_CUDA(cudaStreamSynchronize(st));
    FLD_CUDA(cudaMemcpy(hbuf.data(), p.trace, hbuf.size() * 8, cudaMemcpyDeviceToHost));
    char name[256];
    snprintf(name, sizeof(name), "%s/trace_deconv_%02d_s%d_mode%d.txt", getenv("FLD_TC_TRACE"), n_dump++, p.s, p.mode);
    if (FILE* f = fopen(name, "w")) {
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < kTraceN && hbuf[r * kTraceN + i]; ++i)
          fprintf(f, "%d %llu %llu\n", r, hbuf[r * kTraceN + i] & 15, hbuf[r * kTraceN + i] >> 4);
      fclose(f);
    }
  }
  return FLD_OK;
}
