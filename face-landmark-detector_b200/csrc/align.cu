// align.cu — batched Umeyama similarity fit fused with a cv2.warpAffine-exact bilinear warp.
//
// Build-defined stage: nothing in the reference computes this; semantics
// are "fp64 closed-form Umeyama, then cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of OpenCV
// 4.13", reproduced bit-exactly with OpenCV's fixed-point scheme (1/32 px sub-pixel, 15-bit weights).
//
// The fit is warp-parallel: lane l accumulates points l, l+32, ... in index order, then a fixed xor-butterfly
// (16, 8, 4, 2, 1) combines the lanes in fp64 without FMA contraction — a deterministic summation tree that
// oracle/align.py mirrors, so M is bit-identical to the CPU oracle.
//
// Two warp kernels:
//   * align_tile_kernel (round 2; C == 3, 16-byte-aligned frame rows, output a multiple of 16 x 16, <= 256 x 256):
//     round 1's per-pixel global gathers were L1-tag / issue bound (ncu: ~21 sectors per warp load, 107 instructions per
//     output pixel, DRAM at 16 %).  Here the four warps of a CTA walk the 16 x 16-pixel output tiles of a face over a pooled box ring
//     (as many warps take tiles as boxes fit; warp pairs share a tile when at most two fit — see the kernel); the source
//     bounding box of a tile (a square for a similarity transform) arrives by ONE TMA load into a shared-memory ring
//     (out-of-frame bytes are zero-filled by the tensor map = BORDER_CONSTANT 0, so the blend has no bounds checks and
//     32-bit addresses), with the next tiles' boxes in flight while the current one is blended.  The tensor-map box is
//     fixed per map, so eight maps (box sides 16..80 source pixels) are passed and the face picks the smallest that fits.
//     The blend is exact integer arithmetic: out = (sum_ij a_i b_j p_ij + 512) >> 10 with a = (32 - fx, fx),
//     b = (32 - fy, fy) (OpenCV's 15-bit table is exactly 32 a_i b_j for 5-bit fractions; its +-1 fix-ups at fx = fy = 0
//     cannot change the rounded byte), computed as two IDP.2A per channel (16-bit weights a_i b_j against the gathered (p00, p01, p10, p11) bytes).
//     Tiles whose box does not fit the class (never for similarity transforms; possible for caller matrices with
//     shear) and faces scaled down by more than ~3.7x take the per-pixel global path below.
//     Batches of >= 2048 faces with a caller scratch buffer (fld_align_ordered) run the fit in align_fit_kernel (one warp per
//     face) and a counting sort in align_order_kernel, and the tile kernel then takes the faces big boxes first (no scheduling tail).
//   * align_warp_kernel (round 1): any C in {1, 3, 4}, any shape; one CTA per (face, row block), per-pixel global loads.
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int kAlignThreads = 256;

struct Fit {
  double M[6];
  double iM[6];
  int ok;
};

// fp64 helpers that forbid FMA contraction (numpy / OpenCV evaluate mul and add separately)
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }

// fixed summation tree over the warp: every lane ends up with the same value
__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = da(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// Whole warp (all 32 lanes converged).  Source point i: five_point ? reduced iBUG-68 point : marks[i].
// Sums: lane l takes points l, l+32, ... sequentially, then warp_tree_sum.  Lane 0 writes M[6].
__device__ void fit_similarity_warp(const float* __restrict__ marks, int N, const double* __restrict__ tmpl, int five_point,
                                    double* M, int lane) {
  const int n = five_point ? 5 : N;
  auto src_pt = [&](int i, double& x, double& y) {
    if (!five_point) { x = (double)marks[2 * i]; y = (double)marks[2 * i + 1]; return; }
    if (i < 2) {  // eye centres: six landmarks summed in index order, / 6
      const int b0 = i == 0 ? 36 : 42;
      double ax = 0, ay = 0;
      for (int k = b0; k < b0 + 6; ++k) { ax = da(ax, (double)marks[2 * k]); ay = da(ay, (double)marks[2 * k + 1]); }
      x = ax / 6.0; y = ay / 6.0;
    } else {
      const int k = i == 2 ? 30 : (i == 3 ? 48 : 54);
      x = (double)marks[2 * k]; y = (double)marks[2 * k + 1];
    }
  };
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = lane; i < n; i += 32) {
    double x, y;
    src_pt(i, x, y);
    s0 = da(s0, x); s1 = da(s1, y); s2 = da(s2, tmpl[2 * i]); s3 = da(s3, tmpl[2 * i + 1]);
  }
  const double dn = (double)n;
  const double mpx = warp_tree_sum(s0) / dn, mpy = warp_tree_sum(s1) / dn;
  const double mqx = warp_tree_sum(s2) / dn, mqy = warp_tree_sum(s3) / dn;
  double var = 0, a = 0, b = 0, c = 0, d = 0;
  for (int i = lane; i < n; i += 32) {
    double sx, sy;
    src_pt(i, sx, sy);
    const double x = da(sx, -mpx), y = da(sy, -mpy);
    const double qx = da(tmpl[2 * i], -mqx), qy = da(tmpl[2 * i + 1], -mqy);
    var = da(var, da(dm(x, x), dm(y, y)));
    a = da(a, dm(qx, x)); b = da(b, dm(qx, y)); c = da(c, dm(qy, x)); d = da(d, dm(qy, y));
  }
  var = warp_tree_sum(var); a = warp_tree_sum(a); b = warp_tree_sum(b); c = warp_tree_sum(c); d = warp_tree_sum(d);
  if (lane != 0) return;
  const double P = da(a, d), Q = da(c, -b);
  if (var == 0.0 || (P == 0.0 && Q == 0.0) || !isfinite(var) || !isfinite(P) || !isfinite(Q)) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < 6; ++i) M[i] = nan;
    return;
  }
  const double l00 = P / var, l01 = -Q / var, l10 = Q / var, l11 = P / var;
  M[0] = l00; M[1] = l01; M[2] = da(mqx, -da(dm(l00, mpx), dm(l01, mpy)));
  M[3] = l10; M[4] = l11; M[5] = da(mqy, -da(dm(l10, mpx), dm(l11, mpy)));
}

__device__ void invert_affine(const double* M, double* iM) {
  double D = da(dm(M[0], M[4]), -dm(M[1], M[3]));
  D = (D != 0.0) ? 1.0 / D : 0.0;
  const double i00 = dm(M[4], D), i01 = dm(-M[1], D), i10 = dm(-M[3], D), i11 = dm(M[0], D);
  iM[0] = i00; iM[1] = i01; iM[2] = da(dm(-i00, M[2]), -dm(i01, M[5]));
  iM[3] = i10; iM[4] = i11; iM[5] = da(dm(-i10, M[2]), -dm(i11, M[5]));
}

// One warp: fit (or take the caller's matrix), validate, invert.  `fit` is in shared memory; lane 0 finishes it.
__device__ __forceinline__ void warp_fit(Fit& fit, int face, int F, const int32_t* __restrict__ face2frame,
                                         const float* __restrict__ marks, int N, const double* __restrict__ tmpl, int five_point,
                                         const double* __restrict__ M_in, double* __restrict__ M_out, bool write_M, int lane) {
  if (M_in) {
    if (lane < 6) fit.M[lane] = M_in[(size_t)face * 6 + lane];
  } else {
    fit_similarity_warp(marks + (size_t)face * N * 2, N, tmpl, five_point, fit.M, lane);
  }
  __syncwarp();
  if (lane == 0) {
    int ok = 1;
    for (int i = 0; i < 6; ++i) ok &= isfinite(fit.M[i]) ? 1 : 0;
    const int fr = face2frame[face];
    if (fr < 0 || fr >= F) ok = 0;
    if (ok) invert_affine(fit.M, fit.iM);
    fit.ok = ok;
    if (M_out && write_M) for (int i = 0; i < 6; ++i) M_out[(size_t)face * 6 + i] = fit.M[i];
  }
}

// Warp 0 of the CTA fits; every thread reads `fit` after a __syncthreads.
__device__ __forceinline__ void cta_fit(Fit& fit, int face, int F, const int32_t* __restrict__ face2frame,
                                        const float* __restrict__ marks, int N, const double* __restrict__ tmpl, int five_point,
                                        const double* __restrict__ M_in, double* __restrict__ M_out, bool write_M, int tid) {
  if (tid >= 32) return;
  warp_fit(fit, face, F, face2frame, marks, N, tmpl, five_point, M_in, M_out, write_M, tid);
}

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

// cv2's fixed-point coordinate tables (App. B.2 step 2): saturate_cast<int>(rint(v)).  In range (|v| < 2^31, i.e. source coordinates
// below 2 M pixels) the native F2I.S32.F64 conversion is exact; beyond it cv2 itself returns INT_MIN garbage — reproduced here so
// that the 64-bit conversion (a ~20-instruction emulation per table entry) is not needed.
__device__ __forceinline__ int cv_round(double v) { return (v >= -2147483648.0 && v < 2147483648.0) ? __double2int_rn(v) : INT_MIN; }
__device__ __forceinline__ int tab_ad(const Fit& fit, int x) { return cv_round(dm(dm(fit.iM[0], (double)x), 1024.0)); }
__device__ __forceinline__ int tab_bd(const Fit& fit, int x) { return cv_round(dm(dm(fit.iM[3], (double)x), 1024.0)); }
__device__ __forceinline__ int tab_X0(const Fit& fit, int y) { return cv_round(dm(da(dm(fit.iM[1], (double)y), fit.iM[2]), 1024.0)) + 16; }
__device__ __forceinline__ int tab_Y0(const Fit& fit, int y) { return cv_round(dm(da(dm(fit.iM[4], (double)y), fit.iM[5]), 1024.0)) + 16; }

// six consecutive bytes starting at an arbitrary address, from two aligned 8-byte loads
__device__ __forceinline__ uint64_t load6(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  const unsigned sh = (unsigned)(a & 7) * 8;
  const uint64_t lo = __ldg(q);
  if (sh <= 16) return lo >> sh;
  const uint64_t hi = __ldg(q + 1);
  return (lo >> sh) | (hi << (64 - sh));
}

// One output pixel of a 3-channel frame by per-pixel global loads with full bounds handling (BORDER_CONSTANT 0).
// X, Y: fixed-point source coordinates (1/32 px).  Returns the three bytes in bits [0,24).
__device__ __forceinline__ uint32_t blend_px_global3(const uint8_t* __restrict__ frame, int H, int W, size_t row, int X, int Y) {
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
  const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  int c[3] = {0, 0, 0};
  if (sx >= -1 && sy >= -1 && sx < W && sy < H) {
    const bool x0in = sx >= 0, x1in = sx + 1 < W, y0in = sy >= 0, y1in = sy + 1 < H;
    const uint8_t* p0 = frame + (ptrdiff_t)sy * (ptrdiff_t)row + (ptrdiff_t)sx * 3;
    const uint8_t* p1 = p0 + row;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int t0 = (y0in && x0in) ? p0[ch] : 0, t1 = (y0in && x1in) ? p0[3 + ch] : 0;
      const int t2 = (y1in && x0in) ? p1[ch] : 0, t3 = (y1in && x1in) ? p1[3 + ch] : 0;
      c[ch] = w00 * t0 + w01 * t1 + w10 * t2 + w11 * t3;
    }
  }
  return (uint32_t)((c[0] + 16384) >> 15) | ((uint32_t)((c[1] + 16384) >> 15) << 8) | ((uint32_t)((c[2] + 16384) >> 15) << 16);
}

// ------------------------------------------------------------------------------------------------ PTX (TMA / mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
#pragma unroll 1
  for (uint32_t it = 0; it < 20000000u; ++it)
    if (mbar_try_wait(bar, parity)) return;
  printf("fld align: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while its predecessor in
// the stream still runs; it must wait here before touching the predecessor's results (a no-op for an ordinary launch)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// ------------------------------------------------------------------------------------------------ tile kernel
constexpr int kTile = 16;          // output tile edge
constexpr int kTileWarps = 4;      // warps of a CTA; how many of them take tiles depends on the face's box size (see the ring pool below)
constexpr int kTileThreads = 32 * kTileWarps;
constexpr int kNumCls = 8;
constexpr int kMaxOut = 128;       // tables are static shared arrays
constexpr int kMaxBuf = 4;
constexpr int kRingBytes = 27 * 1024;
// box side (source pixels) per class; rows = side, bytes per row = round16(3 * side + 15) (start is floored to 16 bytes)
__host__ __device__ constexpr int cls_side(int k) { return k == 0 ? 16 : k == 1 ? 20 : k == 2 ? 24 : k == 3 ? 32 : k == 4 ? 40 : k == 5 ? 48 : k == 6 ? 64 : 80; }
__host__ __device__ constexpr int cls_bw(int k) { return (3 * cls_side(k) + 15 + 15) & ~15; }

// Box class of a face from the linear part of its inverse map: a 16 x 16 tile spans 15 * (|i00| + |i01|) source columns and
// 15 * (|i10| + |i11|) rows, plus the second tap and the floor / round slack.  -1: no class fits (per-pixel global path).
__device__ __forceinline__ int box_class(const Fit& fit, int ring_bytes) {
  const double ex = 15.0 * (fabs(fit.iM[0]) + fabs(fit.iM[1])), ey = 15.0 * (fabs(fit.iM[3]) + fabs(fit.iM[4]));
  const double e = fmax(ex, ey) + 3.0;
  int cls = -1;
  if (e < 4096.0) {
    const int side = (int)ceil(e);
    for (int k = 0; k < kNumCls; ++k) if (side <= cls_side(k)) { cls = k; break; }
  }
  if (cls >= 0 && (cls_bw(cls) * cls_side(cls) + 16 + 127) / 128 * 128 > ring_bytes) cls = -1;   // box larger than the ring
  return cls;
}

struct AlignMaps { CUtensorMap m[kNumCls]; };

// Ordered mode (large batches): the fit runs in its own kernel, one warp per face, and a counting sort orders the faces by
// decreasing box size, so that the expensive faces (strongly reduced ones: up to 20 KB per tile box) start first and the cheap
// ones fill the tail.  Measured on config C4 (4096 faces, scales 0.28 - 1.4): random order 0.194 ms, big-first 0.168 ms.
struct PreFit { double iM[6]; int ok; int key; };
constexpr int kOrderKeys = 16;        // key 0: degenerate fit (zero crop), 1 + class, kNumCls + 1: global path

struct TileArgs {
  const uint8_t* frames;
  const int32_t* face2frame;
  const float* marks;
  const double* tmpl;
  const double* M_in;
  double* M_out;
  uint8_t* crops;
  int F, H, W, N, five_point, out_h, out_w, ysplit;
  int ring_bytes;   // dynamic shared memory of the box ring (multiple of 256)
  int min_bufs;     // buffers a warp should have before another warp is activated
  int pair_max;     // warps pair up on a tile when at most this many boxes fit the ring
  const struct PreFit* fits;   // ordered mode: per-face inverse maps from align_fit_kernel (null: the CTA fits its face itself)
  const int32_t* perm;         // ordered mode: faces by decreasing box size
};

// Blend of four consecutive output pixels from the staged source box.  bxv / byv: the row's X0 / Y0; av / bv: adelta / bdelta of
// the four columns; base: shared-memory address of the box; corr = -oy * BW - ox.  Returns the 12 output bytes in three words.
__device__ __forceinline__ void blend4_smem(int bxv, int byv, const int4& a4, const int4& b4, uint32_t base, int corr, int BW, uint32_t (&w)[3]) {
  const int av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
  uint32_t q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int X = (bxv + av[j]) >> 5, Y = (byv + bv[j]) >> 5;
    const int sx = X >> 5, sy = Y >> 5;
    const uint32_t fx = X & 31, fy = Y & 31;
    const uint32_t o = base + (uint32_t)(sy * BW + sx * 3 + corr);
    const uint32_t oa = o & ~3u, k8 = (o & 3u) << 3;
    // bytes o .. o+5 of a row: the third word is needed only when o = 3 (mod 4) — a predicated load for a quarter of the lanes
    // (the shared-memory gathers are bank-conflict bound: 2.4 wavefronts per load with all lanes active)
    const uint32_t l0 = lds32(oa), l1 = lds32(oa + 4), m0 = lds32(oa + BW), m1 = lds32(oa + BW + 4);
    uint32_t l2 = 0u, m2 = 0u;
    if (k8 == 24u) { l2 = lds32(oa + 8); m2 = lds32(oa + BW + 8); }
    const uint32_t w0 = __funnelshift_r(l0, l1, k8), w1 = __funnelshift_r(l1, l2, k8);   // row sy  : c0 c1 c2 c0' | c1' c2'
    const uint32_t v0 = __funnelshift_r(m0, m1, k8), v1 = __funnelshift_r(m1, m2, k8);   // row sy+1
    const uint32_t t0 = prmt(w0, w1, 0x5241u), t1 = prmt(v0, v1, 0x5241u);                 // c1 c1' c2 c2'
    const uint32_t P0 = prmt(w0, v0, 0x7430u), P1 = prmt(t0, t1, 0x5410u), P2 = prmt(t0, t1, 0x7632u);   // p00 p01 p10 p11 per channel
    // weights w_ij = a_i b_j (<= 1024) as 16-bit pairs: out = (sum_ij w_ij p_ij + 512) >> 10, two IDP.2A per channel
    const uint32_t A16 = fx * 65535u + 32u;          // (32 - fx) | fx << 16
    const uint32_t W01 = (32u - fy) * A16, W23 = fy * A16;
    // (S + 512) >> 10 is byte 2 of (S + 512) * 64: the multiply runs on the FMA pipe, which is idle next to the ALU pipe here
    const uint32_t S0 = __dp2a_hi(W23, P0, __dp2a_lo(W01, P0, 512u)) * 64u;
    const uint32_t S1 = __dp2a_hi(W23, P1, __dp2a_lo(W01, P1, 512u)) * 64u;
    const uint32_t S2 = __dp2a_hi(W23, P2, __dp2a_lo(W01, P2, 512u)) * 64u;
    q[j] = prmt(prmt(S0, S1, 0x0062u), S2, 0x0610u);
  }
  w[0] = prmt(q[0], q[1], 0x4210u);
  w[1] = prmt(q[1], q[2], 0x5421u);
  w[2] = prmt(q[2], q[3], 0x6542u);
}

// Warp-autonomous pipelines: the CTA (4 warps) shares the face's fit, coordinate tables and tile origins; after that each active
// warp walks its own tiles (every n_act-th tile of the CTA's tile rows) with its own slice of the shared-memory ring and its own mbarriers:
// lane 0 issues the TMA load of the tile after next into the buffer the warp has just finished reading (__syncwarp), so there is
// no CTA-wide barrier in the loop.  A lane blends 2 x 4 pixels of a tile (rows r and r + 8, four consecutive columns).
__global__ void __launch_bounds__(kTileThreads)
align_tile_kernel(const __grid_constant__ AlignMaps maps, const TileArgs p) {
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ Fit fit;
  __shared__ __align__(16) int s_ad[kMaxOut], s_bd[kMaxOut], s_X0[kMaxOut], s_Y0[kMaxOut];
  __shared__ int s_ox[(kMaxOut / kTile) * (kMaxOut / kTile)], s_oy[(kMaxOut / kTile) * (kMaxOut / kTile)];   // per tile of this CTA: box origin (byte column, row); ox = INT_MIN -> global path
  __shared__ __align__(8) uint64_t full_bar[kTileWarps][kMaxBuf], empty_bar[kTileWarps][kMaxBuf];
  __shared__ int s_cls;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.perm) pdl_wait();                       // ordered mode: the order kernel's permutation (and, through it, the fit kernel's maps)
  const int face = p.perm ? p.perm[blockIdx.x] : blockIdx.x;
  const int tiles_x = p.out_w / kTile, tiles_y = p.out_h / kTile;
  const int ty_per = (tiles_y + p.ysplit - 1) / p.ysplit;
  const int ty0 = blockIdx.y * ty_per, ty1 = min(tiles_y, ty0 + ty_per);
  const int T = max(0, ty1 - ty0) * tiles_x;
  const uint32_t ring = (smem_u32(ring_raw) + 127u) & ~127u;

  if (p.fits) {
    if (tid < 6) fit.iM[tid] = p.fits[face].iM[tid];
    if (tid == 6) fit.ok = p.fits[face].ok;
  } else {
    cta_fit(fit, face, p.F, p.face2frame, p.marks, p.N, p.tmpl, p.five_point, p.M_in, p.M_out, blockIdx.y == 0, tid);
  }
  __syncthreads();
  uint8_t* crop = p.crops + (size_t)face * p.out_h * p.out_w * 3;
  if (T == 0) return;
  if (!fit.ok) {
    uint32_t* c4 = reinterpret_cast<uint32_t*>(crop + (size_t)ty0 * kTile * p.out_w * 3);   // 16 * out_w * 3 bytes per tile row: 4-byte multiple
    const int n4 = (ty1 - ty0) * kTile * p.out_w * 3 / 4;
    for (int i = tid; i < n4; i += kTileThreads) c4[i] = 0u;
    return;
  }
  for (int x = tid; x < p.out_w; x += kTileThreads) { s_ad[x] = tab_ad(fit, x); s_bd[x] = tab_bd(fit, x); }
  for (int y = ty0 * kTile + tid; y < ty1 * kTile; y += kTileThreads) { s_X0[y] = tab_X0(fit, y); s_Y0[y] = tab_Y0(fit, y); }
  if (tid == 0) {
    const int cls = box_class(fit, p.ring_bytes);
    s_cls = cls;
    for (int i = 0; i < kTileWarps * kMaxBuf; ++i) {
      mbar_init(smem_u32(&full_bar[0][0]) + 8 * i, 1);
      mbar_init(smem_u32(&empty_bar[0][0]) + 8 * i, 2);     // used by warp pairs only
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int cls = s_cls;
  const int BW = cls >= 0 ? cls_bw(cls) : 0, BH = cls >= 0 ? cls_side(cls) : 0;
  // tile origins (exact extremes: the tables are monotone, so min / max sit at the tile's first / last column and row)
  for (int t = tid; t < T; t += kTileThreads) {
    const int ty = ty0 + t / tiles_x, tx = t - (t / tiles_x) * tiles_x;
    const int xa = s_ad[tx * kTile], xb = s_ad[tx * kTile + kTile - 1], ya = s_X0[ty * kTile], yb = s_X0[ty * kTile + kTile - 1];
    const int ua = s_bd[tx * kTile], ub = s_bd[tx * kTile + kTile - 1], va = s_Y0[ty * kTile], vb = s_Y0[ty * kTile + kTile - 1];
    const long long xmin = (long long)min(xa, xb) + min(ya, yb), xmax = (long long)max(xa, xb) + max(ya, yb);
    const long long ymin = (long long)min(ua, ub) + min(va, vb), ymax = (long long)max(ua, ub) + max(va, vb);
    int ox = INT_MIN, oy = 0;
    if (cls >= 0 && xmin > -(1ll << 30) && xmax < (1ll << 30) && ymin > -(1ll << 30) && ymax < (1ll << 30)) {
      const int sx0 = (int)(xmin >> 10), sx1 = (int)(xmax >> 10), sy0 = (int)(ymin >> 10), sy1 = (int)(ymax >> 10);
      const int bx = (3 * sx0) & ~15;                  // floor to 16 bytes (two's complement: also for negatives)
      if (3 * (sx1 + 2) - bx <= BW && sy1 + 2 - sy0 <= BH && sx0 > -30000 && sx1 < 30000 && sy0 > -30000 && sy1 < 30000) { ox = bx; oy = sy0; }
    }
    s_ox[t] = ox; s_oy[t] = oy;
  }
  __syncthreads();

  // ---- pipelines.  The ring is a pool shared by the CTA's warps.  A face with small boxes (magnified or 1:1 faces, where the
  // blend is instruction / shared-memory bound) runs four autonomous warps, each with its own tiles, buffers and mbarriers and no
  // barrier between warps.  When at most two boxes fit the pool (strongly reduced faces: HBM bound, and with one warp per box a
  // face took 49 x ~3 us — the long poles of a mixed batch) the warps work in PAIRS on one tile, eight rows each, and hand the
  // buffer back through an mbarrier with two arrivals.
  const int frame_idx = p.face2frame[face];
  const uint8_t* frame = p.frames + (size_t)frame_idx * p.H * p.W * 3;
  const size_t row = (size_t)p.W * 3;
  const uint32_t box_bytes = (uint32_t)(BW * BH);
  const uint32_t buf_pitch = (box_bytes + 16u + 127u) & ~127u;       // + 16: the last pixel's third word may lie past the box
  int n_grp = kTileWarps, g = 1, nbuf = 1;
  if (cls >= 0) {
    const int fit_boxes = (int)((uint32_t)p.ring_bytes / ((uint32_t)p.min_bufs * buf_pitch));
    if (fit_boxes <= p.pair_max) { g = 2; n_grp = max(1, min(kTileWarps / 2, fit_boxes)); }
    else n_grp = min(kTileWarps, fit_boxes);
    nbuf = max(1, min(kMaxBuf, (int)((uint32_t)p.ring_bytes / ((uint32_t)n_grp * buf_pitch))));
  }
  const int grp = g == 2 ? warp >> 1 : warp, wi = g == 2 ? warp & 1 : 0;
  if (grp >= n_grp) return;
  const uint32_t my_ring = ring + (uint32_t)(grp * nbuf) * buf_pitch;
  const uint32_t my_bar = smem_u32(&full_bar[grp][0]), my_empty = smem_u32(&empty_bar[grp][0]);
  const int t_first = grp, t_step = n_grp;
  const int n_my = t_first < T ? (T - t_first + t_step - 1) / t_step : 0;
  const CUtensorMap* tm = &maps.m[cls >= 0 ? cls : 0];
  const bool issuer = lane == 0 && wi == 0;

  auto issue = [&](int i, int b) {   // issuer lane: load the box of this group's i-th tile into buffer b (= i % nbuf)
    const int t = t_first + i * t_step;
    const int ox = s_ox[t];
    if (ox == INT_MIN) return;
    mbar_arrive_expect_tx(my_bar + 8u * b, box_bytes);
    tma_load_3d(my_ring + b * buf_pitch, tm, my_bar + 8u * b, ox >> 2, s_oy[t], frame_idx);   // 32-bit elements: column = byte / 4
  };
  if (issuer) for (int i = 0; i < min(nbuf, n_my); ++i) issue(i, i);

  const int yl = lane >> 2, xg = (lane & 3) << 2;
  uint32_t phase_bits = 0u, empty_bits = 0u;   // bit b = parity the next wait on buffer b expects (global-path tiles skip their slot)
  int tx = t_first % tiles_x, ty = ty0 + t_first / tiles_x, buf = 0;
  for (int i = 0; i < n_my; ++i) {
    const int t = t_first + i * t_step;
    const int x = tx * kTile + xg;
    const int4 a4 = *reinterpret_cast<const int4*>(&s_ad[x]);
    const int4 b4 = *reinterpret_cast<const int4*>(&s_bd[x]);
    const int ox = s_ox[t], oy = s_oy[t];
    const bool staged = ox != INT_MIN;
    uint32_t base = 0u;
    int corr = 0;
    if (staged) {
      mbar_wait(my_bar + 8u * buf, (phase_bits >> buf) & 1u);
      phase_bits ^= 1u << buf;
      base = my_ring + buf * buf_pitch;
      corr = -oy * BW - ox;
    }
    auto do_row = [&](int y) {      // four pixels of output row y
      uint32_t w[3];
      if (staged) {
        blend4_smem(s_X0[y], s_Y0[y], a4, b4, base, corr, BW, w);
      } else {
        const int av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = blend_px_global3(frame, p.H, p.W, row, (s_X0[y] + av[j]) >> 5, (s_Y0[y] + bv[j]) >> 5);
        w[0] = prmt(q[0], q[1], 0x4210u); w[1] = prmt(q[1], q[2], 0x5421u); w[2] = prmt(q[2], q[3], 0x6542u);
      }
      uint32_t* d = reinterpret_cast<uint32_t*>(crop + (uint32_t)(y * p.out_w + x) * 3u);
      d[0] = w[0]; d[1] = w[1]; d[2] = w[2];
    };
    if (g == 1) {
      do_row(ty * kTile + yl);
      do_row(ty * kTile + yl + 8);
      __syncwarp();          // every lane has finished reading this tile's buffer
      if (issuer && i + nbuf < n_my) issue(i + nbuf, buf);
    } else {
      do_row(ty * kTile + wi * 8 + yl);
      __syncwarp();
      if (lane == 0 && staged) mbar_arrive(my_empty + 8u * buf);
      if (issuer && i + nbuf < n_my) {
        if (staged) { mbar_wait(my_empty + 8u * buf, (empty_bits >> buf) & 1u); empty_bits ^= 1u << buf; }
        issue(i + nbuf, buf);
      }
    }
    if (++buf == nbuf) buf = 0;
    tx += t_step;
    while (tx >= tiles_x) { tx -= tiles_x; ++ty; }     // a narrow output can have fewer tile columns than the step
  }
}

// ------------------------------------------------------------------------------------------------ ordered mode: fit + order
constexpr int kFitWarps = 4;
__global__ void __launch_bounds__(32 * kFitWarps)
align_fit_kernel(const int32_t* __restrict__ face2frame, int F, const float* __restrict__ marks, int N, const double* __restrict__ tmpl,
                 int five_point, const double* __restrict__ M_in, double* __restrict__ M_out, int B, int ring_bytes,
                 PreFit* __restrict__ fits, int32_t* __restrict__ keys) {
  __shared__ Fit sfit[kFitWarps];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int face = blockIdx.x * kFitWarps + warp;
  if (face >= B) return;
  Fit& fit = sfit[warp];
  warp_fit(fit, face, F, face2frame, marks, N, tmpl, five_point, M_in, M_out, true, lane);
  __syncwarp();
  if (lane < 6) fits[face].iM[lane] = fit.ok ? fit.iM[lane] : 0.0;
  if (lane == 6) {
    const int cls = fit.ok ? box_class(fit, ring_bytes) : 0;
    fits[face].ok = fit.ok;
    keys[face] = fits[face].key = !fit.ok ? 0 : (cls < 0 ? kNumCls + 1 : cls + 1);
  }
}

// One CTA: counting sort of the faces by key, largest key first.  Faces with equal keys land in arbitrary order (shared-memory
// atomics) — the order only schedules the work, every face's result is independent of it.
__global__ void __launch_bounds__(1024)
align_order_kernel(const int32_t* __restrict__ keys, int B, int32_t* __restrict__ perm) {
  __shared__ int hist[kOrderKeys], offs[kOrderKeys];
  pdl_launch_dependents();
  if (threadIdx.x < kOrderKeys) hist[threadIdx.x] = 0;
  __syncthreads();
  pdl_wait();                                   // the fit kernel's keys
  for (int i = threadIdx.x; i < B; i += blockDim.x) atomicAdd(&hist[keys[i]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = kOrderKeys - 1; k >= 0; --k) { offs[k] = acc; acc += hist[k]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) perm[atomicAdd(&offs[keys[i]], 1)] = i;
}

// ------------------------------------------------------------------------------------------------ generic kernel
template <int C, bool FAST>
__global__ void __launch_bounds__(kAlignThreads)
align_warp_kernel(const uint8_t* __restrict__ frames, int F, int H, int W, const int32_t* __restrict__ face2frame,
                  const float* __restrict__ marks, int N, const double* __restrict__ tmpl, int Nt, int five_point,
                  const double* __restrict__ M_in, double* __restrict__ M_out, uint8_t* __restrict__ crops,
                  int out_h, int out_w, int rtC, int rows_per) {
  __shared__ Fit fit;
  extern __shared__ int tab[];  // adelta[out_w], bdelta[out_w], X0[rows_per], Y0[rows_per]
  int* adelta = tab;
  int* bdelta = tab + out_w;
  int* X0 = tab + 2 * out_w;
  int* Y0 = tab + 2 * out_w + rows_per;
  const int face = blockIdx.x;
  const int y_beg = blockIdx.y * rows_per, y_end = min(out_h, y_beg + rows_per);  // this CTA's rows of the face
  const int tid = threadIdx.x;
  const int nc = FAST ? C : rtC;

  cta_fit(fit, face, F, face2frame, marks, N, tmpl, five_point, M_in, M_out, blockIdx.y == 0, tid);
  __syncthreads();
  uint8_t* crop = crops + (size_t)face * out_h * out_w * nc;
  if (!fit.ok) {
    for (int i = y_beg * out_w * nc + tid; i < y_end * out_w * nc; i += kAlignThreads) crop[i] = 0;
    return;
  }
  for (int x = tid; x < out_w; x += kAlignThreads) { adelta[x] = tab_ad(fit, x); bdelta[x] = tab_bd(fit, x); }
  for (int y = y_beg + tid; y < y_end; y += kAlignThreads) { X0[y - y_beg] = tab_X0(fit, y); Y0[y - y_beg] = tab_Y0(fit, y); }
  __syncthreads();

  const uint8_t* frame = frames + (size_t)face2frame[face] * H * W * nc;
  const size_t row = (size_t)W * nc;

  if (FAST) {
    // C == 3, out_w % 4 == 0: 4 pixels (12 bytes) per thread per step
    const uint8_t* safe_end = frames + (size_t)F * H * W * 3 - 16;  // last address load6 may start from
    const int groups_per_row = out_w >> 2;
    const int n_groups = (y_end - y_beg) * groups_per_row;
    for (int g = tid; g < n_groups; g += kAlignThreads) {
      const int yl = g / groups_per_row;
      const int y = y_beg + yl;
      const int xg = (g - yl * groups_per_row) << 2;
      const int bx = X0[yl], by = Y0[yl];
      uint32_t outw[3] = {0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int X = (bx + adelta[xg + j]) >> 5;
        const int Y = (by + bdelta[xg + j]) >> 5;
        const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
        const int fx = X & 31, fy = Y & 31;
        const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
        const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
        int c0 = 0, c1 = 0, c2 = 0;
        if (sx >= 0 && sy >= 0 && sx + 1 < W && sy + 1 < H) {
          const uint8_t* p0 = frame + (size_t)sy * row + (size_t)sx * 3;
          const uint8_t* p1 = p0 + row;
          uint64_t v0, v1;
          if (p1 <= safe_end) {
            v0 = load6(p0);
            v1 = load6(p1);
          } else {
            v0 = 0; v1 = 0;
            for (int k = 5; k >= 0; --k) { v0 = (v0 << 8) | p0[k]; v1 = (v1 << 8) | p1[k]; }
          }
          c0 = w00 * (int)(v0 & 255) + w01 * (int)((v0 >> 24) & 255) + w10 * (int)(v1 & 255) + w11 * (int)((v1 >> 24) & 255);
          c1 = w00 * (int)((v0 >> 8) & 255) + w01 * (int)((v0 >> 32) & 255) + w10 * (int)((v1 >> 8) & 255) + w11 * (int)((v1 >> 32) & 255);
          c2 = w00 * (int)((v0 >> 16) & 255) + w01 * (int)((v0 >> 40) & 255) + w10 * (int)((v1 >> 16) & 255) + w11 * (int)((v1 >> 40) & 255);
        } else if (sx >= -1 && sy >= -1 && sx < W && sy < H) {
          // partially inside: out-of-image taps contribute the border value 0
          const bool x0in = sx >= 0, x1in = sx + 1 < W, y0in = sy >= 0, y1in = sy + 1 < H;
          const uint8_t* p0 = frame + (ptrdiff_t)sy * (ptrdiff_t)row + (ptrdiff_t)sx * 3;
          const uint8_t* p1 = p0 + row;
          int t[4][3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            t[0][ch] = (y0in && x0in) ? p0[ch] : 0;
            t[1][ch] = (y0in && x1in) ? p0[3 + ch] : 0;
            t[2][ch] = (y1in && x0in) ? p1[ch] : 0;
            t[3][ch] = (y1in && x1in) ? p1[3 + ch] : 0;
          }
          c0 = w00 * t[0][0] + w01 * t[1][0] + w10 * t[2][0] + w11 * t[3][0];
          c1 = w00 * t[0][1] + w01 * t[1][1] + w10 * t[2][1] + w11 * t[3][1];
          c2 = w00 * t[0][2] + w01 * t[1][2] + w10 * t[2][2] + w11 * t[3][2];
        }
        const uint32_t o0 = (uint32_t)((c0 + 16384) >> 15), o1 = (uint32_t)((c1 + 16384) >> 15),
                       o2 = (uint32_t)((c2 + 16384) >> 15);
        // pixel j occupies bytes 3j..3j+2 of the 12-byte group
        const int b = 3 * j;
        outw[b >> 2] |= o0 << (8 * (b & 3));
        outw[(b + 1) >> 2] |= o1 << (8 * ((b + 1) & 3));
        outw[(b + 2) >> 2] |= o2 << (8 * ((b + 2) & 3));
      }
      uint32_t* dst = reinterpret_cast<uint32_t*>(crop + ((size_t)y * out_w + xg) * 3);
      dst[0] = outw[0]; dst[1] = outw[1]; dst[2] = outw[2];
    }
  } else {
    const int n_px = (y_end - y_beg) * out_w;
    for (int i = tid; i < n_px; i += kAlignThreads) {
      const int yl = i / out_w, x = i - yl * out_w;
      const int X = (X0[yl] + adelta[x]) >> 5;
      const int Y = (Y0[yl] + bdelta[x]) >> 5;
      const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
      const int fx = X & 31, fy = Y & 31;
      const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
      const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
      const bool x0in = sx >= 0 && sx < W, x1in = sx + 1 >= 0 && sx + 1 < W;
      const bool y0in = sy >= 0 && sy < H, y1in = sy + 1 >= 0 && sy + 1 < H;
      for (int ch = 0; ch < nc; ++ch) {
        const ptrdiff_t o = (ptrdiff_t)sy * (ptrdiff_t)row + (ptrdiff_t)sx * nc + ch;
        const int p00 = (y0in && x0in) ? frame[o] : 0;
        const int p01 = (y0in && x1in) ? frame[o + nc] : 0;
        const int p10 = (y1in && x0in) ? frame[o + (ptrdiff_t)row] : 0;
        const int p11 = (y1in && x1in) ? frame[o + (ptrdiff_t)row + nc] : 0;
        crop[((size_t)(y_beg + yl) * out_w + x) * nc + ch] = (uint8_t)((w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 16384) >> 15);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor maps over the frames for the eight box classes, cached per (handle, buffer, shape): encoding costs microseconds
struct MapsEntry { const fld_handle* h; const void* frames; int F, H, W; AlignMaps maps; };
std::mutex g_maps_mu;
std::vector<MapsEntry> g_maps_cache;

int get_align_maps(const fld_handle* h, const uint8_t* frames, int F, int H, int W, AlignMaps* out) {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  for (const auto& e : g_maps_cache)
    if (e.h == h && e.frames == frames && e.F == F && e.H == H && e.W == W) { *out = e.maps; return FLD_OK; }
  EncodeTiledFn enc = (EncodeTiledFn)h->encode_tiled;
  MapsEntry e;
  e.h = h; e.frames = frames; e.F = F; e.H = H; e.W = W;
  memset(&e.maps, 0, sizeof(e.maps));
  for (int k = 0; k < kNumCls; ++k) {
    // the frame bytes as 32-bit elements: [F][H][W*3/4]; out-of-range elements are zero-filled (BORDER_CONSTANT 0)
    cuuint64_t dims[3] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)H * W * 3};
    cuuint32_t box[3] = {(cuuint32_t)cls_bw(k) / 4, (cuuint32_t)cls_side(k), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&e.maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(frames), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fld_set_error("cuTensorMapEncodeTiled(align frames, class %d) failed: %d", k, (int)r); return FLD_ERR_CUDA; }
  }
  if (g_maps_cache.size() >= 32) g_maps_cache.erase(g_maps_cache.begin());
  g_maps_cache.push_back(e);
  *out = e.maps;
  return FLD_OK;
}

size_t align_scratch_bytes(int B) { return (size_t)std::max(B, 0) * (sizeof(PreFit) + 2 * sizeof(int32_t)) + 64; }

int launch_align(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                 const float* marks, int N, const double* tmpl, int Nt, int five_point, const double* M_in, int B,
                 int out_h, int out_w, double* M_out, uint8_t* crops, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;   // empty batch: nothing to read or write (empty tensors have null data pointers)
  FLD_REQUIRE(frames && face2frame && crops, "fld_align: null pointer");
  FLD_REQUIRE(C == 1 || C == 3 || C == 4, "fld_align: C must be 1, 3 or 4 (got %d)", C);
  FLD_REQUIRE(F > 0 && H > 1 && W > 1 && out_h > 0 && out_w > 0 && B >= 0, "fld_align: bad shape");
  FLD_REQUIRE(out_h <= 4096 && out_w <= 4096, "fld_align: output larger than 4096");
  if (!M_in) {
    FLD_REQUIRE(marks && tmpl, "fld_align: null marks/template");
    if (five_point) FLD_REQUIRE(N == 68 && Nt == 5, "fld_align: five_point mode needs N=68, Nt=5 (got %d, %d)", N, Nt);
    else FLD_REQUIRE(N == Nt && N >= 2, "fld_align: N (%d) must equal Nt (%d) and be >= 2", N, Nt);
  }
  static const bool tile_off = getenv("FLD_ALIGN_TILE_OFF") != nullptr;
  const bool tile_ok = !tile_off && C == 3 && h->encode_tiled && out_w % kTile == 0 && out_h % kTile == 0 && out_w <= kMaxOut &&
                       out_h <= kMaxOut && ((size_t)W * 3) % 16 == 0 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(crops) & 3) == 0 && H < 30000 && W < 30000;
  if (tile_ok) {
    AlignMaps maps;
    rc = get_align_maps(h, frames, F, H, W, &maps);
    if (rc) return rc;
    TileArgs a;
    a.frames = frames; a.face2frame = face2frame; a.marks = marks; a.tmpl = tmpl; a.M_in = M_in; a.M_out = M_out; a.crops = crops;
    a.F = F; a.H = H; a.W = W; a.N = N; a.five_point = five_point; a.out_h = out_h; a.out_w = out_w;
    // split a face's tile rows over several CTAs while the grid would otherwise be only a few waves deep
    const int tiles_y = out_h / kTile;
    // measured (4-warp CTAs, fit fused): 1024 faces 84 / 66 / 64 / 77 us with 1 / 2 / 4 / 7 CTAs per face, 2047 faces 125 / 109 / 117 / 146,
    // 4096 faces 202 / 187 / 220 / -: aim at ~4 k CTAs, at least two per face (the per-CTA fit + tables dominate beyond that)
    a.ysplit = (int)std::max(1ll, std::min((long long)tiles_y, std::max(2ll, (4096ll + B - 1) / B)));
    { const char* e = getenv("FLD_ALIGN_YSPLIT"); if (e && atoi(e) > 0) a.ysplit = std::min(tiles_y, atoi(e)); }
    a.ring_bytes = kRingBytes;
    a.min_bufs = 1; a.pair_max = 2;
    { const char* e = getenv("FLD_ALIGN_PAIR_MAX"); if (e && atoi(e) >= 0 && atoi(e) <= 8) a.pair_max = atoi(e); }
    { const char* e = getenv("FLD_ALIGN_MIN_BUFS"); if (e && atoi(e) >= 1 && atoi(e) <= 4) a.min_bufs = atoi(e); }
    { const char* e = getenv("FLD_ALIGN_RING_KB"); if (e && atoi(e) >= 4 && atoi(e) <= 200) a.ring_bytes = atoi(e) * 1024; }
    const size_t smem = (size_t)a.ring_bytes + 128 + 32;
    FLD_CUDA(cudaFuncSetAttribute(align_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLD_CUDA(cudaFuncSetAttribute(align_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    a.fits = nullptr; a.perm = nullptr;
    // ordered mode: worth its two small extra launches once the grid is several waves deep
    int order_min = 2048;
    { const char* e = getenv("FLD_ALIGN_ORDER_MIN"); if (e) order_min = atoi(e); }
    if (scratch && B >= order_min) {
      FLD_REQUIRE(scratch_bytes >= align_scratch_bytes(B), "fld_align: scratch of %zu bytes, need %zu", scratch_bytes, align_scratch_bytes(B));
      FLD_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "fld_align: scratch must be 16-byte aligned");
      PreFit* fits = reinterpret_cast<PreFit*>(scratch);
      int32_t* keys = reinterpret_cast<int32_t*>(fits + B);
      int32_t* perm = keys + B;
      align_fit_kernel<<<fld_div_up(B, kFitWarps), 32 * kFitWarps, 0, st>>>(face2frame, F, marks, N, tmpl, five_point, M_in, M_out, B, a.ring_bytes, fits, keys);
      FLD_LAUNCHED();
      // the two dependents start under programmatic serialization: their launch latency and prologues overlap the predecessor
      cudaLaunchAttribute pdl[1];
      pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      pdl[0].val.programmaticStreamSerializationAllowed = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = pdl; cfg.numAttrs = 1;
      FLD_CUDA(cudaLaunchKernelEx(&cfg, align_order_kernel, (const int32_t*)keys, B, perm));
      FLD_LAUNCHED();
      a.fits = fits; a.perm = perm;
      a.ysplit = 1;
      { const char* e = getenv("FLD_ALIGN_YSPLIT"); if (e && atoi(e) > 0) a.ysplit = std::min(tiles_y, atoi(e)); }
      cfg.gridDim = dim3(B, a.ysplit); cfg.blockDim = dim3(kTileThreads); cfg.dynamicSmemBytes = smem;
      FLD_CUDA(cudaLaunchKernelEx(&cfg, align_tile_kernel, maps, a));
      FLD_LAUNCHED();
      return FLD_OK;
    }
    align_tile_kernel<<<dim3(B, a.ysplit), kTileThreads, smem, st>>>(maps, a);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  // small batches: split each face over several CTAs (row blocks) so that the grid covers the machine
  int rb = 1;
  while (rb < 8 && (long long)B * rb < 4ll * h->sm_count && out_h / (rb * 2) >= 8) rb *= 2;
  const int rows_per = fld_div_up(out_h, rb);
  const dim3 grid(B, fld_div_up(out_h, rows_per));
  const size_t smem = (size_t)(2 * out_w + 2 * rows_per) * sizeof(int);
  const bool fast = (C == 3) && (out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(crops) & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(frames) & 7) == 0) && ((size_t)F * H * W * 3 >= 32);
  if (fast) {
    align_warp_kernel<3, true><<<grid, kAlignThreads, smem, st>>>(frames, F, H, W, face2frame, marks, N, tmpl, Nt, five_point,
                                                               M_in, M_out, crops, out_h, out_w, 3, rows_per);
  } else {
    align_warp_kernel<1, false><<<grid, kAlignThreads, smem, st>>>(frames, F, H, W, face2frame, marks, N, tmpl, Nt, five_point,
                                                                M_in, M_out, crops, out_h, out_w, C, rows_per);
  }
  FLD_LAUNCHED();
  return FLD_OK;
}

}  // namespace

extern "C" int fld_align(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                         const float* marks, int N, const double* tmpl, int Nt, int five_point, int B, int out_h, int out_w,
                         double* M_out, uint8_t* crops, fld_stream stream) {
  return launch_align(h, frames, F, H, W, C, face2frame, marks, N, tmpl, Nt, five_point, nullptr, B, out_h, out_w, M_out,
                      crops, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int fld_warp_affine(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                               const double* M, int B, int out_h, int out_w, uint8_t* crops, fld_stream stream) {
  if (!M) { fld_set_error("fld_warp_affine: null M"); return FLD_ERR_INVALID; }
  return launch_align(h, frames, F, H, W, C, face2frame, nullptr, 0, nullptr, 0, 0, M, B, out_h, out_w, nullptr, crops,
                      nullptr, 0, (cudaStream_t)stream);
}

extern "C" size_t fld_align_scratch_bytes(fld_handle* h, int B) {
  (void)h;
  return align_scratch_bytes(B);
}

extern "C" int fld_align_ordered(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                                 const float* marks, int N, const double* tmpl, int Nt, int five_point, int B, int out_h, int out_w,
                                 double* M_out, uint8_t* crops, void* scratch, size_t scratch_bytes, fld_stream stream) {
  if (!scratch) { fld_set_error("fld_align_ordered: null scratch"); return FLD_ERR_INVALID; }
  return launch_align(h, frames, F, H, W, C, face2frame, marks, N, tmpl, Nt, five_point, nullptr, B, out_h, out_w, M_out,
                      crops, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int fld_warp_affine_ordered(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                                       const double* M, int B, int out_h, int out_w, uint8_t* crops, void* scratch,
                                       size_t scratch_bytes, fld_stream stream) {
  if (!M) { fld_set_error("fld_warp_affine_ordered: null M"); return FLD_ERR_INVALID; }
  if (!scratch) { fld_set_error("fld_warp_affine_ordered: null scratch"); return FLD_ERR_INVALID; }
  return launch_align(h, frames, F, H, W, C, face2frame, nullptr, 0, nullptr, 0, 0, M, B, out_h, out_w, nullptr, crops,
                      scratch, scratch_bytes, (cudaStream_t)stream);
}
