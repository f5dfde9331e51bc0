// decode.cu — landmark decodes: regression scale-back, per-pixel class map, heat-map centroids.
//
// Replaces reference prediction.py:88-94 (regression decode), prediction.py:209 (argmax over
// classes) and utils/metrics.py:46-109 (get_average_xy / transfer_xy_coord / transfer_target).
// All three are HBM-bound streaming reductions: coalesced loads, warp/block partials, no GEMM.
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"


namespace {

// ---------------------------------------------------------------- regression decode (prediction.py:88-94)
__global__ void regress_decode_kernel(const float* __restrict__ out136, int stride, const int32_t* __restrict__ fb, int B,
                                      float* __restrict__ marks, unsigned long long* __restrict__ marks_u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 136) return;
  const int b = i / 136, j = i - b * 136;
  const int x0 = fb[4 * b + 0], y0 = fb[4 * b + 1], x1 = fb[4 * b + 2];
  // marks *= (x1-x0); marks[:,0] += x0; marks[:,1] += y0   (float32, separate mul / add)
  float v = __fmul_rn(out136[(size_t)b * stride + j], (float)(x1 - x0));
  v = __fadd_rn(v, (float)((j & 1) ? y0 : x0));
  marks[i] = v;
  if (marks_u) marks_u[i] = (v > 0.f) ? (unsigned long long)v : 0ull;  // astype(np.uint): trunc toward 0, negatives -> 0
}

// ---------------------------------------------------------------- class map (prediction.py:209)
constexpr int kCmThreads = 128;
__global__ void __launch_bounds__(kCmThreads)
classmap_kernel(const float* __restrict__ scores, long long n_px, int L, long long* __restrict__ cmap) {
  extern __shared__ float tile[];  // [kCmThreads][L+1]
  const long long px0 = (long long)blockIdx.x * kCmThreads;
  const int npx = (int)min((long long)kCmThreads, n_px - px0);
  const float* src = scores + px0 * L;
  const int n = npx * L;
  const int ld = L | 1;  // odd pitch: conflict-free row scans
  for (int i = threadIdx.x; i < n; i += kCmThreads) {
    const int p = i / L, c = i - p * L;
    tile[p * ld + c] = src[i];
  }
  __syncthreads();
  if (threadIdx.x < npx) {
    const float* row = tile + threadIdx.x * ld;
    float best = row[0];
    int bi = 0;
    for (int c = 1; c < L; ++c) {
      const float v = row[c];
      if (v > best) { best = v; bi = c; }  // first maximum wins (numpy argmax)
    }
    cmap[px0 + threadIdx.x] = bi;
  }
}

// L % 4 == 0: 16-byte loads, 16-byte shared-memory stores and row scans (pitch L floats: a quarter warp of float4 accesses at a
// 4L-byte pitch hits 32 distinct banks when L/4 is odd, e.g. 68 classes) -- a quarter of the memory instructions
__global__ void __launch_bounds__(kCmThreads)
classmap_vec4_kernel(const float4* __restrict__ scores, long long n_px, int L4, long long* __restrict__ cmap) {
  extern __shared__ float4 tile4[];  // [kCmThreads][L4]
  const long long px0 = (long long)blockIdx.x * kCmThreads;
  const int npx = (int)min((long long)kCmThreads, n_px - px0);
  const float4* src = scores + px0 * L4;
  const int n = npx * L4;
  for (int i = threadIdx.x; i < n; i += kCmThreads) tile4[i] = __ldg(src + i);
  __syncthreads();
  if (threadIdx.x < npx) {
    const float4* row = tile4 + threadIdx.x * L4;
    float best = -INFINITY;
    int bi = 0;
    for (int q = 0; q < L4; ++q) {
      const float4 v = row[q];
      if (q == 0) { best = v.x; bi = 0; } else if (v.x > best) { best = v.x; bi = 4 * q; }   // first maximum wins (numpy argmax)
      if (v.y > best) { best = v.y; bi = 4 * q + 1; }
      if (v.z > best) { best = v.z; bi = 4 * q + 2; }
      if (v.w > best) { best = v.w; bi = 4 * q + 3; }
    }
    cmap[px0 + threadIdx.x] = bi;
  }
}

// ---------------------------------------------------------------- heat-map centroids (utils/metrics.py:46-80)
struct Cand { float v; int idx; };
__device__ __forceinline__ bool better(float av, int ai, float bv, int bi) {
  return av > bv || (av == bv && ai > bi);  // ties -> higher flat index (stable ascending sort, take the tail)
}

// partial[((b*S + s)*L + l)*n + k] : per-slab top-n candidates (unsorted), idx = -1 for empty
template <int NMAX>
__global__ void __launch_bounds__(1024)   // blockDim = R * L can reach 1024 (L up to 1024): keep every variant launchable
topn_partial_kernel(const float* __restrict__ hm, int HW, int L, int R, int S, int n,
                                    Cand* __restrict__ partial) {
  const int b = blockIdx.x, s = blockIdx.y;
  const int t = threadIdx.x;
  const int l = t % L, r = t / L;   // blockDim.x == R * L exactly: every thread reaches the barrier below
  Cand top[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) { top[k].v = -INFINITY; top[k].idx = -1; }
  int mn = 0;  // position of the current worst entry
  float wv = -INFINITY;   // ... and a register copy of it: the common (no insertion) path is one compare per value
  int wi = -1;
  const int per = (HW + S - 1) / S;
  const int p0 = s * per, p1 = min(HW, p0 + per);
  const float* base = hm + (size_t)b * HW * L;
  auto consider = [&](float v, int p) {
    if (better(v, p, wv, wi)) {
#pragma unroll
      for (int k = 0; k < NMAX; ++k) if (k == mn) { top[k].v = v; top[k].idx = p; }
      mn = 0;  // rescan for the worst
#pragma unroll
      for (int k = 1; k < NMAX; ++k) if (k < n && better(top[mn].v, top[mn].idx, top[k].v, top[k].idx)) mn = k;
      wv = top[0].v; wi = top[0].idx;
#pragma unroll
      for (int k = 1; k < NMAX; ++k) if (k == mn) { wv = top[k].v; wi = top[k].idx; }
    }
  };
  // Pixels are visited in DESCENDING index order: among equal values the higher index wins, so a tie arriving later never
  // displaces an entry (heat maps with long runs of equal values, e.g. saturated softmax outputs, stay on the one-compare path).
  int cnt = (p0 + r < p1) ? (p1 - (p0 + r) + R - 1) / R : 0;
  int p = p0 + r + (cnt - 1) * R;
  for (; cnt >= 4; cnt -= 4, p -= 4 * R) {  // four independent loads in flight per thread
    const float v0 = __ldg(base + (size_t)p * L + l), v1 = __ldg(base + (size_t)(p - R) * L + l);
    const float v2 = __ldg(base + (size_t)(p - 2 * R) * L + l), v3 = __ldg(base + (size_t)(p - 3 * R) * L + l);
    consider(v0, p); consider(v1, p - R); consider(v2, p - 2 * R); consider(v3, p - 3 * R);
  }
  for (; cnt > 0; --cnt, p -= R) consider(__ldg(base + (size_t)p * L + l), p);
  if (NMAX > 16) {  // large n: every (slab, r) lane writes its n candidates; the merge kernel selects among S*R*n
    Cand* dst = partial + ((((size_t)b * S + s) * R + r) * L + l) * n;
    for (int k = 0; k < n; ++k) dst[k] = top[k];
    return;
  }
  // n <= 16: block-level selection over the R pixel lanes, so that only ONE list per (slab, class) reaches the (serial) merge kernel
  extern __shared__ Cand cand_s[];   // [R][L][n]
#pragma unroll
  for (int k = 0; k < NMAX; ++k)
    if (k < n) cand_s[((size_t)r * L + l) * n + k] = top[k];
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < R; ++rr)
      for (int k = 0; k < n; ++k) {
        const Cand c = cand_s[((size_t)rr * L + l) * n + k];
        if (c.idx >= 0) consider(c.v, c.idx);
      }
    Cand* dst = partial + (((size_t)b * S + s) * L + l) * n;
#pragma unroll
    for (int k = 0; k < NMAX; ++k)
      if (k < n) dst[k] = top[k];
  }
}

__global__ void topn_merge_kernel(const Cand* __restrict__ partial, int W, int L, int SR, int n, int B, float thresh,
                                  double* __restrict__ xy, Cand* __restrict__ sel_scratch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, l)
  if (i >= B * L) return;
  const int b = i / L, l = i - b * L;
  Cand* sel = sel_scratch + (size_t)i * n;  // selected, kept sorted descending by (v, idx)
  int cnt = 0;
  for (int sr = 0; sr < SR; ++sr) {
    const Cand* src = partial + (((size_t)b * SR + sr) * L + l) * n;
    for (int k = 0; k < n; ++k) {
      const Cand c = src[k];
      if (c.idx < 0) continue;
      if (cnt == n && !better(c.v, c.idx, sel[n - 1].v, sel[n - 1].idx)) continue;
      int pos = (cnt < n) ? cnt : n - 1;
      while (pos > 0 && better(c.v, c.idx, sel[pos - 1].v, sel[pos - 1].idx)) { sel[pos] = sel[pos - 1]; --pos; }
      sel[pos] = c;
      if (cnt < n) ++cnt;
    }
  }
  // accumulate in ascending order like metrics.py:70-74: hsum in float32 (np.float32 scalars),
  // i0/i1 in float64 (np.int64 * np.float32 -> float64)
  float hsum = 0.f;
  double i0 = 0.0, i1 = 0.0;
  for (int k = cnt - 1; k >= 0; --k) {
    const float h = sel[k].v;
    const int row = sel[k].idx / W, col = sel[k].idx - row * W;
    hsum = __fadd_rn(hsum, h);
    i0 = __dadd_rn(i0, __dmul_rn((double)row, (double)h));
    i1 = __dadd_rn(i1, __dmul_rn((double)col, (double)h));
  }
  i0 = i0 / (double)hsum;
  i1 = i1 / (double)hsum;
  if (__fdiv_rn(hsum, (float)n) <= thresh) { i0 = -1.0; i1 = -1.0; }  // metrics.py:78-79
  xy[(size_t)i * 2 + 0] = i1;
  xy[(size_t)i * 2 + 1] = i0;
}

// full soft-centroid (metrics.py:58-64): per-slab fp64 partial sums, then merge
__global__ void soft_partial_kernel(const float* __restrict__ hm, int HW, int W, int L, int R, int S,
                                    double* __restrict__ partial) {
  const int b = blockIdx.x, s = blockIdx.y;
  const int t = threadIdx.x;
  const int l = t % L, r = t / L;
  if (r >= R) return;
  const int per = (HW + S - 1) / S;
  const int p0 = s * per, p1 = min(HW, p0 + per);
  const float* base = hm + (size_t)b * HW * L;
  double sh = 0, sx = 0, sy = 0;
  int p = p0 + r;
  for (; p + 3 * R < p1; p += 4 * R) {  // four independent loads in flight per thread
    const float v[4] = {__ldg(base + (size_t)p * L + l), __ldg(base + (size_t)(p + R) * L + l),
                        __ldg(base + (size_t)(p + 2 * R) * L + l), __ldg(base + (size_t)(p + 3 * R) * L + l)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pp = p + q * R;
      const int row = pp / W, col = pp - row * W;
      sh += (double)v[q]; sx += (double)v[q] * col; sy += (double)v[q] * row;
    }
  }
  for (; p < p1; p += R) {
    const double v = (double)__ldg(base + (size_t)p * L + l);
    const int row = p / W, col = p - row * W;
    sh += v; sx += v * col; sy += v * row;
  }
  double* dst = partial + ((((size_t)b * S + s) * R + r) * L + l) * 3;
  dst[0] = sh; dst[1] = sx; dst[2] = sy;
}

__global__ void soft_merge_kernel(const double* __restrict__ partial, int HW, int L, int SR, int B, float thresh,
                                  double* __restrict__ xy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, l = i - b * L;
  double sh = 0, sx = 0, sy = 0;
  for (int sr = 0; sr < SR; ++sr) {
    const double* src = partial + (((size_t)b * SR + sr) * L + l) * 3;
    sh += src[0]; sx += src[1]; sy += src[2];
  }
  double x = sx / sh, y = sy / sh;
  if ((float)sh / (float)HW <= thresh) { x = -1.0; y = -1.0; }
  xy[(size_t)i * 2 + 0] = x;
  xy[(size_t)i * 2 + 1] = y;
}

// ---- L % 4 == 0 variants: thread (r, q) owns classes 4q .. 4q+3 of pixel lane r and reads them with ONE 16-byte load; a warp
// reads 512 contiguous bytes.  Same partial layout as the scalar kernels (the merge kernels are shared).
__global__ void soft_partial_vec4_kernel(const float4* __restrict__ hm, int HW, int W, int L4, int R, int S, double* __restrict__ partial) {
  const int b = blockIdx.x, s = blockIdx.y;
  const int t = threadIdx.x;
  const int q = t % L4, r = t / L4;   // blockDim.x == R * L4 exactly: no thread leaves before the barrier
  const int per = (HW + S - 1) / S;
  const int p0 = s * per, p1 = min(HW, p0 + per);
  const float4* base = hm + (size_t)b * HW * L4 + q;
  // per-thread sums in fp32 (a thread adds ~200 values; the reference itself sums float32 arrays, metrics.py:58-64), fp64 from the
  // block-level reduction on
  float sh[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0}, sy[4] = {0, 0, 0, 0};
  auto acc = [&](const float4& v, int p) {
    const int row = p / W, col = p - row * W;
    const float fc = (float)col, fr = (float)row;
    sh[0] += v.x; sx[0] = fmaf(v.x, fc, sx[0]); sy[0] = fmaf(v.x, fr, sy[0]);
    sh[1] += v.y; sx[1] = fmaf(v.y, fc, sx[1]); sy[1] = fmaf(v.y, fr, sy[1]);
    sh[2] += v.z; sx[2] = fmaf(v.z, fc, sx[2]); sy[2] = fmaf(v.z, fr, sy[2]);
    sh[3] += v.w; sx[3] = fmaf(v.w, fc, sx[3]); sy[3] = fmaf(v.w, fr, sy[3]);
  };
  int p = p0 + r;
  for (; p + 3 * R < p1; p += 4 * R) {  // four independent 16-byte loads in flight per thread
    const float4 v0 = __ldg(base + (size_t)p * L4), v1 = __ldg(base + (size_t)(p + R) * L4);
    const float4 v2 = __ldg(base + (size_t)(p + 2 * R) * L4), v3 = __ldg(base + (size_t)(p + 3 * R) * L4);
    acc(v0, p); acc(v1, p + R); acc(v2, p + 2 * R); acc(v3, p + 3 * R);
  }
  for (; p < p1; p += R) acc(__ldg(base + (size_t)p * L4), p);
  // block-level reduction over the R pixel lanes: one partial triple per (slab, class) reaches the merge kernel
  extern __shared__ double red[];   // [R][L][3]
  const int L = L4 * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double* d = red + ((size_t)r * L + 4 * q + k) * 3;
    d[0] = (double)sh[k]; d[1] = (double)sx[k]; d[2] = (double)sy[k];
  }
  __syncthreads();
  if (t < L) {
    double a0 = 0, a1 = 0, a2 = 0;
    for (int rr = 0; rr < R; ++rr) {
      const double* d = red + ((size_t)rr * L + t) * 3;
      a0 += d[0]; a1 += d[1]; a2 += d[2];
    }
    double* dst = partial + (((size_t)b * S + s) * L + t) * 3;
    dst[0] = a0; dst[1] = a1; dst[2] = a2;
  }
}

// n <= 4, L % 4 == 0.  Thread (r, q) reads classes 4q .. 4q+3 of pixel lane r (one 16-byte load per pixel).  The running top four
// of a class are NOT per-thread lists (round 1 / early round 2: every thread kept 4 x 4 sorted entries in registers; a warp took the
// ~18-instruction insertion path whenever ANY of its 128 lists inserted, which with ~560 pixels per thread was nearly every step —
// ncu: 13 of 32 lanes active on average, 65 warp instructions per 16-byte load, 0.39 - 0.48 of HBM).  They are four 64-bit slots per
// class in SHARED memory, common to the CTA's pixel lanes: key = (order-preserving bits of the value) << 32 | flat pixel index, so
// a larger key is exactly the reference's "larger value, ties to the higher index" (metrics.py:66-77).  A value is first compared
// with the class's lower bound s_thr (the value in slot 3: that element lost to three larger keys, so four elements >= it exist);
// only values that reach it go on — about 4 ln(pixels / 4) per class and CTA instead of per thread — and bubble down the slots with
// atomicMax: each step leaves the larger key in the slot and carries the smaller on.  The multiset {slots} + {carried} always equals
// the keys inserted so far minus keys that lost four times, so the final slots are the slab's top four whatever the interleaving.
// (Replacing the smallest slot with one compare-and-swap instead measured 2x slower: 82 registers, retries during the warm-up.)
__device__ __forceinline__ unsigned long long topn_key(float v, int p) {
  const uint32_t b = __float_as_uint(v);
  const uint32_t ou = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)ou << 32) | (uint32_t)p;
}
__device__ __forceinline__ float topn_key_value(unsigned long long key) {
  const uint32_t ou = (uint32_t)(key >> 32);
  return __uint_as_float((ou & 0x80000000u) ? (ou ^ 0x80000000u) : ~ou);
}
__device__ __noinline__ void topn4_insert(unsigned long long* slots, float* thr, float v, int p) {
  unsigned long long key = topn_key(v, p);
  if (key <= *reinterpret_cast<volatile unsigned long long*>(&slots[3])) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long old = atomicMax(&slots[k], key);
    key = old < key ? old : key;
    if (key == 0ull) break;               // the slot was empty: nothing to carry on
  }
  const unsigned long long s3 = *reinterpret_cast<volatile unsigned long long*>(&slots[3]);
  if (s3 != 0ull) {
    const float w = topn_key_value(s3);
    if (w >= 0.f) atomicMax(reinterpret_cast<int*>(thr), __float_as_int(w));
    else atomicMin(reinterpret_cast<unsigned*>(thr), __float_as_uint(w));
  }
}

// Work split: one CTA per (image, slab of `per` pixels), B * S CTAs = one wave of resident CTAs.  (Equal shares of the whole batch
// across exactly the resident CTAs, straddling image boundaries, measured SLOWER: 0.289 vs 0.256 ms — a second warm-up of the
// slots per CTA costs more than the 148-SM imbalance, which an HBM-bound kernel does not feel.)
__global__ void __launch_bounds__(512)
topn4_vec4_kernel(const float4* __restrict__ hm, int HW, int L4, int R, int S, int n, Cand* __restrict__ partial) {
  const int t = threadIdx.x;
  const int q = t % L4, r = t / L4;   // blockDim.x == R * L4 exactly
  const int L = L4 * 4;
  const int b = blockIdx.x, s = blockIdx.y;
  const int Sb = S;
  extern __shared__ __align__(16) unsigned long long s_slots[];   // [L][4] keys, then [L] floats: the lower bounds
  float* s_thr = reinterpret_cast<float*>(s_slots + (size_t)L * 4);
  const uint32_t thr_addr = (uint32_t)__cvta_generic_to_shared(&s_thr[4 * q]);
  unsigned long long* my_slots = s_slots + (size_t)(4 * q) * 4;
  const int per = (HW + S - 1) / S;
  const int p0 = s * per, p1 = min(HW, p0 + per);
  {
  for (int i = t; i < L * 4; i += blockDim.x) s_slots[i] = 0ull;
  for (int i = t; i < L; i += blockDim.x) s_thr[i] = -INFINITY;
  __syncthreads();
  const float4* base = hm + (size_t)b * HW * L4 + q;
  float thr[4];
#define FLD_TOPN4_LOAD_THR() \
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(thr[0]), "=f"(thr[1]), "=f"(thr[2]), "=f"(thr[3]) : "r"(thr_addr))
#define FLD_TOPN4_CONSIDER4(V4, P)                                                          \
  {                                                                                         \
    if ((V4).x >= thr[0]) topn4_insert(my_slots + 0, s_thr + 4 * q + 0, (V4).x, (P));       \
    if ((V4).y >= thr[1]) topn4_insert(my_slots + 4, s_thr + 4 * q + 1, (V4).y, (P));       \
    if ((V4).z >= thr[2]) topn4_insert(my_slots + 8, s_thr + 4 * q + 2, (V4).z, (P));       \
    if ((V4).w >= thr[3]) topn4_insert(my_slots + 12, s_thr + 4 * q + 3, (V4).w, (P));      \
  }
  int p = p0 + r;
  for (; p + 7 * R < p1; p += 8 * R) {  // eight independent 16-byte loads in flight per thread
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(base + (size_t)(p + k * R) * L4);
    FLD_TOPN4_LOAD_THR();
#pragma unroll
    for (int k = 0; k < 8; ++k) FLD_TOPN4_CONSIDER4(v[k], p + k * R)
  }
  FLD_TOPN4_LOAD_THR();
  for (; p < p1; p += R) {
    const float4 v = __ldg(base + (size_t)p * L4);
    FLD_TOPN4_CONSIDER4(v, p)
  }
#undef FLD_TOPN4_CONSIDER4
#undef FLD_TOPN4_LOAD_THR
  __syncthreads();
  // one thread per class: sort the four slots (descending) and emit the first n
  if (t < L) {
    unsigned long long k4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) k4[k] = s_slots[(size_t)t * 4 + k];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3 - a; ++c)
        if (k4[c] < k4[c + 1]) { const unsigned long long x = k4[c]; k4[c] = k4[c + 1]; k4[c + 1] = x; }
    Cand* dst = partial + (((size_t)b * Sb + s) * L + t) * n;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n) {
        if (k4[k] == 0ull) { dst[k].v = -INFINITY; dst[k].idx = -1; }
        else { dst[k].v = topn_key_value(k4[k]); dst[k].idx = (int)(uint32_t)k4[k]; }
      }
  }
  }
}

// Merge of the per-slab top-n lists of topn4_vec4_kernel (n <= 4): one thread per (image, class), the running top four as keys in
// registers (the generic merge keeps its selection in global scratch: 18 us for 64 x 68 lists), then the reference's accumulation.
__global__ void topn4_merge_kernel(const Cand* __restrict__ partial, int W, int L, int Sb, int n, int B, float thresh,
                                   double* __restrict__ xy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, l)
  if (i >= B * L) return;
  const int b = i / L, l = i - b * L;
  unsigned long long k4[4] = {0ull, 0ull, 0ull, 0ull};
  for (int j = 0; j < Sb; ++j) {
    const Cand* src = partial + (((size_t)b * Sb + j) * L + l) * n;
    for (int k = 0; k < n; ++k) {
      const Cand c = src[k];
      if (c.idx < 0) continue;
      unsigned long long key = topn_key(c.v, c.idx);
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (key > k4[a]) { const unsigned long long x = k4[a]; k4[a] = key; key = x; }
    }
  }
  // accumulate in ascending order like metrics.py:70-74: hsum in float32 (np.float32 scalars), i0/i1 in float64
  float hsum = 0.f;
  double i0 = 0.0, i1 = 0.0;
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    if (k >= n || k4[k] == 0ull) continue;
    const float h = topn_key_value(k4[k]);
    const int idx = (int)(uint32_t)k4[k];
    const int row = idx / W, col = idx - row * W;
    hsum = __fadd_rn(hsum, h);
    i0 = __dadd_rn(i0, __dmul_rn((double)row, (double)h));
    i1 = __dadd_rn(i1, __dmul_rn((double)col, (double)h));
  }
  i0 = i0 / (double)hsum;
  i1 = i1 / (double)hsum;
  if (__fdiv_rn(hsum, (float)n) <= thresh) { i0 = -1.0; i1 = -1.0; }  // metrics.py:78-79
  xy[(size_t)i * 2 + 0] = i1;
  xy[(size_t)i * 2 + 1] = i0;
}

}  // namespace

extern "C" int fld_decode_regress(fld_handle* h, const float* out136, int stride, const int32_t* faceboxes, int B,
                                  float* marks_f32, uint64_t* marks_u64, fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(out136 && faceboxes && marks_f32, "fld_decode_regress: null pointer");
  FLD_REQUIRE(stride >= 136 && B >= 0, "fld_decode_regress: stride must be >= 136");
  if (B == 0) return FLD_OK;
  const int n = B * 136;
  regress_decode_kernel<<<fld_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(out136, stride, faceboxes, B, marks_f32,
                                                                             (unsigned long long*)marks_u64);
  FLD_LAUNCHED();
  return FLD_OK;
}

extern "C" int fld_decode_classmap(fld_handle* h, const float* scores, int B, int hw, int L, int64_t* class_map,
                                   fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(scores && class_map, "fld_decode_classmap: null pointer");
  FLD_REQUIRE(B >= 0 && hw > 0 && L > 0 && L <= 384, "fld_decode_classmap: need 0 < L <= 384");
  if (B == 0) return FLD_OK;
  const long long n_px = (long long)B * hw;
  if (L % 4 == 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0 && !getenv("FLD_DECODE_SCALAR")) {
    const size_t smem4 = (size_t)kCmThreads * L * sizeof(float);
    if (smem4 > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(classmap_vec4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    const long long blocks4 = (n_px + kCmThreads - 1) / kCmThreads;
    FLD_REQUIRE(blocks4 < (1ll << 31), "fld_decode_classmap: too many pixels");
    classmap_vec4_kernel<<<(unsigned)blocks4, kCmThreads, smem4, (cudaStream_t)stream>>>((const float4*)scores, n_px, L / 4, (long long*)class_map);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  const size_t smem = (size_t)kCmThreads * (L | 1) * sizeof(float);
  if (smem > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(classmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n_px + kCmThreads - 1) / kCmThreads;
  FLD_REQUIRE(blocks < (1ll << 31), "fld_decode_classmap: too many pixels");
  classmap_kernel<<<(unsigned)blocks, kCmThreads, smem, (cudaStream_t)stream>>>(scores, n_px, L, (long long*)class_map);
  FLD_LAUNCHED();
  return FLD_OK;
}

namespace {
struct XyPlan { bool vec4; int R, S, threads; size_t bytes; };

// launch geometry + scratch bytes of the two-pass heat-map decode; deterministic per device
XyPlan xy_plan(const fld_handle* h, int B, int H, int W, int L, int n_points, bool vec4) {
  XyPlan p;
  const int HW = H * W;
  p.vec4 = vec4;
  const int Lt = vec4 ? L / 4 : L;                      // threads per pixel lane
  p.R = (Lt >= 256) ? 1 : fld_div_up(256, Lt);          // pixel lanes per CTA
  if (vec4 && n_points >= 1) {
    // top-n: the CTA's lanes share the four slots of a class, and the insertions (the expensive, divergent part) number ~4 ln(pixels
    // of the CTA / 4) per class and CTA — fewer, larger CTAs mean fewer insertions in total.  Measured on 64 x 232 x 232 x 68:
    // 8 / 12 / 16 / 20 / 24 / 28 / 30 lanes -> 0.311 / 0.262 / 0.248 / 0.240 / 0.224 / 0.215 / 0.212 ms.
    p.R = std::max(1, 512 / Lt);
    const char* e = getenv("FLD_TOPN_R");
    if (e && atoi(e) > 0 && atoi(e) * Lt <= 512) p.R = atoi(e);
  }
  p.threads = p.R * Lt;
  const int R = p.R, threads = p.threads;
  // slabs: B * S CTAs fill ONE wave of resident CTAs (a 2.05-wave grid measured 68 % efficient), at least 64 pixels per lane
  int occ = 0;
  if (n_points < 1) {
    if (vec4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, soft_partial_vec4_kernel, threads, (size_t)R * L * 3 * sizeof(double));
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, soft_partial_kernel, threads, 0);
  } else if (vec4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topn4_vec4_kernel, threads, (size_t)L * 36);
  else if (n_points <= 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topn_partial_kernel<4>, threads, (size_t)R * L * n_points * sizeof(Cand));
  else if (n_points <= 16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topn_partial_kernel<16>, threads, (size_t)R * L * n_points * sizeof(Cand));
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topn_partial_kernel<FLD_MAX_TOPN>, threads, 0);
  occ = max(occ, 1);
  int S = max(1, (h->sm_count * occ) / max(B, 1));
  S = max(1, min(S, HW / (64 * R) > 0 ? HW / (64 * R) : 1));
  p.S = min(S, 65535);
  if (n_points >= 1 && vec4) { p.bytes = (size_t)B * p.S * L * n_points * sizeof(Cand); return p; }
  if (n_points < 1) p.bytes = (size_t)B * p.S * R * L * 3 * sizeof(double);
  else p.bytes = (size_t)B * p.S * R * L * n_points * sizeof(Cand) + (size_t)B * L * n_points * sizeof(Cand);
  return p;
}
bool xy_vec4_shape(int L, int n_points) { return L % 4 == 0 && n_points <= 4 && L <= 128 && !getenv("FLD_DECODE_SCALAR"); }
}  // namespace

extern "C" size_t fld_decode_heatmap_scratch_bytes(fld_handle* h, int B, int H, int W, int L, int n_points) {
  if (!h || B <= 0 || H <= 0 || W <= 0 || L <= 0 || L > 1024 || n_points > FLD_MAX_TOPN) return 0;
  if (fld_enter(h)) return 0;
  size_t b = xy_plan(h, B, H, W, L, n_points, false).bytes;   // the pointer's alignment picks the variant at call time: cover both
  if (xy_vec4_shape(L, n_points)) b = std::max(b, xy_plan(h, B, H, W, L, n_points, true).bytes);
  return b + 256;
}

extern "C" int fld_decode_heatmap_xy(fld_handle* h, const float* hm, int B, int H, int W, int L, int n_points, double thresh,
                                     double* xy, void* scratch_in, size_t scratch_bytes, fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(hm && xy, "fld_decode_heatmap_xy: null pointer");
  FLD_REQUIRE(B >= 0 && H > 0 && W > 0 && L > 0 && L <= 1024, "fld_decode_heatmap_xy: need 0 < L <= 1024");
  FLD_REQUIRE(n_points <= FLD_MAX_TOPN, "fld_decode_heatmap_xy: n_points must be <= %d", FLD_MAX_TOPN);
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const bool vec4 = xy_vec4_shape(L, n_points) && (reinterpret_cast<uintptr_t>(hm) & 15) == 0;
  const XyPlan pl = xy_plan(h, B, H, W, L, n_points, vec4);
  const int R = pl.R, S = pl.S, threads = pl.threads;
  // caller-provided scratch (fld_decode_heatmap_scratch_bytes): decodes in flight on different streams never share partials
  FLD_REQUIRE(scratch_in, "fld_decode_heatmap_xy: null scratch (size it with fld_decode_heatmap_scratch_bytes)");
  char* scratch = (char*)(((uintptr_t)scratch_in + 255) & ~(uintptr_t)255);
  if ((size_t)(scratch - (char*)scratch_in) + pl.bytes > scratch_bytes) {
    fld_set_error("fld_decode_heatmap_xy: scratch %zu < required %zu", scratch_bytes, pl.bytes + 256);
    return FLD_ERR_WORKSPACE;
  }
  dim3 grid(B, S);
  const int nBL = B * L;
  if (n_points < 1) {
    if (vec4) soft_partial_vec4_kernel<<<grid, threads, (size_t)R * L * 3 * sizeof(double), st>>>((const float4*)hm, HW, W, L / 4, R, S, (double*)scratch);
    else soft_partial_kernel<<<grid, threads, 0, st>>>(hm, HW, W, L, R, S, (double*)scratch);
    FLD_LAUNCHED();
    soft_merge_kernel<<<fld_div_up(nBL, 128), 128, 0, st>>>((const double*)scratch, HW, L, vec4 ? S : S * R, B, (float)thresh, xy);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  const int n = n_points;
  FLD_REQUIRE(n <= HW, "fld_decode_heatmap_xy: n_points (%d) exceeds H*W (%d)", n, HW);
  const size_t part_bytes = (size_t)B * S * R * L * n * sizeof(Cand);
  Cand* partial = (Cand*)scratch;
  Cand* sel = (Cand*)((char*)scratch + part_bytes);
  const size_t csm = (size_t)R * L * n * sizeof(Cand);   // block-level candidate exchange (n <= 16)
  if (n <= 16 && csm > 48 * 1024) {
    FLD_CUDA(cudaFuncSetAttribute(topn_partial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
    FLD_CUDA(cudaFuncSetAttribute(topn_partial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  }
  if (vec4) {
    topn4_vec4_kernel<<<grid, threads, (size_t)L * 36, st>>>((const float4*)hm, HW, L / 4, R, S, n, partial);
    FLD_LAUNCHED();
    topn4_merge_kernel<<<fld_div_up(nBL, 128), 128, 0, st>>>(partial, W, L, S, n, B, (float)thresh, xy);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  if (n <= 4) topn_partial_kernel<4><<<grid, threads, csm, st>>>(hm, HW, L, R, S, n, partial);
  else if (n <= 16) topn_partial_kernel<16><<<grid, threads, csm, st>>>(hm, HW, L, R, S, n, partial);
  else topn_partial_kernel<FLD_MAX_TOPN><<<grid, threads, 0, st>>>(hm, HW, L, R, S, n, partial);
  FLD_LAUNCHED();
  topn_merge_kernel<<<fld_div_up(nBL, 128), 128, 0, st>>>(partial, W, L, n <= 16 ? S : S * R, n, B, (float)thresh, xy, sel);
  FLD_LAUNCHED();
  return FLD_OK;
}
