// decode.cu — landmark decodes: regression scale-back, per-pixel class map, heat-map centroids.
//
// Replaces reference prediction.py:88-94 (regression decode), prediction.py:209 (argmax over
// classes) and utils/metrics.py:46-109 (get_average_xy / transfer_xy_coord / transfer_target).
// All three are HBM-bound streaming reductions: coalesced loads, warp/block partials, no GEMM.
#include "common.cuh"

int fld_scratch(fld_handle* h, size_t bytes, void** out);  // api.cu: handle-owned scratch that grows on demand

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

// ---------------------------------------------------------------- heat-map centroids (utils/metrics.py:46-80)
struct Cand { float v; int idx; };
__device__ __forceinline__ bool better(float av, int ai, float bv, int bi) {
  return av > bv || (av == bv && ai > bi);  // ties -> higher flat index (stable ascending sort, take the tail)
}

// partial[((b*S + s)*L + l)*n + k] : per-slab top-n candidates (unsorted), idx = -1 for empty
template <int NMAX>
__global__ void topn_partial_kernel(const float* __restrict__ hm, int HW, int L, int R, int S, int n,
                                    Cand* __restrict__ partial) {
  const int b = blockIdx.x, s = blockIdx.y;
  const int t = threadIdx.x;
  const int l = t % L, r = t / L;
  if (r >= R) return;
  Cand top[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) { top[k].v = -INFINITY; top[k].idx = -1; }
  int mn = 0;  // position of the current worst entry
  const int per = (HW + S - 1) / S;
  const int p0 = s * per, p1 = min(HW, p0 + per);
  const float* base = hm + (size_t)b * HW * L;
  auto consider = [&](float v, int p) {
    if (better(v, p, top[mn].v, top[mn].idx)) {
#pragma unroll
      for (int k = 0; k < NMAX; ++k) if (k == mn) { top[k].v = v; top[k].idx = p; }
      mn = 0;  // rescan for the worst
#pragma unroll
      for (int k = 1; k < NMAX; ++k) if (k < n && better(top[mn].v, top[mn].idx, top[k].v, top[k].idx)) mn = k;
    }
  };
  int p = p0 + r;
  for (; p + 3 * R < p1; p += 4 * R) {  // four independent loads in flight per thread
    const float v0 = __ldg(base + (size_t)p * L + l), v1 = __ldg(base + (size_t)(p + R) * L + l);
    const float v2 = __ldg(base + (size_t)(p + 2 * R) * L + l), v3 = __ldg(base + (size_t)(p + 3 * R) * L + l);
    consider(v0, p); consider(v1, p + R); consider(v2, p + 2 * R); consider(v3, p + 3 * R);
  }
  for (; p < p1; p += R) consider(__ldg(base + (size_t)p * L + l), p);
  // every (slab, r) lane writes its n candidates; merge kernel selects among S*R*n
  Cand* dst = partial + ((((size_t)b * S + s) * R + r) * L + l) * n;
  for (int k = 0; k < n; ++k) dst[k] = top[k];
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
  const size_t smem = (size_t)kCmThreads * (L | 1) * sizeof(float);
  if (smem > 48 * 1024) FLD_CUDA(cudaFuncSetAttribute(classmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n_px + kCmThreads - 1) / kCmThreads;
  FLD_REQUIRE(blocks < (1ll << 31), "fld_decode_classmap: too many pixels");
  classmap_kernel<<<(unsigned)blocks, kCmThreads, smem, (cudaStream_t)stream>>>(scores, n_px, L, (long long*)class_map);
  FLD_LAUNCHED();
  return FLD_OK;
}

extern "C" int fld_decode_heatmap_xy(fld_handle* h, const float* hm, int B, int H, int W, int L, int n_points, double thresh,
                                     double* xy, fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(hm && xy, "fld_decode_heatmap_xy: null pointer");
  FLD_REQUIRE(B >= 0 && H > 0 && W > 0 && L > 0 && L <= 1024, "fld_decode_heatmap_xy: need 0 < L <= 1024");
  FLD_REQUIRE(n_points <= FLD_MAX_TOPN, "fld_decode_heatmap_xy: n_points must be <= %d", FLD_MAX_TOPN);
  if (B == 0) return FLD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int R = (L >= 256) ? 1 : fld_div_up(256, L);  // pixel lanes per CTA
  const int threads = R * L;
  // slabs: enough CTAs to cover the machine ~8x (the kernels are latency-bound: one strided stream per thread),
  // at least 64 pixels per lane
  int S = fld_div_up(8 * h->sm_count, B);
  S = max(1, min(S, HW / (64 * R) > 0 ? HW / (64 * R) : 1));
  S = min(S, 65535);
  dim3 grid(B, S);
  const int nBL = B * L;
  if (n_points < 1) {
    void* scratch;
    rc = fld_scratch(h, (size_t)B * S * R * L * 3 * sizeof(double), &scratch);
    if (rc) return rc;
    soft_partial_kernel<<<grid, threads, 0, st>>>(hm, HW, W, L, R, S, (double*)scratch);
    FLD_LAUNCHED();
    soft_merge_kernel<<<fld_div_up(nBL, 128), 128, 0, st>>>((const double*)scratch, HW, L, S * R, B, (float)thresh, xy);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  const int n = n_points;
  FLD_REQUIRE(n <= HW, "fld_decode_heatmap_xy: n_points (%d) exceeds H*W (%d)", n, HW);
  const size_t part_bytes = (size_t)B * S * R * L * n * sizeof(Cand);
  const size_t sel_bytes = (size_t)nBL * n * sizeof(Cand);
  void* scratch;
  rc = fld_scratch(h, part_bytes + sel_bytes, &scratch);
  if (rc) return rc;
  Cand* partial = (Cand*)scratch;
  Cand* sel = (Cand*)((char*)scratch + part_bytes);
  if (n <= 4) topn_partial_kernel<4><<<grid, threads, 0, st>>>(hm, HW, L, R, S, n, partial);
  else if (n <= 16) topn_partial_kernel<16><<<grid, threads, 0, st>>>(hm, HW, L, R, S, n, partial);
  else topn_partial_kernel<FLD_MAX_TOPN><<<grid, threads, 0, st>>>(hm, HW, L, R, S, n, partial);
  FLD_LAUNCHED();
  topn_merge_kernel<<<fld_div_up(nBL, 128), 128, 0, st>>>(partial, W, L, S * R, n, B, (float)thresh, xy, sel);
  FLD_LAUNCHED();
  return FLD_OK;
}
