// simt_ops.cu — fp32 CUDA-core layer kernels (FLD_F32 parity mode; also the non-GEMM layers of the
// FLD_BF16 mode: transposed convs, crop-add, softmax, dense).
//
// Restates with BN folded (net.cu) the Keras layers used by reference networks/fcn.py:10-51,89-150 and
// networks/utils.py:28-30.  All tensors NHWC.
#include "ops.cuh"

namespace {

// grid.x is limited to 2^31 - 1: a larger request yields 0 blocks, which the launch rejects loudly (FLD_LAUNCHED)
unsigned blocks256(long long total) { const long long b = (total + 255) / 256; return b < (1ll << 31) ? (unsigned)b : 0u; }

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
// SPLIT tensors (FLD_BF16X3 mode, see tc_common.cuh): per pixel [hi(C) | lo(C)] bf16.  As an output type of the SIMT conv it is
// addressed in bf16 elements with pixel pitch 2 * Cout; st_split writes both halves.
struct SplitBf16 { __nv_bfloat16 v; };
__device__ __forceinline__ void st_split(__nv_bfloat16* p, int cout, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  p[0] = h;
  p[cout] = __float2bfloat16_rn(v - __bfloat162float(h));
}
template <typename T> struct OutTraits { static constexpr bool split = false; typedef T elem; };
template <> struct OutTraits<SplitBf16> { static constexpr bool split = true; typedef __nv_bfloat16 elem; };
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == FLD_ACT_RELU) return fmaxf(v, 0.f);
  if (act == FLD_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution: one CTA = 8x8 output pixels of one image x 64 output channels.
// K = kh*kw*Cin walked in chunks of 16; 256 threads, 4x4 register tile each.
// Optional fused MaxPool2x2 through shared memory (pool partners live inside the 8x8 patch).
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16, PT = 8;

template <typename TIn, typename TOut, bool POOL>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const TIn* __restrict__ in, const float* __restrict__ w /*[K][Cout]*/, const float* __restrict__ bias,
                 TOut* __restrict__ out_, ConvGeom g) {
  typedef typename OutTraits<TOut>::elem TO;
  constexpr bool SPLIT = OutTraits<TOut>::split;
  TO* __restrict__ out = reinterpret_cast<TO*>(out_);
  const int opitch = SPLIT ? 2 * g.Cout : g.Cout;   // elements per output pixel
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ float Cs[POOL ? BM : 1][POOL ? BN + 1 : 1];
  const int tiles_x = (g.OW + PT - 1) / PT, tiles_y = (g.OH + PT - 1) / PT;
  int bid = blockIdx.x;
  const int tx_ = bid % tiles_x; bid /= tiles_x;
  const int ty_ = bid % tiles_y; bid /= tiles_y;
  const int b = bid;
  const int n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  const int K = g.kh * g.kw * g.Cin;

  // A-load mapping: k_local = t % 16, pixels t/16 + 16*j
  const int ak = t & 15;
  int a_oy[4], a_ox[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = (t >> 4) + 16 * j;
    a_oy[j] = ty_ * PT + (m >> 3);
    a_ox[j] = tx_ * PT + (m & 7);
  }
  const int bn = t & 63, bk = t >> 6;  // B-load mapping

  const int cty = t >> 4, ctx = t & 15;  // compute mapping: rows cty*4.., cols ctx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const TIn* in_b = in + (size_t)b * g.IH * g.IW * g.Cin;
  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int k = k0 + ak;
      int c = 0, ky = 0, kx = 0;
      const bool kin = k < K;
      if (kin) { const int tap = k / g.Cin; c = k - tap * g.Cin; ky = tap / g.kw; kx = tap - ky * g.kw; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int iy = a_oy[j] * g.stride + ky - g.pad_t;
        const int ix = a_ox[j] * g.stride + kx - g.pad_l;
        float v = 0.f;
        if (kin && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) v = ld_f<TIn>(in_b + ((size_t)iy * g.IW + ix) * g.Cin + c);
        As[ak][(t >> 4) + 16 * j] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + bk + 4 * j;
      const int n = n0 + bn;
      Bs[bk + 4 * j][bn] = (k < K && n < g.Cout) ? __ldg(w + (size_t)k * g.Cout + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][cty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][ctx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ctx * 4 + j;
    const float bsv = (bias && n < g.Cout) ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][j] = apply_act(acc[i][j] + bsv, g.act);
  }
  if (!POOL) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = cty * 4 + i;
      const int oy = ty_ * PT + (m >> 3), ox = tx_ * PT + (m & 7);
      if (oy >= g.OH || ox >= g.OW) continue;
      TO* o = out + (((size_t)b * g.OH + oy) * g.OW + ox) * opitch;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + ctx * 4 + j;
        if (n < g.Cout) {
          if constexpr (SPLIT) st_split(o + n, g.Cout, acc[i][j]);
          else st_f<TO>(o + n, acc[i][j]);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Cs[cty * 4 + i][ctx * 4 + j] = acc[i][j];
    __syncthreads();
    const int PH = g.OH >> 1, PW = g.OW >> 1;  // MaxPooling2D 'valid': floor
    for (int e = t; e < 16 * BN; e += 256) {
      const int n = e & 63, pp = e >> 6;
      const int py = pp >> 2, px = pp & 3;
      const int oy = ty_ * (PT / 2) + py, ox = tx_ * (PT / 2) + px;
      if (oy >= PH || ox >= PW || n0 + n >= g.Cout) continue;
      const int m = (py * 2) * 8 + px * 2;
      const float v = fmaxf(fmaxf(Cs[m][n], Cs[m + 1][n]), fmaxf(Cs[m + 8][n], Cs[m + 9][n]));
      TO* o = out + (((size_t)b * PH + oy) * PW + ox) * opitch + n0 + n;
      if constexpr (SPLIT) st_split(o, g.Cout, v);
      else st_f<TO>(o, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Conv2DTranspose 'valid', no bias (fcn.py:104,114,121,145), gather form.
// weights repacked to [kh][kw][Cin][Cout] so consecutive threads (cout) read consecutive floats.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void deconv_simt_kernel(const TIn* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, int B, int IH,
                                   int IW, int Cin, int OH, int OW, int Cout, int k, int s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * Cout;
  if (i >= total) return;
  const int o = (int)(i % Cout);
  long long r = i / Cout;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float acc = 0.f;
  for (int a = y % s; a < k; a += s) {
    const int iy = (y - a) / s;
    if (iy < 0 || iy >= IH) continue;
    for (int c2 = x % s; c2 < k; c2 += s) {
      const int ix = (x - c2) / s;
      if (ix < 0 || ix >= IW) continue;
      const TIn* ip = in + (((size_t)b * IH + iy) * IW + ix) * Cin;
      const float* wp = w + ((size_t)(a * k + c2) * Cin) * Cout + o;
      for (int c = 0; c < Cin; ++c) acc = fmaf(ld_f<TIn>(ip + c), __ldg(wp + (size_t)c * Cout), acc);
    }
  }
  out[i] = acc;
}

// fcn.py:55-86 crop + Add: both operands cropped (bottom/right) to the common size
__global__ void add_crop_kernel(const float* __restrict__ a, int AH, int AW, const float* __restrict__ b2, int BH, int BW,
                                float* __restrict__ out, int B, int OH, int OW, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  out[i] = a[(((size_t)b * AH + y) * AW + x) * C + c] + b2[(((size_t)b * BH + y) * BW + x) * C + c];
}

// C % 4 == 0: 16-byte accesses (the FCN skip adds work on 68-channel fp32 score maps)
__global__ void add_crop_vec4_kernel(const float4* __restrict__ a, int AH, int AW, const float4* __restrict__ b2, int BH, int BW,
                                     float4* __restrict__ out, int B, int OH, int OW, int C4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * OH * OW * C4) return;
  const int c = (int)(i % C4);
  long long r = i / C4;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  const float4 u = __ldg(a + (((size_t)b * AH + y) * AW + x) * C4 + c), v = __ldg(b2 + (((size_t)b * BH + y) * BW + x) * C4 + c);
  out[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
}

// networks/utils.py:28-30: softmax over the channel axis, one warp per pixel
__global__ void softmax_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_px, int C) {
  const long long px = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (px >= n_px) return;
  const float* p = in + px * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, p[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(p[c] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float* q = out + px * C;
  for (int c = lane; c < C; c += 32) q[c] = expf(p[c] - mx) / sum;
}

// Flatten + Dense, split-K: out[b][j] = sum_k in[b][k] * W[k][j] + bias[j].
// pass 1: one CTA = 8 items x 128 outputs x one K slice -> partial sums; pass 2: fixed-order reduction over the
// slices + bias + activation (deterministic; no atomics).  K = 4096 in one serial loop was latency-bound.
constexpr int DF = 8, DK = 128, DSLICE = 256;
template <typename TIn>
__global__ void __launch_bounds__(128)
dense_partial_kernel(const TIn* __restrict__ in, const float* __restrict__ w, float* __restrict__ part, int B, int In, int Out, int csplit) {
  __shared__ float xs[DF][DK];
  const int b0 = blockIdx.x * DF;
  const int j = blockIdx.y * 128 + threadIdx.x;
  const int kbeg = blockIdx.z * DSLICE, kend = min(In, kbeg + DSLICE);
  float acc[DF];
#pragma unroll
  for (int f = 0; f < DF; ++f) acc[f] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += DK) {
    for (int e = threadIdx.x; e < DF * DK; e += 128) {
      const int f = e / DK, k = e - f * DK;
      float xv = 0.f;
      if (b0 + f < B && k0 + k < kend) {
        if (csplit) {   // SPLIT input: logical element kk = (pixel, c) lives at pixel * 2C + c (hi) and + C (lo)
          const int kk = k0 + k, px = kk / csplit, c = kk - px * csplit;
          const TIn* q = in + (size_t)(b0 + f) * In * 2 + (size_t)px * 2 * csplit + c;
          xv = ld_f<TIn>(q) + ld_f<TIn>(q + csplit);
        } else {
          xv = ld_f<TIn>(in + (size_t)(b0 + f) * In + k0 + k);
        }
      }
      xs[f][k] = xv;
    }
    __syncthreads();
    if (j < Out) {
      const int kmax = min(DK, kend - k0);
      const float* wp = w + (size_t)k0 * Out + j;
      int k = 0;
      for (; k + 8 <= kmax; k += 8) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = __ldg(wp + (size_t)(k + u) * Out);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int f = 0; f < DF; ++f) acc[f] = fmaf(xs[f][k + u], wv[u], acc[f]);
      }
      for (; k < kmax; ++k) {
        const float wv = __ldg(wp + (size_t)k * Out);
#pragma unroll
        for (int f = 0; f < DF; ++f) acc[f] = fmaf(xs[f][k], wv, acc[f]);
      }
    }
    __syncthreads();
  }
  if (j < Out) {
#pragma unroll
    for (int f = 0; f < DF; ++f)
      if (b0 + f < B) part[((size_t)blockIdx.z * B + b0 + f) * Out + j] = acc[f];
  }
}

__global__ void dense_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ out, int B,
                                    int Out, int KS, int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Out) return;
  float s = 0.f;
  for (int z = 0; z < KS; ++z) s += part[(size_t)z * B * Out + i];
  out[i] = apply_act(s + (bias ? bias[i % Out] : 0.f), act);
}

__global__ void maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int IH, int IW, int C, int OH, int OW,
                               int k, int s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float m = -INFINITY;
  for (int a = 0; a < k; ++a)
    for (int d = 0; d < k; ++d) {
      const int iy = y * s + a, ix = x * s + d;
      if (iy < IH && ix < IW) m = fmaxf(m, in[(((size_t)b * IH + iy) * IW + ix) * C + c]);
    }
  out[i] = m;
}

__global__ void cvt_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

template <typename TIn, typename TOut>
int launch_conv_t(const void* in, const float* w, const float* bias, void* out, const ConvGeom& g, int B, cudaStream_t st) {
  const int tiles = B * fld_div_up(g.OH, PT) * fld_div_up(g.OW, PT);
  dim3 grid(tiles, fld_div_up(g.Cout, BN));
  if (g.pool)
    conv_simt_kernel<TIn, TOut, true><<<grid, 256, 0, st>>>((const TIn*)in, w, bias, (TOut*)out, g);
  else
    conv_simt_kernel<TIn, TOut, false><<<grid, 256, 0, st>>>((const TIn*)in, w, bias, (TOut*)out, g);
  FLD_LAUNCHED();
  return FLD_OK;
}

}  // namespace

int simt_conv(const void* in, int in_dtype, const float* w, const float* bias, void* out, int out_dtype, const ConvGeom& g, int B,
              cudaStream_t st) {
  if (B == 0) return FLD_OK;
  if (g.pool && g.pool != 2) { fld_set_error("simt_conv: only 2x2 fused pooling"); return FLD_ERR_INVALID; }
  if (out_dtype == FLD_F32) {
    if (in_dtype == FLD_U8) return launch_conv_t<uint8_t, float>(in, w, bias, out, g, B, st);
    if (in_dtype == FLD_F32) return launch_conv_t<float, float>(in, w, bias, out, g, B, st);
    if (in_dtype == FLD_BF16) return launch_conv_t<__nv_bfloat16, float>(in, w, bias, out, g, B, st);
  } else if (out_dtype == FLD_BF16) {
    if (in_dtype == FLD_U8) return launch_conv_t<uint8_t, __nv_bfloat16>(in, w, bias, out, g, B, st);
    if (in_dtype == FLD_F32) return launch_conv_t<float, __nv_bfloat16>(in, w, bias, out, g, B, st);
    if (in_dtype == FLD_BF16) return launch_conv_t<__nv_bfloat16, __nv_bfloat16>(in, w, bias, out, g, B, st);
  } else if (out_dtype == FLD_BF16X3) {   // SPLIT output feeding a FLD_BF16X3 tensor-core conv
    if (in_dtype == FLD_U8) return launch_conv_t<uint8_t, SplitBf16>(in, w, bias, out, g, B, st);
    if (in_dtype == FLD_F32) return launch_conv_t<float, SplitBf16>(in, w, bias, out, g, B, st);
  }
  fld_set_error("simt_conv: unsupported dtype combination %d -> %d", in_dtype, out_dtype);
  return FLD_ERR_INVALID;
}

int simt_deconv(const void* in, int in_dtype, const float* w, float* out, int B, int IH, int IW, int Cin, int OH, int OW, int Cout,
                int k, int s, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  const long long total = (long long)B * OH * OW * Cout;
  const long long blocks = (total + 127) / 128;
  if (blocks >= (1ll << 31)) { fld_set_error("simt_deconv: tensor too large"); return FLD_ERR_INVALID; }
  if (in_dtype == FLD_F32)
    deconv_simt_kernel<float><<<(unsigned)blocks, 128, 0, st>>>((const float*)in, w, out, B, IH, IW, Cin, OH, OW, Cout, k, s);
  else if (in_dtype == FLD_BF16)
    deconv_simt_kernel<__nv_bfloat16><<<(unsigned)blocks, 128, 0, st>>>((const __nv_bfloat16*)in, w, out, B, IH, IW, Cin, OH, OW,
                                                                       Cout, k, s);
  else { fld_set_error("simt_deconv: unsupported input dtype"); return FLD_ERR_INVALID; }
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_add_crop(const float* a, int AH, int AW, const float* b, int BH, int BW, float* out, int B, int OH, int OW, int C,
                  cudaStream_t st) {
  if (B == 0) return FLD_OK;
  const long long total = (long long)B * OH * OW * C;
  if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    add_crop_vec4_kernel<<<blocks256(total / 4), 256, 0, st>>>((const float4*)a, AH, AW, (const float4*)b, BH, BW, (float4*)out, B, OH,
                                                                            OW, C / 4);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  add_crop_kernel<<<blocks256(total), 256, 0, st>>>(a, AH, AW, b, BH, BW, out, B, OH, OW, C);
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_softmax(const float* in, float* out, long long n_px, int C, cudaStream_t st) {
  if (n_px == 0) return FLD_OK;
  const long long blocks = (n_px * 32 + 255) / 256;
  if (blocks >= (1ll << 31)) { fld_set_error("simt_softmax: tensor too large"); return FLD_ERR_INVALID; }
  softmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, n_px, C);
  FLD_LAUNCHED();
  return FLD_OK;
}

size_t simt_dense_scratch_bytes(int B, int In, int Out) { return (size_t)fld_div_up(In, DSLICE) * B * Out * sizeof(float); }

int simt_dense(const void* in, int in_dtype, const float* w, const float* bias, float* out, float* scratch, int B, int In, int Out,
               int act, cudaStream_t st, int in_channels) {
  if (B == 0) return FLD_OK;
  const int KS = fld_div_up(In, DSLICE);
  dim3 grid(fld_div_up(B, DF), fld_div_up(Out, 128), KS);
  if (in_dtype == FLD_F32) dense_partial_kernel<float><<<grid, 128, 0, st>>>((const float*)in, w, scratch, B, In, Out, 0);
  else if (in_dtype == FLD_BF16) dense_partial_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)in, w, scratch, B, In, Out, 0);
  else if (in_dtype == FLD_BF16X3 && in_channels > 0)   // SPLIT tensor: hi + lo summed on load
    dense_partial_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)in, w, scratch, B, In, Out, in_channels);
  else { fld_set_error("simt_dense: unsupported input dtype"); return FLD_ERR_INVALID; }
  FLD_LAUNCHED();
  dense_reduce_kernel<<<fld_div_up(B * Out, 256), 256, 0, st>>>(scratch, bias, out, B, Out, KS, act);
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_dense_reduce(const float* part, const float* bias, float* out, int B, int Out, int KS, int act, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  dense_reduce_kernel<<<fld_div_up(B * Out, 256), 256, 0, st>>>(part, bias, out, B, Out, KS, act);
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_maxpool(const float* in, float* out, int B, int IH, int IW, int C, int OH, int OW, int k, int s, cudaStream_t st) {
  if (B == 0) return FLD_OK;
  const long long total = (long long)B * OH * OW * C;
  maxpool_kernel<<<blocks256(total), 256, 0, st>>>(in, out, B, IH, IW, C, OH, OW, k, s);
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_cvt_bf16_f32(const void* in, float* out, long long n, cudaStream_t st) {
  if (n == 0) return FLD_OK;
  cvt_bf16_f32_kernel<<<blocks256(n), 256, 0, st>>>((const __nv_bfloat16*)in, out, n);
  FLD_LAUNCHED();
  return FLD_OK;
}

