// simt_extra.cu — CUDA-core layers of the MobileNet-v1 / ResNet50 encoders (HBM-bound, no GEMM):
//   depthwise 3x3 conv + folded BN + ReLU6      reference networks/mobilenet.py:37-47
//   MaxPooling2D(k, stride) 'valid'             reference networks/resnet50.py:149
//   residual add (+ReLU), crop-add              reference networks/resnet50.py:68-69,117-118; networks/fcn.py:55-86,112,119
// One thread per output element with the channel index fastest: all accesses are coalesced over channels.
#include "ops.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float actf(float v, int act) {
  if (act == FLD_ACT_RELU) return fmaxf(v, 0.f);
  if (act == FLD_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

template <typename TIn, typename TOut>
__global__ void dwconv_kernel(const TIn* __restrict__ in, const float* __restrict__ w /*[kh*kw][C]*/, const float* __restrict__ bias,
                              TOut* __restrict__ out, int B, int IH, int IW, int C, int OH, int OW, int kh, int kw, int stride,
                              int pad_t, int pad_l, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float acc = bias ? bias[c] : 0.f;
  for (int a = 0; a < kh; ++a) {
    const int iy = y * stride + a - pad_t;
    if (iy < 0 || iy >= IH) continue;
    for (int d = 0; d < kw; ++d) {
      const int ix = x * stride + d - pad_l;
      if (ix < 0 || ix >= IW) continue;
      acc = fmaf(ldv<TIn>(in + (((size_t)b * IH + iy) * IW + ix) * C + c), __ldg(w + (size_t)(a * kw + d) * C + c), acc);
    }
  }
  stv<TOut>(out + i, actf(acc, act));
}

template <typename TIn, typename TOut>
__global__ void maxpool2d_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int B, int IH, int IW, int C, int OH, int OW, int k,
                                 int s) {
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
      if (iy < IH && ix < IW) m = fmaxf(m, ldv<TIn>(in + (((size_t)b * IH + iy) * IW + ix) * C + c));
    }
  stv<TOut>(out + i, m);
}

template <typename TA, typename TB, typename TOut>
__global__ void add_act_kernel(const TA* __restrict__ a, int AH, int AW, const TB* __restrict__ b2, int BH, int BW, TOut* __restrict__ out,
                               int B, int OH, int OW, int C, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  const float v = ldv<TA>(a + (((size_t)b * AH + y) * AW + x) * C + c) + ldv<TB>(b2 + (((size_t)b * BH + y) * BW + x) * C + c);
  stv<TOut>(out + i, actf(v, act));
}


// ---- bf16 x 8 fast paths (C % 8 == 0): one thread = one pixel x 8 channels, 16-byte accesses
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

__global__ void dwconv_bf16x8_kernel(const uint4* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                     uint4* __restrict__ out, int B, int IH, int IW, int C8, int OH, int OW, int kh, int kw, int stride,
                                     int pad_t, int pad_l, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * OH * OW * C8) return;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  const int C = C8 * 8;
  float acc[8];
  if (bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  }
  for (int a = 0; a < kh; ++a) {
    const int iy = y * stride + a - pad_t;
    if (iy < 0 || iy >= IH) continue;
    for (int d = 0; d < kw; ++d) {
      const int ix = x * stride + d - pad_l;
      if (ix < 0 || ix >= IW) continue;
      float v[8];
      unpack8(__ldg(in + (((size_t)b * IH + iy) * IW + ix) * C8 + c8), v);
      const float* wp = w + (size_t)(a * kw + d) * C + c8 * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]); acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
      acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]); acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = actf(acc[e], act);
  out[i] = pack8(acc);
}

__global__ void maxpool_bf16x8_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int IH, int IW, int C8, int OH, int OW,
                                      int k, int s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * OH * OW * C8) return;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
  for (int a = 0; a < k; ++a)
    for (int d = 0; d < k; ++d) {
      const int iy = y * s + a, ix = x * s + d;
      if (iy < IH && ix < IW) {
        float v[8];
        unpack8(__ldg(in + (((size_t)b * IH + iy) * IW + ix) * C8 + c8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    }
  out[i] = pack8(m);   // max of bf16 values is a bf16 value: exact
}

__global__ void add_bf16x8_kernel(const uint4* __restrict__ a, int AH, int AW, const uint4* __restrict__ b2, int BH, int BW,
                                  uint4* __restrict__ out, int B, int OH, int OW, int C8, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * OH * OW * C8) return;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int x = (int)(r % OW); r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float u[8], v[8];
  unpack8(__ldg(a + (((size_t)b * AH + y) * AW + x) * C8 + c8), u);
  unpack8(__ldg(b2 + (((size_t)b * BH + y) * BW + x) * C8 + c8), v);
#pragma unroll
  for (int e = 0; e < 8; ++e) u[e] = actf(u[e] + v[e], act);
  out[i] = pack8(u);
}

// grid.x is limited to 2^31 - 1: a larger request returns 0 blocks, which the launch rejects loudly (FLD_LAUNCHED) instead of
// silently truncating the element count
unsigned blocks_for(long long total) { const long long b = (total + 255) / 256; return b < (1ll << 31) ? (unsigned)b : 0u; }

}  // namespace

int simt_dwconv(const void* in, int in_dtype, const float* w, const float* bias, void* out, int out_dtype, int B, int IH, int IW, int C,
                int OH, int OW, int kh, int kw, int stride, int pad_t, int pad_l, int act, cudaStream_t st) {
  const long long total = (long long)B * OH * OW * C;
  if (total == 0) return FLD_OK;
  if (in_dtype == FLD_BF16 && out_dtype == FLD_BF16 && C % 8 == 0) {
    dwconv_bf16x8_kernel<<<blocks_for(total / 8), 256, 0, st>>>((const uint4*)in, w, bias, (uint4*)out, B, IH, IW, C / 8, OH, OW, kh, kw, stride,
                                                               pad_t, pad_l, act);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  if (out_dtype == FLD_F32) {
    float* o = (float*)out;
    if (in_dtype == FLD_F32) dwconv_kernel<float, float><<<blocks_for(total), 256, 0, st>>>((const float*)in, w, bias, o, B, IH, IW, C, OH, OW, kh, kw, stride, pad_t, pad_l, act);
    else if (in_dtype == FLD_BF16) dwconv_kernel<__nv_bfloat16, float><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)in, w, bias, o, B, IH, IW, C, OH, OW, kh, kw, stride, pad_t, pad_l, act);
    else { fld_set_error("simt_dwconv: unsupported input dtype"); return FLD_ERR_INVALID; }
  } else {
    __nv_bfloat16* o = (__nv_bfloat16*)out;
    if (in_dtype == FLD_F32) dwconv_kernel<float, __nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const float*)in, w, bias, o, B, IH, IW, C, OH, OW, kh, kw, stride, pad_t, pad_l, act);
    else if (in_dtype == FLD_BF16) dwconv_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)in, w, bias, o, B, IH, IW, C, OH, OW, kh, kw, stride, pad_t, pad_l, act);
    else { fld_set_error("simt_dwconv: unsupported input dtype"); return FLD_ERR_INVALID; }
  }
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_maxpool2d(const void* in, int in_dtype, void* out, int out_dtype, int B, int IH, int IW, int C, int OH, int OW, int k, int s,
                   cudaStream_t st) {
  const long long total = (long long)B * OH * OW * C;
  if (total == 0) return FLD_OK;
  if (in_dtype == FLD_BF16 && out_dtype == FLD_BF16 && C % 8 == 0) {
    maxpool_bf16x8_kernel<<<blocks_for(total / 8), 256, 0, st>>>((const uint4*)in, (uint4*)out, B, IH, IW, C / 8, OH, OW, k, s);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  if (out_dtype == FLD_F32) {
    float* o = (float*)out;
    if (in_dtype == FLD_F32) maxpool2d_kernel<float, float><<<blocks_for(total), 256, 0, st>>>((const float*)in, o, B, IH, IW, C, OH, OW, k, s);
    else if (in_dtype == FLD_BF16) maxpool2d_kernel<__nv_bfloat16, float><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)in, o, B, IH, IW, C, OH, OW, k, s);
    else { fld_set_error("simt_maxpool2d: unsupported input dtype"); return FLD_ERR_INVALID; }
  } else {
    __nv_bfloat16* o = (__nv_bfloat16*)out;
    if (in_dtype == FLD_F32) maxpool2d_kernel<float, __nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const float*)in, o, B, IH, IW, C, OH, OW, k, s);
    else if (in_dtype == FLD_BF16) maxpool2d_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)in, o, B, IH, IW, C, OH, OW, k, s);
    else { fld_set_error("simt_maxpool2d: unsupported input dtype"); return FLD_ERR_INVALID; }
  }
  FLD_LAUNCHED();
  return FLD_OK;
}

template <typename TA, typename TB>
static int add_out(const void* a, int AH, int AW, const void* b, int BH, int BW, void* out, int out_dtype, int B, int OH, int OW, int C,
                   int act, cudaStream_t st) {
  const long long total = (long long)B * OH * OW * C;
  if (out_dtype == FLD_F32)
    add_act_kernel<TA, TB, float><<<blocks_for(total), 256, 0, st>>>((const TA*)a, AH, AW, (const TB*)b, BH, BW, (float*)out, B, OH, OW, C, act);
  else
    add_act_kernel<TA, TB, __nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const TA*)a, AH, AW, (const TB*)b, BH, BW, (__nv_bfloat16*)out, B, OH, OW, C, act);
  FLD_LAUNCHED();
  return FLD_OK;
}

int simt_add_act(const void* a, int a_dtype, int AH, int AW, const void* b, int b_dtype, int BH, int BW, void* out, int out_dtype, int B,
                 int OH, int OW, int C, int act, cudaStream_t st) {
  if ((long long)B * OH * OW * C == 0) return FLD_OK;
  if (a_dtype == FLD_BF16 && b_dtype == FLD_BF16 && out_dtype == FLD_BF16 && C % 8 == 0) {
    add_bf16x8_kernel<<<blocks_for((long long)B * OH * OW * C / 8), 256, 0, st>>>((const uint4*)a, AH, AW, (const uint4*)b, BH, BW, (uint4*)out, B,
                                                                               OH, OW, C / 8, act);
    FLD_LAUNCHED();
    return FLD_OK;
  }
  if (a_dtype == FLD_F32 && b_dtype == FLD_F32) return add_out<float, float>(a, AH, AW, b, BH, BW, out, out_dtype, B, OH, OW, C, act, st);
  if (a_dtype == FLD_F32 && b_dtype == FLD_BF16) return add_out<float, __nv_bfloat16>(a, AH, AW, b, BH, BW, out, out_dtype, B, OH, OW, C, act, st);
  if (a_dtype == FLD_BF16 && b_dtype == FLD_F32) return add_out<__nv_bfloat16, float>(a, AH, AW, b, BH, BW, out, out_dtype, B, OH, OW, C, act, st);
  if (a_dtype == FLD_BF16 && b_dtype == FLD_BF16) return add_out<__nv_bfloat16, __nv_bfloat16>(a, AH, AW, b, BH, BW, out, out_dtype, B, OH, OW, C, act, st);
  fld_set_error("simt_add_act: unsupported dtypes %d, %d", a_dtype, b_dtype);
  return FLD_ERR_INVALID;
}
