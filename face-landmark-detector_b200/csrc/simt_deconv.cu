// simt_deconv.cu — Conv2DTranspose (padding='valid', no bias) with kernel = 2 x stride as stride^2 PHASE convolutions.
//
// Replaces the Keras layers at reference networks/fcn.py:104,114 (4x4 s2), :121 (16x16 s8) and :145 (64x64 s32).
// With k = 2s every output pixel (y, x) = (s*i + a, s*j + b), 0 <= a, b < s, receives exactly 2 x 2 taps:
//     out[s*i+a, s*j+b, o] = sum_{u,v in {0,1}} sum_c in[i+u-1, j+v-1, c] * W[a + s(1-u), b + s(1-v), o, c]
// i.e. phase (a, b) is a 2x2 convolution over the input zero-padded by one pixel, written depth-to-space.  Each
// phase is an implicit GEMM [pixels] x [K = 4*Cin] x [Cout]: same 8x8-pixel register-tiled scheme as
// conv_simt_kernel, blockIdx.z = phase, weights repacked per phase on the host (net.cu) to [phase][u][v][Cin][Cout].
// The naive one-thread-per-output gather kernel this replaces spent 21 ms on fcn_8's up8 at batch 32.
#include "ops.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

constexpr int BM = 64, BK = 16, PT = 8;

// TN = output channels per thread (4 -> 64-wide N tile, 5 -> 80-wide: Cout = 68 fits one tile)
template <typename TIn, int TN>
__global__ void __launch_bounds__(256)
deconv_phase_kernel(const TIn* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, int IH, int IW, int Cin,
                    int Cout, int s) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 1];
  const int PHh = IH + 1, PWw = IW + 1;           // phase-image size
  const int tiles_x = (PWw + PT - 1) / PT, tiles_y = (PHh + PT - 1) / PT;
  int bid = blockIdx.x;
  const int tx_ = bid % tiles_x; bid /= tiles_x;
  const int ty_ = bid % tiles_y; bid /= tiles_y;
  const int b = bid;
  const int n0 = blockIdx.y * BN;
  const int phase = blockIdx.z, pa = phase / s, pb = phase - pa * s;
  const int t = threadIdx.x;
  const int K = 4 * Cin;
  const float* wp = w + (size_t)phase * K * Cout;

  const int ak = t & 15;
  int a_i[4], a_j[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int m = (t >> 4) + 16 * q;
    a_i[q] = ty_ * PT + (m >> 3);
    a_j[q] = tx_ * PT + (m & 7);
  }
  const int cty = t >> 4, ctx = t & 15;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const TIn* in_b = in + (size_t)b * IH * IW * Cin;
  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int k = k0 + ak;
      const bool kin = k < K;
      int c = 0, u = 0, v = 0;
      if (kin) { const int tap = k / Cin; c = k - tap * Cin; u = tap >> 1; v = tap & 1; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int iy = a_i[q] + u - 1, ix = a_j[q] + v - 1;
        float val = 0.f;
        if (kin && iy >= 0 && iy < IH && ix >= 0 && ix < IW) val = ldf<TIn>(in_b + ((size_t)iy * IW + ix) * Cin + c);
        As[ak][(t >> 4) + 16 * q] = val;
      }
    }
    for (int e = t; e < BK * BN; e += 256) {
      const int kk = e / BN, n = e - kk * BN;
      Bs[kk][n] = (k0 + kk < K && n0 + n < Cout) ? __ldg(wp + (size_t)(k0 + kk) * Cout + n0 + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][cty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][ctx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int OH = s * PHh, OW = s * PWw;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = cty * 4 + i;
    const int pi = ty_ * PT + (m >> 3), pj = tx_ * PT + (m & 7);
    if (pi >= PHh || pj >= PWw) continue;
    float* o = out + (((size_t)b * OH + (s * pi + pa)) * OW + (s * pj + pb)) * Cout;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + ctx * TN + j;
      if (n < Cout) o[n] = acc[i][j];
    }
  }
}

template <typename TIn>
int launch(const void* in, const float* w, float* out, int B, int IH, int IW, int Cin, int Cout, int s, cudaStream_t st) {
  const int tiles = B * fld_div_up(IH + 1, PT) * fld_div_up(IW + 1, PT);
  const int rem = Cout % 64;
  const bool wide = rem > 0 && rem <= 16;  // e.g. 68 -> one 80-wide tile instead of 64 + 4
  if (s * s > 65535) { fld_set_error("simt_deconv_phase: stride too large"); return FLD_ERR_INVALID; }
  if (wide) {
    dim3 grid(tiles, fld_div_up(Cout, 80), s * s);
    deconv_phase_kernel<TIn, 5><<<grid, 256, 0, st>>>((const TIn*)in, w, out, IH, IW, Cin, Cout, s);
  } else {
    dim3 grid(tiles, fld_div_up(Cout, 64), s * s);
    deconv_phase_kernel<TIn, 4><<<grid, 256, 0, st>>>((const TIn*)in, w, out, IH, IW, Cin, Cout, s);
  }
  FLD_LAUNCHED();
  return FLD_OK;
}

}  // namespace

// weights: [s*s phases][2][2][Cin][Cout] (see header); k must equal 2*s
int simt_deconv_phase(const void* in, int in_dtype, const float* w_phase, float* out, int B, int IH, int IW, int Cin, int Cout, int s,
                      cudaStream_t st) {
  if (B == 0) return FLD_OK;
  if (in_dtype == FLD_F32) return launch<float>(in, w_phase, out, B, IH, IW, Cin, Cout, s, st);
  if (in_dtype == FLD_BF16) return launch<__nv_bfloat16>(in, w_phase, out, B, IH, IW, Cin, Cout, s, st);
  fld_set_error("simt_deconv_phase: unsupported input dtype");
  return FLD_ERR_INVALID;
}
