// preprocess.cu — batched face crop + cv2.resize-exact bilinear resize (+BGR->RGB), and
// get_image_array (resize + normalise to float32).
//
// Replaces reference prediction.py:36-83 (box math, crop, cv2.resize, cvtColor) and
// data/generator.py:50-69.  The resize reproduces OpenCV 4.13's uint8 INTER_LINEAR bit-exactly:
// 11-bit fixed-point coefficients computed from float32 fractions, horizontal then vertical pass with
// OpenCV's intermediate truncations, x-border fraction clamp / y-border index clamp, the exact-2x
// INTER_AREA shortcut and the same-size copy shortcut (oracle/preprocess.py states the scheme).
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
// output rows per CTA: the x coefficient table (fp64 arithmetic per entry) and the box math are per-CTA work, so large batches take
// more rows per CTA; small ones keep 16 so that the grid still covers the machine (measured on 4096 faces: 16 / 32 / 64 / 128 rows -> 0.419 / 0.386 / 0.374 / 0.374 ms)
static int rows_per_cta(long long items, int rows, int sm_count) {
  int rpc = 16;
  { const char* e = getenv("FLD_RESIZE_RPC"); if (e && atoi(e) >= 4) return atoi(e); }
  while (rpc < 64 && rpc * 2 <= rows && items * ((rows + rpc * 2 - 1) / (rpc * 2)) >= 4ll * sm_count) rpc *= 2;
  return rpc;
}

// fp64 without FMA contraction: OpenCV evaluates (d+0.5)*scale-0.5 with separate mul / sub
__device__ __forceinline__ void axis_coeff(int d, double scale, int src, bool clamp_frac, int* ofs, int* a0, int* a1) {
  const float f = __double2float_rn(__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5));
  int s = (int)floorf(f);
  float fr = __fsub_rn(f, (float)s);
  if (clamp_frac) {
    if (s < 0) { s = 0; fr = 0.f; }
    if (s >= src - 1) { s = src - 1; fr = 0.f; }
  }
  *ofs = s;
  *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, fr), 2048.f));
  *a1 = __float2int_rn(__fmul_rn(fr, 2048.f));
}

struct Rect { int x0, y0, w, h; };

// six consecutive bytes starting at an arbitrary address, from one or two aligned 8-byte loads (the buffer is 8-byte aligned and
// the caller keeps the 16-byte window inside it)
__device__ __forceinline__ uint64_t load6(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  const unsigned sh = (unsigned)(a & 7) * 8;
  const uint64_t lo = __ldg(q);
  if (sh <= 16) return lo >> sh;
  const uint64_t hi = __ldg(q + 1);
  return (lo >> sh) | (hi << (64 - sh));
}

// prediction.py:67-78 + :36-65 with Python integer semantics
__device__ void square_box(const int32_t* face, int* fb) {
  int x0 = face[0], y0 = face[1], x1 = face[2], y1 = face[3];
  // int(abs((y1 - y0) * 0.1)): double product truncated toward zero
  const int off = (int)fabs(__dmul_rn((double)(y1 - y0), 0.1));
  y0 += off; y1 += off;
  const int bw = x1 - x0, bh = y1 - y0;
  const int diff = bh - bw;
  const int delta = (diff < 0 ? -diff : diff) / 2;  // int(abs(diff)/2)
  const bool odd = (diff & 1) != 0;                  // Python: diff % 2 == 1 (true for odd negatives too)
  if (diff > 0) { x0 -= delta; x1 += delta; if (odd) x1 += 1; }
  else if (diff < 0) { y0 -= delta; y1 += delta; if (odd) y1 += 1; }
  fb[0] = x0; fb[1] = y0; fb[2] = x1; fb[3] = y1;
}

// MODE 0: faces -> uint8 (optional R/B swap).  MODE 1: whole image -> float32 normalised.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
resize_kernel(const uint8_t* __restrict__ src, int F, int H, int W, const int32_t* __restrict__ boxes,
              const int32_t* __restrict__ face2frame, int dw, int dh, int swap_rb, int norm, uint8_t* __restrict__ out_u8,
              float* __restrict__ out_f32, int32_t* __restrict__ faceboxes, uint4* __restrict__ staged, int n_items, int vec_ok, int rpc) {
  extern __shared__ int sm[];
  int* xofs = sm;            // [dw]
  int* xa0 = sm + dw;        // [dw]
  int* xa1 = sm + 2 * dw;    // [dw]
  int* yt = sm + 3 * dw;     // [rpc][4]: y0, y1, b0, b1
  __shared__ Rect rc;
  __shared__ int s_frame;
  const int item = blockIdx.x;
  const int row0 = blockIdx.y * rpc;
  const int tid = threadIdx.x;

  if (tid == 0) {
    if (MODE == 0) {
      int fb[4];
      square_box(boxes + 4 * item, fb);
      if (blockIdx.y == 0 && faceboxes) { for (int i = 0; i < 4; ++i) faceboxes[4 * item + i] = fb[i]; }
      const int cx0 = max(fb[0], 0), cy0 = max(fb[1], 0);
      const int cx1 = min(fb[2], W), cy1 = min(fb[3], H);
      rc.x0 = cx0; rc.y0 = cy0; rc.w = cx1 - cx0; rc.h = cy1 - cy0;
      int fr = face2frame ? face2frame[item] : 0;
      s_frame = (fr >= 0 && fr < F) ? fr : -1;
    } else {
      rc.x0 = 0; rc.y0 = 0; rc.w = W; rc.h = H;
      s_frame = item;
    }
  }
  __syncthreads();
  const int cw = rc.w, ch = rc.h;
  const int nrows = min(rpc, dh - row0);
  const bool empty = (cw <= 0 || ch <= 0 || s_frame < 0);
  const bool area2 = (cw == 2 * dw && ch == 2 * dh);
  const bool same = (cw == dw && ch == dh);
  if (!empty && !area2 && !same) {
    const double sx = 1.0 / ((double)dw / (double)cw);
    const double sy = 1.0 / ((double)dh / (double)ch);
    for (int x = tid; x < dw; x += kThreads) axis_coeff(x, sx, cw, true, &xofs[x], &xa0[x], &xa1[x]);
    for (int r = tid; r < nrows; r += kThreads) {
      int s, b0, b1;
      axis_coeff(row0 + r, sy, ch, false, &s, &b0, &b1);
      yt[4 * r + 0] = min(max(s, 0), ch - 1);
      yt[4 * r + 1] = min(max(s + 1, 0), ch - 1);
      yt[4 * r + 2] = b0;
      yt[4 * r + 3] = b1;
    }
  }
  __syncthreads();
  const uint8_t* base = empty ? src : src + ((size_t)s_frame * H + rc.y0) * (size_t)W * 3 + (size_t)rc.x0 * 3;
  const size_t rs = (size_t)W * 3;
  const int n = nrows * dw;
  // Vector path (the general bilinear case, dw % 4 == 0, aligned buffers): a thread produces FOUR consecutive output pixels.  The two
  // source pixels of a tap pair are six consecutive bytes, fetched with one or two aligned 8-byte loads instead of six byte loads,
  // and the 12 output bytes (or 12 floats) leave as three 4-byte (16-byte) stores.  The byte-per-instruction version issued 15
  // memory instructions per pixel and was bound by L1 requests (0.37 of HBM on 4096 faces).  Same integer arithmetic, same bits.
  if (vec_ok && !empty && !same && !area2) {
    const uint8_t* safe_end = src + (size_t)F * H * W * 3 - 16;     // last address an aligned 16-byte window may start at
    const int gpr = dw >> 2;
    const int ng = nrows * gpr;
    for (int g = tid; g < ng; g += kThreads) {
      const int r = g / gpr, x4 = (g - r * gpr) << 2;
      const int y = row0 + r;
      const uint8_t* p0 = base + (size_t)yt[4 * r + 0] * rs;
      const uint8_t* p1 = base + (size_t)yt[4 * r + 1] * rs;
      const int b0 = yt[4 * r + 2], b1 = yt[4 * r + 3];
      int v[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xo = xofs[x4 + j], a0 = xa0[x4 + j], a1 = xa1[x4 + j];
        const uint8_t* q0 = p0 + xo * 3;
        const uint8_t* q1 = p1 + xo * 3;
        uint64_t w0, w1;
        if (xo + 1 < cw && q1 <= safe_end && q0 <= safe_end) {
          w0 = load6(q0); w1 = load6(q1);
        } else {                             // right border (the second tap has weight 0 there) or the buffer's last bytes: byte loads
          const int x1 = min(xo + 1, cw - 1) - xo;
          w0 = 0; w1 = 0;
#pragma unroll
          for (int c = 2; c >= 0; --c) {
            w0 = (w0 << 8) | q0[x1 * 3 + c]; w1 = (w1 << 8) | q1[x1 * 3 + c];
          }
#pragma unroll
          for (int c = 2; c >= 0; --c) {
            w0 = (w0 << 8) | q0[c]; w1 = (w1 << 8) | q1[c];
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int t0 = (int)((w0 >> (8 * c)) & 255) * a0 + (int)((w0 >> (24 + 8 * c)) & 255) * a1;
          const int t1 = (int)((w1 >> (8 * c)) & 255) * a0 + (int)((w1 >> (24 + 8 * c)) & 255) * a1;
          const int o = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
          v[j][c] = min(max(o, 0), 255);
        }
      }
      const size_t o = ((size_t)item * dh + y) * dw + x4;
      if (MODE == 0) {
        uint32_t by[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          by[3 * j + 0] = (uint32_t)(swap_rb ? v[j][2] : v[j][0]);
          by[3 * j + 1] = (uint32_t)v[j][1];
          by[3 * j + 2] = (uint32_t)(swap_rb ? v[j][0] : v[j][2]);
        }
        uint32_t* q = reinterpret_cast<uint32_t*>(out_u8 + o * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = by[4 * k] | (by[4 * k + 1] << 8) | (by[4 * k + 2] << 16) | (by[4 * k + 3] << 24);
        if (staged) {
          const int H2 = dh >> 1, W2 = dw >> 1;
          const size_t plane = (size_t)n_items * H2 * W2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = x4 + j;
            __nv_bfloat162 lo = __floats2bfloat162_rn((float)by[3 * j], (float)by[3 * j + 1]), hi = __floats2bfloat162_rn((float)by[3 * j + 2], 0.f);
            staged[(size_t)((y & 1) * 2 + (x & 1)) * plane + ((size_t)item * H2 + (y >> 1)) * W2 + (x >> 1)] =
                make_uint4(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), 0u, 0u);
          }
        }
      } else {
        float f[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (norm == 0) {  // sub_mean then reverse channels (generator.py:53-61)
            f[3 * j + 2] = __fsub_rn((float)v[j][0], 103.939f);
            f[3 * j + 1] = __fsub_rn((float)v[j][1], 116.779f);
            f[3 * j + 0] = __fsub_rn((float)v[j][2], 123.68f);
          } else if (norm == 1) {  // generator.py:51
#pragma unroll
            for (int c = 0; c < 3; ++c) f[3 * j + c] = __fsub_rn(__fdiv_rn((float)v[j][c], 127.5f), 1.0f);
          } else {  // generator.py:63-65
#pragma unroll
            for (int c = 0; c < 3; ++c) f[3 * j + c] = __fdiv_rn((float)v[j][c], 255.0f);
          }
        }
        float4* q = reinterpret_cast<float4*>(out_f32 + o * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
      }
    }
    return;
  }
  for (int i = tid; i < n; i += kThreads) {
    const int r = i / dw, x = i - r * dw;
    const int y = row0 + r;
    int v[3];
    if (empty) {
      v[0] = v[1] = v[2] = 0;
    } else if (same) {
      const uint8_t* p = base + (size_t)y * rs + (size_t)x * 3;
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    } else if (area2) {
      const uint8_t* p = base + (size_t)(2 * y) * rs + (size_t)(2 * x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (p[c] + p[3 + c] + p[rs + c] + p[rs + 3 + c] + 2) >> 2;
    } else {
      const int xo = xofs[x], a0 = xa0[x], a1 = xa1[x];
      const int x1 = min(xo + 1, cw - 1);
      const uint8_t* p0 = base + (size_t)yt[4 * r + 0] * rs;
      const uint8_t* p1 = base + (size_t)yt[4 * r + 1] * rs;
      const int b0 = yt[4 * r + 2], b1 = yt[4 * r + 3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int t0 = p0[xo * 3 + c] * a0 + p0[x1 * 3 + c] * a1;
        const int t1 = p1[xo * 3 + c] * a0 + p1[x1 * 3 + c] * a1;
        int o = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
        v[c] = min(max(o, 0), 255);
      }
    }
    const size_t o = ((size_t)item * dh + y) * dw + x;
    if (MODE == 0) {
      uint8_t* q = out_u8 + o * 3;
      const int c0 = swap_rb ? v[2] : v[0], c2 = swap_rb ? v[0] : v[2];
      q[0] = (uint8_t)c0; q[1] = (uint8_t)v[1]; q[2] = (uint8_t)c2;
      if (staged) {
        // the first conv layer's operand staging (csrc/tc_conv_s2d.cu): the same pixel widened to 8 bf16 channels in the
        // space-to-depth plane (y & 1, x & 1) — written here so that the network's widening pass is skipped
        const int H2 = dh >> 1, W2 = dw >> 1;
        const size_t plane = (size_t)n_items * H2 * W2;
        __nv_bfloat162 lo = __floats2bfloat162_rn((float)c0, (float)v[1]), hi = __floats2bfloat162_rn((float)c2, 0.f);
        staged[(size_t)((y & 1) * 2 + (x & 1)) * plane + ((size_t)item * H2 + (y >> 1)) * W2 + (x >> 1)] =
            make_uint4(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), 0u, 0u);
      }
    } else {
      float* q = out_f32 + o * 3;
      if (norm == 0) {  // sub_mean then reverse channels (generator.py:53-61)
        q[2] = __fsub_rn((float)v[0], 103.939f);
        q[1] = __fsub_rn((float)v[1], 116.779f);
        q[0] = __fsub_rn((float)v[2], 123.68f);
      } else if (norm == 1) {  // generator.py:51
#pragma unroll
        for (int c = 0; c < 3; ++c) q[c] = __fsub_rn(__fdiv_rn((float)v[c], 127.5f), 1.0f);
      } else {  // generator.py:63-65
#pragma unroll
        for (int c = 0; c < 3; ++c) q[c] = __fdiv_rn((float)v[c], 255.0f);
      }
    }
  }
}

}  // namespace

static int preprocess_faces(fld_handle* h, const uint8_t* frames, int F, int H, int W, const int32_t* boxes, const int32_t* face2frame,
                            int B, int S, int swap_rb, uint8_t* out, int32_t* faceboxes, void* staging, fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(frames && boxes && out, "fld_preprocess_faces: null pointer");
  FLD_REQUIRE(F > 0 && H > 0 && W > 0 && S > 0 && S <= 4096 && B >= 0, "fld_preprocess_faces: bad shape");
  FLD_REQUIRE(!staging || (S % 4 == 0 && (reinterpret_cast<uintptr_t>(staging) & 15) == 0),
              "fld_preprocess_faces_staged: the staging layout needs S % 4 == 0 and a 16-byte aligned buffer");
  const int rpc = rows_per_cta(B, S, h->sm_count);
  const size_t smem = (size_t)(3 * S + 4 * rpc) * sizeof(int);
  dim3 grid(B, fld_div_up(S, rpc));
  const int vec_ok = (S % 4 == 0 && (reinterpret_cast<uintptr_t>(frames) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 &&
                      (size_t)F * H * W * 3 >= 32 && !getenv("FLD_RESIZE_SCALAR")) ? 1 : 0;
  resize_kernel<0><<<grid, kThreads, smem, (cudaStream_t)stream>>>(frames, F, H, W, boxes, face2frame, S, S, swap_rb, 0, out,
                                                                   nullptr, faceboxes, (uint4*)staging, B, vec_ok, rpc);
  FLD_LAUNCHED();
  return FLD_OK;
}

extern "C" int fld_preprocess_faces(fld_handle* h, const uint8_t* frames, int F, int H, int W, const int32_t* boxes,
                                    const int32_t* face2frame, int B, int S, int swap_rb, uint8_t* out, int32_t* faceboxes,
                                    fld_stream stream) {
  return preprocess_faces(h, frames, F, H, W, boxes, face2frame, B, S, swap_rb, out, faceboxes, nullptr, stream);
}

extern "C" int fld_preprocess_faces_staged(fld_handle* h, const uint8_t* frames, int F, int H, int W, const int32_t* boxes,
                                           const int32_t* face2frame, int B, int S, int swap_rb, uint8_t* out, int32_t* faceboxes,
                                           void* staging, fld_stream stream) {
  if (!staging) { fld_set_error("fld_preprocess_faces_staged: null staging (fld_net_input_staging)"); return FLD_ERR_INVALID; }
  return preprocess_faces(h, frames, F, H, W, boxes, face2frame, B, S, swap_rb, out, faceboxes, staging, stream);
}

extern "C" int fld_image_array(fld_handle* h, const uint8_t* images, int B, int H, int W, int ow, int oh, int norm, float* out,
                               fld_stream stream) {
  int rc = fld_enter(h);
  if (rc) return rc;
  if (B == 0) return FLD_OK;
  FLD_REQUIRE(images && out, "fld_image_array: null pointer");
  FLD_REQUIRE(H > 0 && W > 0 && ow > 0 && oh > 0 && ow <= 8192 && B >= 0, "fld_image_array: bad shape");
  FLD_REQUIRE(norm >= 0 && norm <= 2, "fld_image_array: norm must be 0 (sub_mean), 1 (sub_and_divide) or 2 (divide)");
  if (B == 0) return FLD_OK;
  const int rpc = rows_per_cta(B, oh, h->sm_count);
  const size_t smem = (size_t)(3 * ow + 4 * rpc) * sizeof(int);
  dim3 grid(B, fld_div_up(oh, rpc));
  const int vec_ok = (ow % 4 == 0 && (reinterpret_cast<uintptr_t>(images) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (size_t)B * H * W * 3 >= 32 && !getenv("FLD_RESIZE_SCALAR")) ? 1 : 0;
  resize_kernel<1><<<grid, kThreads, smem, (cudaStream_t)stream>>>(images, B, H, W, nullptr, nullptr, ow, oh, 0, norm, nullptr,
                                                                   out, nullptr, nullptr, B, vec_ok, rpc);
  FLD_LAUNCHED();
  return FLD_OK;
}
