// align.cu — batched Umeyama similarity fit fused with a cv2.warpAffine-exact bilinear warp.
//
// Build-defined stage: nothing in the reference computes this; semantics
// are "fp64 closed-form Umeyama, then cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of OpenCV
// 4.13", reproduced bit-exactly with OpenCV's fixed-point scheme (1/32 px sub-pixel, 15-bit weights).
//
// One CTA per face.  Thread 0 fits (sequential fp64 sums, no FMA contraction, so M is bit-identical
// to the CPU oracle), the CTA builds the per-column / per-row fixed-point coordinate tables in shared
// memory, then every thread produces 4 consecutive output pixels (12 bytes -> three 32-bit stores)
// from 8-byte-aligned source loads.  Algorithmic traffic ~37.6 kB written + source footprint read per face.
// MEASURED (ncu, round 1, config C4): DRAM runs at 17 % while the L1 tag stage is the busiest unit — every warp-level
// load touches ~21 sectors on many lines (rotated faces put each lane on its own source row).  Two leaner-ALU rewrites
// (IDP.4A blends, 32-bit funnel loads, loads hoisted for MLP) were SLOWER (0.34 / 0.39 ms vs 0.32 ms) because they issue
// more, narrower requests; the next step is a 2-D warp->pixel mapping with shared-memory staged stores.
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

__device__ void fit_similarity(const float* __restrict__ marks, int N, const double* __restrict__ tmpl, int Nt,
                               int five_point, double* M) {
  // source points
  double px[5], py[5];
  const int n = five_point ? 5 : N;
  double mpx = 0, mpy = 0, mqx = 0, mqy = 0;
  if (five_point) {
    double ax = 0, ay = 0, bx = 0, by = 0;
    for (int i = 36; i < 42; ++i) { ax = da(ax, (double)marks[2 * i]); ay = da(ay, (double)marks[2 * i + 1]); }
    for (int i = 42; i < 48; ++i) { bx = da(bx, (double)marks[2 * i]); by = da(by, (double)marks[2 * i + 1]); }
    px[0] = ax / 6.0; py[0] = ay / 6.0;
    px[1] = bx / 6.0; py[1] = by / 6.0;
    px[2] = marks[60]; py[2] = marks[61];
    px[3] = marks[96]; py[3] = marks[97];
    px[4] = marks[108]; py[4] = marks[109];
    for (int i = 0; i < 5; ++i) {
      mpx = da(mpx, px[i]); mpy = da(mpy, py[i]);
      mqx = da(mqx, tmpl[2 * i]); mqy = da(mqy, tmpl[2 * i + 1]);
    }
  } else {
    for (int i = 0; i < n; ++i) {
      mpx = da(mpx, (double)marks[2 * i]); mpy = da(mpy, (double)marks[2 * i + 1]);
      mqx = da(mqx, tmpl[2 * i]); mqy = da(mqy, tmpl[2 * i + 1]);
    }
  }
  const double dn = (double)n;
  mpx = mpx / dn; mpy = mpy / dn; mqx = mqx / dn; mqy = mqy / dn;
  double var = 0, a = 0, b = 0, c = 0, d = 0;
  for (int i = 0; i < n; ++i) {
    const double sx = five_point ? px[i] : (double)marks[2 * i];
    const double sy = five_point ? py[i] : (double)marks[2 * i + 1];
    const double x = da(sx, -mpx), y = da(sy, -mpy);
    const double qx = da(tmpl[2 * i], -mqx), qy = da(tmpl[2 * i + 1], -mqy);
    var = da(var, da(dm(x, x), dm(y, y)));
    a = da(a, dm(qx, x)); b = da(b, dm(qx, y)); c = da(c, dm(qy, x)); d = da(d, dm(qy, y));
  }
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

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

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

  if (tid == 0) {
    if (M_in) {
      for (int i = 0; i < 6; ++i) fit.M[i] = M_in[(size_t)face * 6 + i];
    } else {
      fit_similarity(marks + (size_t)face * N * 2, N, tmpl, Nt, five_point, fit.M);
    }
    int ok = 1;
    for (int i = 0; i < 6; ++i) ok &= isfinite(fit.M[i]) ? 1 : 0;
    const int fr = face2frame[face];
    if (fr < 0 || fr >= F) ok = 0;
    if (ok) invert_affine(fit.M, fit.iM);
    fit.ok = ok;
    if (M_out && blockIdx.y == 0) for (int i = 0; i < 6; ++i) M_out[(size_t)face * 6 + i] = fit.M[i];
  }
  __syncthreads();
  uint8_t* crop = crops + (size_t)face * out_h * out_w * nc;
  if (!fit.ok) {
    for (int i = y_beg * out_w * nc + tid; i < y_end * out_w * nc; i += kAlignThreads) crop[i] = 0;
    return;
  }
  const double AB = 1024.0;
  for (int x = tid; x < out_w; x += kAlignThreads) {
    adelta[x] = (int)__double2ll_rn(dm(dm(fit.iM[0], (double)x), AB));
    bdelta[x] = (int)__double2ll_rn(dm(dm(fit.iM[3], (double)x), AB));
  }
  for (int y = y_beg + tid; y < y_end; y += kAlignThreads) {
    X0[y - y_beg] = (int)__double2ll_rn(dm(da(dm(fit.iM[1], (double)y), fit.iM[2]), AB)) + 16;
    Y0[y - y_beg] = (int)__double2ll_rn(dm(da(dm(fit.iM[4], (double)y), fit.iM[5]), AB)) + 16;
  }
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

int launch_align(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                 const float* marks, int N, const double* tmpl, int Nt, int five_point, const double* M_in, int B,
                 int out_h, int out_w, double* M_out, uint8_t* crops, cudaStream_t st) {
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
  if (B == 0) return FLD_OK;
  // small batches: split each face over several CTAs (row blocks) so that the grid covers the machine
  int rb = 1;
  while (rb < 8 && (long long)B * rb < 4ll * h->sm_count && out_h / (rb * 2) >= 8) rb *= 2;
  const int rows_per = fld_div_up(out_h, rb);
  const dim3 grid(B, fld_div_up(out_h, rows_per));
  const size_t smem = (size_t)(2 * out_w + 2 * rows_per) * sizeof(int);
  const bool fast = (C == 3) && (out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(crops) & 3) == 0) &&
                    ((size_t)F * H * W * 3 >= 32);
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
                      crops, (cudaStream_t)stream);
}

extern "C" int fld_warp_affine(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                               const double* M, int B, int out_h, int out_w, uint8_t* crops, fld_stream stream) {
  if (!M) { fld_set_error("fld_warp_affine: null M"); return FLD_ERR_INVALID; }
  return launch_align(h, frames, F, H, W, C, face2frame, nullptr, 0, nullptr, 0, 0, M, B, out_h, out_w, nullptr, crops,
                      (cudaStream_t)stream);
}
