// ops.cuh — internal layer-op entry points shared by net.cu / simt_ops.cu / tc_conv.cu
#pragma once
#include "common.cuh"
#include <vector>

struct ConvGeom {
  int IH, IW, Cin;      // input (un-padded)
  int OH, OW, Cout;     // conv output BEFORE the optional fused pool
  int kh, kw, stride;
  int pad_t, pad_l;     // zero padding on top / left (bottom / right implied by OH, OW)
  int act;              // fld_act
  int pool;             // 0 or 2
};

// ---- fp32 CUDA-core kernels (simt_ops.cu)
int simt_conv(const void* in, int in_dtype, const float* w /*[kh*kw*Cin][Cout]*/, const float* bias, void* out, int out_dtype,
              const ConvGeom& g, int B, cudaStream_t st);
int simt_deconv(const void* in, int in_dtype, const float* w /*[k][k][Cin][Cout]*/, float* out, int B, int IH, int IW, int Cin,
                int OH, int OW, int Cout, int k, int s, cudaStream_t st);
// k == 2*s: stride^2 phase convolutions (simt_deconv.cu); weights [s*s][2][2][Cin][Cout]
int simt_deconv_phase(const void* in, int in_dtype, const float* w_phase, float* out, int B, int IH, int IW, int Cin, int Cout, int s,
                      cudaStream_t st);
int simt_add_crop(const float* a, int AH, int AW, const float* b, int BH, int BW, float* out, int B, int OH, int OW, int C,
                  cudaStream_t st);
int simt_softmax(const float* in, float* out, long long n_px, int C, cudaStream_t st);
size_t simt_dense_scratch_bytes(int B, int In, int Out);
// in_dtype FLD_BF16X3: `in` is a SPLIT tensor ([hi | lo] bf16 per pixel of in_channels channels); In counts logical elements
int simt_dense(const void* in, int in_dtype, const float* w /*[In][Out]*/, const float* bias, float* out, float* scratch /*split-K partials*/,
               int B, int In, int Out, int act, cudaStream_t st, int in_channels = 0);
// fixed-order sum of split-K partials [KS][B][Out] + bias + activation (deterministic: no atomics)
int simt_dense_reduce(const float* part, const float* bias, float* out, int B, int Out, int KS, int act, cudaStream_t st);
int simt_maxpool(const float* in, float* out, int B, int IH, int IW, int C, int OH, int OW, int k, int s, cudaStream_t st);
int simt_cvt_bf16_f32(const void* in, float* out, long long n, cudaStream_t st);

// ---- tcgen05 tensor-core kernels (tc_conv.cu)
struct TcConvPlan;  // opaque: tensor maps + launch geometry for one (layer, batch, buffers) combination
// first layer: 3x3, Cin = 3, pad 1, stride 1, fused bias + ReLU (+pool2); input u8 or f32 NHWC, output bf16 NHWC
// x3: FLD_BF16X3 variant (uint8 input only): weights split hi / lo (K = 80), SPLIT output
int tc_conv_first(const fld_handle* h, const void* in, int in_dtype, const __nv_bfloat16* w_packed /*[Cout][32]*/,
                  const float* bias, __nv_bfloat16* out, const ConvGeom& g, int B, cudaStream_t st, int x3 = 0);
bool tc_conv_first_supported(const ConvGeom& g);
void tc_conv_first_pack(const float* w_host /*[27][Cout]*/, const float* bias_host /*[Cout] or null*/, int Cout,
                        uint16_t (*f2bf)(float), uint16_t* out /*[Cout*8*kg]*/, int kg = 6);
// first layer with the 2x2 pool window in the TMEM columns (tc_conv_s2d.cu): space-to-depth planes of the widened image in
// `scratch`, four TMA boxes per tile, bias through a K group; pool = 2 only; bf16 output, or SPLIT output with x3 (uint8 input)
struct TcS2dPlan;
bool tc_conv_s2d_supported(const ConvGeom& g);
size_t tc_conv_s2d_scratch_bytes(const ConvGeom& g, int B);
void tc_conv_s2d_pack(const float* w_host /*[27][Cout]*/, const float* bias_host, int Cout, uint16_t (*f2bf)(float), uint16_t* out /*[Cout*8*(x3 ? 80 : 40)]*/, int x3);
int tc_conv_s2d_plan_create(const fld_handle* h, void* scratch, int in_dtype, const ConvGeom& g, int B, int x3, int split_out, TcS2dPlan** out);
void tc_conv_s2d_plan_destroy(TcS2dPlan* p);
// staged != 0: the planes in `scratch` were already written by the producer of `in` (fld_preprocess_faces_staged): no widening pass
int tc_conv_s2d_run(const TcS2dPlan* p, const void* in, const __nv_bfloat16* w_packed, void* out, cudaStream_t st, int staged = 0);
// strided stems (tc_conv_stem.cu): k x k (3 or 7), stride 2, Cin = 3, any zero padding, bias folded, no pool; bf16 NHWC out
bool tc_conv_stem_supported(const ConvGeom& g);
int tc_conv_stem_kgroups(int ks);
void tc_conv_stem_pack(const float* w_host /*[k*k*3][Cout]*/, const float* bias_host, int ks, int Cout, uint16_t (*f2bf)(float),
                       uint16_t* out /*[kgroups*8][Cout] core-matrix packed*/);
int tc_conv_stem(const fld_handle* h, const void* in, int in_dtype, const __nv_bfloat16* w_packed, __nv_bfloat16* out, const ConvGeom& g,
                 int B, cudaStream_t st);
// generic: stride 1, Cin % 64 == 0, bf16 NHWC in; weights bf16 [kh*kw][Cout_pad][Cin]; out bf16 or f32 NHWC
bool tc_conv_supported(const ConvGeom& g);
// x3: FLD_BF16X3 operands (activations [x_hi | x_lo] 2*Cin channels, weights [w_hi | w_hi | w_lo] 3*Cin); split_out: bf16 output
// stored as a SPLIT tensor; ksplit > 1 (flat 1x1 convs, fp32 output): slice s writes its partial sums to out + s * npx * Cout
int tc_conv_plan_create(const fld_handle* h, const void* in, const __nv_bfloat16* w_packed, int cout_pad, const ConvGeom& g, int B,
                        TcConvPlan** out, int x3 = 0, int split_out = 0, int ksplit = 1);
void tc_conv_plan_destroy(TcConvPlan* p);
int tc_conv_run(const TcConvPlan* p, const float* bias, void* out, int out_dtype, cudaStream_t st);

// 3x3 halo-reuse / stationary-weight variant (tc_conv_halo.cu); bf16 in, bf16 out
struct TcHaloPlan;
bool tc_halo_supported(const ConvGeom& g, int cout_pad);
int tc_halo_plan_create(const fld_handle* h, const void* in, const __nv_bfloat16* w_packed, int cout_pad, const ConvGeom& g, int B,
                        TcHaloPlan** out, int x3 = 0, int split_out = 0);
void tc_halo_plan_destroy(TcHaloPlan* p);
int tc_halo_run(const TcHaloPlan* p, const float* bias, void* out, cudaStream_t st);

// ---- encoder extras (simt_extra.cu): depthwise conv, k/s max-pool, add(+act); f32 or bf16 NHWC on either side
int simt_dwconv(const void* in, int in_dtype, const float* w /*[kh*kw][C]*/, const float* bias, void* out, int out_dtype, int B, int IH,
                int IW, int C, int OH, int OW, int kh, int kw, int stride, int pad_t, int pad_l, int act, cudaStream_t st);
int simt_maxpool2d(const void* in, int in_dtype, void* out, int out_dtype, int B, int IH, int IW, int C, int OH, int OW, int k, int s,
                   cudaStream_t st);
int simt_add_act(const void* a, int a_dtype, int AH, int AW, const void* b, int b_dtype, int BH, int BW, void* out, int out_dtype, int B,
                 int OH, int OW, int C, int act, cudaStream_t st);

// ---- transposed conv (k = 2*stride) as one flat GEMM on the tensor cores with fused softmax / argmax (tc_deconv.cu)
struct TcDeconvPlan;
bool tc_deconv_supported(int k, int s, int Cin, int Cout);
int tc_deconv_kp(int Cin);    // packed K: 4 taps x Cin, padded to 64
int tc_deconv_cpp(int Cout);  // channels per phase, padded to 8
size_t tc_deconv_scratch_bytes(int B, int IH, int IW, int Cin);  // bf16 im2col matrix
void tc_deconv_pack_weights(const float* w_phase /*[s*s][2][2][Cin][Cout]*/, int s, int Cin, int Cout, uint16_t (*f2bf)(float),
                            std::vector<uint16_t>& out /*[s*s*cpp][Kp]*/);
int tc_deconv_plan_create(const fld_handle* h, void* scratch, const __nv_bfloat16* w_packed, int B, int IH, int IW, int Cin, int Cout,
                          int s, TcDeconvPlan** out, int x3 = 0);
// FLD_BF16X3 variant (class maps only, tc_deconv_run mode 2): split operands, three product terms in one accumulation
bool tc_deconv_x3_supported(int k, int s, int Cin, int Cout);
size_t tc_deconv_x3_scratch_bytes(int B, int IH, int IW, int Cin);
void tc_deconv_x3_pack_weights(const float* w_phase, int s, int Cin, int Cout, uint16_t (*f2bf)(float), float (*bf2f)(uint16_t),
                               std::vector<uint16_t>& out);
void tc_deconv_x3_pack_weights_hilo(const float* w_phase, int s, int Cin, int Cout, uint16_t (*f2bf)(float), float (*bf2f)(uint16_t),
                                    std::vector<uint16_t>& out);   // plan variants 2 (this) / 3 (tc_deconv_pack_weights)
void tc_deconv_plan_destroy(TcDeconvPlan* p);
// mode 0: fp32 logits, 1: softmax probabilities, 2: int64 argmax class map,
// 3: fused soft centroid of the softmax output -> out = double [B][Cout][2] (x, y), acc = fp32 scratch of tc_deconv_acc_bytes
size_t tc_deconv_acc_bytes(int B, int Cout);
int tc_deconv_run(const TcDeconvPlan* p, const float* in, void* out, int mode, cudaStream_t st, float* acc = nullptr, double thresh = 0.0,
                  const float* addend = nullptr);
