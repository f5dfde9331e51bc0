// net.cu — layer-graph executor behind fld_net_*: shape inference, BN folding, weight repacking for
// the kernel layouts, workspace planning and per-layer dispatch (tcgen05 kernels in FLD_BF16 mode, fp32
// CUDA-core kernels in FLD_F32 mode and for the non-GEMM layers).
//
// Replaces model.signatures["predict"] (reference prediction.py:84) and model.predict (prediction.py:208)
// for graphs expressed in the vocabulary of networks/fcn.py / networks/utils.py.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ops.cuh"

namespace {

struct TensorInfo {
  int h = 0, w = 0, c = 0;
  int dtype = FLD_F32;
  size_t elems() const { return (size_t)h * w * c; }
  size_t esz() const { return dtype == FLD_U8 ? 1 : dtype == FLD_BF16 ? 2 : 4; }   // FLD_BF16X3 = SPLIT: 2 x bf16 per element
};

enum Path { PATH_SIMT = 0, PATH_TC_FIRST = 1, PATH_TC_TMA = 2 };

struct PlanEntry { int B; const void* in; TcConvPlan* plan; };
struct HaloPlanEntry { int B; const void* in; TcHaloPlan* plan; };
struct DeconvPlanEntry { int B; const void* scratch; TcDeconvPlan* plan; int kind = 0; };   // kind: tc_deconv_plan_create's x3 variant
struct S2dPlanEntry { int B; const void* scratch; TcS2dPlan* plan; };

struct LayerRt {
  fld_layer_desc d;
  ConvGeom g{};
  int path = PATH_SIMT;
  int cout_pad = 0;
  bool x3 = false;              // FLD_BF16X3 tensor-core conv: SPLIT input, weights packed [w_hi | w_hi | w_lo]
  bool s2d = false;             // first layer through tc_conv_s2d.cu (pool window in the TMEM columns)
  bool dc_fuse_softmax = false; // the following SOFTMAX layer is computed in this layer's epilogue (logits never reach HBM)
  bool skip = false;            // SOFTMAX layer folded into the preceding transposed conv
  bool x3_cmap = false;         // FLD_BF16X3: the last transposed conv has a tensor-core variant for class-map-only forwards (d_wbf)
  bool needs_weights = false, has_weights = false;
  std::vector<float> w_host;  // folded fp32: conv [K][Cout]; deconv [k][k][Cin][Cout]; dense [In][Out]
  std::vector<float> b_host;  // folded bias [Cout] (empty = none)
  float* d_w = nullptr;
  float* d_bias = nullptr;
  __nv_bfloat16* d_wbf = nullptr;
  __nv_bfloat16* d_wbf_hilo = nullptr;  // FLD_BF16X3 two-pass transposed conv: [w_hi | w_lo] (pass 1)
  __nv_bfloat16* d_wbf_hi = nullptr;    //   and plain bf16(w) (pass 2)
  float* d_zero = nullptr;      // all-zero bias for split-K partial sums (tensor-core dense)
  std::vector<PlanEntry> plans;
  std::vector<HaloPlanEntry> hplans;
  std::vector<DeconvPlanEntry> dplans;
  std::vector<S2dPlanEntry> splans;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct fld_net {
  fld_handle* h;
  int compute;
  std::vector<TensorInfo> tensors;  // [0] = input
  std::vector<LayerRt> layers;
  bool finalized = false;
  bool profiling = false;
  bool profiled_once = false;
  int retained = 0;                 // > 0 while captured CUDA graphs reference this net's plans: nothing is evicted
  std::vector<cudaEvent_t> events;  // n_layers + 1, recorded around every layer when profiling
};

namespace {

int infer_shapes(fld_net* net) {
  const int bf = net->compute == FLD_BF16;
  const int nL = (int)net->layers.size();
  for (int i = 0; i < nL; ++i) {
    LayerRt& L = net->layers[i];
    const fld_layer_desc& d = L.d;
    FLD_REQUIRE(d.in0 >= 0 && d.in0 <= i, "layer %d: in0=%d must reference an earlier tensor", i, d.in0);
    const TensorInfo& a = net->tensors[d.in0];
    TensorInfo o;
    switch (d.op) {
      case FLD_OP_CONV: {
        FLD_REQUIRE(d.kh > 0 && d.kw > 0 && d.stride > 0 && d.cout > 0, "layer %d: bad conv parameters", i);
        ConvGeom& g = L.g;
        g.IH = a.h; g.IW = a.w; g.Cin = a.c;
        g.kh = d.kh; g.kw = d.kw; g.stride = d.stride; g.pad_t = d.pad_t; g.pad_l = d.pad_l;
        g.OH = (a.h + d.pad_t + d.pad_b - d.kh) / d.stride + 1;
        g.OW = (a.w + d.pad_l + d.pad_r - d.kw) / d.stride + 1;
        g.Cout = d.cout; g.act = d.act; g.pool = d.pool;
        FLD_REQUIRE(g.OH > 0 && g.OW > 0, "layer %d: conv output is empty", i);
        FLD_REQUIRE(d.pool == 0 || d.pool == 2, "layer %d: pool must be 0 or 2", i);
        o.h = d.pool ? g.OH / 2 : g.OH; o.w = d.pool ? g.OW / 2 : g.OW; o.c = d.cout;
        FLD_REQUIRE(o.h > 0 && o.w > 0, "layer %d: pooled output is empty", i);
        L.needs_weights = true;
        break;
      }
      case FLD_OP_DECONV:
        FLD_REQUIRE(d.kh == d.kw && d.kh >= d.stride && d.stride > 0 && d.cout > 0, "layer %d: bad deconv parameters", i);
        o.h = (a.h - 1) * d.stride + d.kh; o.w = (a.w - 1) * d.stride + d.kw; o.c = d.cout;
        L.needs_weights = true;
        break;
      case FLD_OP_ADD: {
        FLD_REQUIRE(d.in1 >= 0 && d.in1 <= i, "layer %d: in1=%d must reference an earlier tensor", i, d.in1);
        const TensorInfo& b2 = net->tensors[d.in1];
        FLD_REQUIRE(a.c == b2.c, "layer %d: ADD channel mismatch %d vs %d", i, a.c, b2.c);
        o.h = a.h < b2.h ? a.h : b2.h; o.w = a.w < b2.w ? a.w : b2.w; o.c = a.c;
        break;
      }
      case FLD_OP_DENSE:
        FLD_REQUIRE(d.cout > 0, "layer %d: bad dense units", i);
        o.h = 1; o.w = 1; o.c = d.cout;
        L.needs_weights = true;
        break;
      case FLD_OP_SOFTMAX:
        o = a;
        break;
      case FLD_OP_MAXPOOL:
        FLD_REQUIRE(d.kh > 0 && d.stride > 0, "layer %d: bad pool parameters", i);
        o.h = (a.h - d.kh) / d.stride + 1; o.w = (a.w - d.kh) / d.stride + 1; o.c = a.c;
        break;
      case FLD_OP_DWCONV:
        FLD_REQUIRE(d.kh > 0 && d.kw > 0 && d.stride > 0, "layer %d: bad depthwise conv parameters", i);
        FLD_REQUIRE(d.cout == 0 || d.cout == a.c, "layer %d: depthwise conv keeps the channel count (%d), got cout=%d", i, a.c, d.cout);
        L.d.cout = a.c;  // depth_multiplier 1 (mobilenet.py:37)
        o.h = (a.h + d.pad_t + d.pad_b - d.kh) / d.stride + 1; o.w = (a.w + d.pad_l + d.pad_r - d.kw) / d.stride + 1; o.c = a.c;
        FLD_REQUIRE(o.h > 0 && o.w > 0, "layer %d: depthwise conv output is empty", i);
        L.needs_weights = true;
        break;
      default:
        fld_set_error("layer %d: op %d not implemented", i, d.op);
        return FLD_ERR_INVALID;
    }
    o.dtype = FLD_F32;
    net->tensors.push_back(o);
  }
  // dtype / kernel-path assignment for the tensor-core mode
  if (bf) {
    std::vector<int> f32_needed(net->tensors.size(), 0);
    f32_needed.back() = 1;  // final tensor is handed out as fp32
    for (int i = 0; i < nL; ++i) {
      const fld_layer_desc& d = net->layers[i].d;
      // FCN skip adds work on fp32 logits; residual adds (add -> ReLU, resnet50.py:68-69) take the trunk's bf16 tensors
      if (d.op == FLD_OP_ADD && d.act == FLD_ACT_NONE) { f32_needed[d.in0] = 1; f32_needed[d.in1] = 1; }
      // logits feeding transposed convs / softmax stay fp32 (they are tiny; bf16 there only costs accuracy)
      if (d.op == FLD_OP_SOFTMAX || d.op == FLD_OP_DECONV) f32_needed[d.in0] = 1;
    }
    for (int i = 0; i < nL; ++i) {
      LayerRt& L = net->layers[i];
      TensorInfo& o = net->tensors[i + 1];
      const TensorInfo& a = net->tensors[L.d.in0];
      if (L.d.op == FLD_OP_DECONV && L.d.kh == L.d.kw && tc_deconv_supported(L.d.kh, L.d.stride, a.c, L.d.cout) && a.dtype == FLD_F32 &&
          !getenv("FLD_TC_DECONV_OFF")) {
        L.path = PATH_TC_TMA;  // the stride^2 phase convolutions as one tensor-core GEMM (tc_deconv.cu)
      }
      const bool want_f32 = f32_needed[i + 1] != 0;
      if (L.d.op == FLD_OP_DWCONV || L.d.op == FLD_OP_MAXPOOL || L.d.op == FLD_OP_ADD) { o.dtype = want_f32 ? FLD_F32 : FLD_BF16; continue; }
      if (L.d.op == FLD_OP_DENSE && a.dtype == FLD_BF16 && a.elems() % 64 == 0 && L.d.act == FLD_ACT_NONE && !getenv("FLD_TC_DENSE_OFF")) {
        // Flatten + Dense as a flat 1x1 conv over B "pixels" with Cin = H*W*C on the tensor cores, split-K (see net_forward)
        L.path = PATH_TC_TMA;
        L.cout_pad = (int)align_up(L.d.cout, 16);
        L.g = ConvGeom{};
        L.g.IH = L.g.IW = L.g.OH = L.g.OW = 1; L.g.Cin = (int)a.elems(); L.g.Cout = L.d.cout;
        L.g.kh = L.g.kw = 1; L.g.stride = 1; L.g.act = FLD_ACT_NONE; L.g.pool = 0;
        if (L.cout_pad > 256 || !tc_conv_supported(L.g)) L.path = PATH_SIMT;
      }
      if (L.d.op != FLD_OP_CONV) { o.dtype = FLD_F32; continue; }
      const bool in_ok_first = (a.dtype == FLD_U8 || a.dtype == FLD_F32) && (tc_conv_first_supported(L.g) || tc_conv_stem_supported(L.g));
      const bool in_ok_tma = (a.dtype == FLD_BF16) && tc_conv_supported(L.g);
      if (in_ok_first && !want_f32) { L.path = PATH_TC_FIRST; o.dtype = FLD_BF16; }
      else if (in_ok_tma && !(want_f32 && L.g.pool)) { L.path = PATH_TC_TMA; o.dtype = want_f32 ? FLD_F32 : FLD_BF16; }
      else { L.path = PATH_SIMT; o.dtype = want_f32 ? FLD_F32 : FLD_BF16; }
      if (L.path == PATH_TC_TMA) L.cout_pad = (int)align_up(L.g.Cout, L.g.Cout > 256 ? 128 : 16);
    }
    // DECONV (tensor cores) immediately followed by the final SOFTMAX over its channels: fuse the softmax into the epilogue
    for (int i = 0; i + 1 < nL; ++i) {
      LayerRt& L = net->layers[i];
      LayerRt& N2 = net->layers[i + 1];
      if (L.d.op != FLD_OP_DECONV || L.path != PATH_TC_TMA || N2.d.op != FLD_OP_SOFTMAX || N2.d.in0 != i + 1) continue;
      bool other_use = false;
      for (int j = i + 2; j < nL; ++j) other_use |= (net->layers[j].d.in0 == i + 1 || net->layers[j].d.in1 == i + 1);
      if (other_use || getenv("FLD_TC_FUSE_OFF")) continue;
      L.dc_fuse_softmax = true;
      N2.skip = true;
    }
  }
  // fp32-accurate tensor-core mode: SPLIT tensors between tensor-core convs, the fp32 kernels everywhere else
  if (net->compute == FLD_BF16X3) {
    std::vector<std::vector<int>> consumers(net->tensors.size());
    for (int i = 0; i < nL; ++i) {
      consumers[net->layers[i].d.in0].push_back(i);
      if (net->layers[i].d.op == FLD_OP_ADD) consumers[net->layers[i].d.in1].push_back(i);
    }
    // out_split[i]: layer i (a conv) stores a SPLIT tensor — possible when every consumer is a tensor-core-capable conv or a
    // dense layer.  geom_ok[i]: conv i can run on the tensor cores given a SPLIT input (a fused pool needs a SPLIT output).
    // Consumers have larger indices, so decide back to front.
    std::vector<char> out_split(nL, 0), geom_ok(nL, 0);
    for (int i = nL - 1; i >= 0; --i) {
      const LayerRt& L = net->layers[i];
      if (L.d.op == FLD_OP_CONV && L.g.Cout % 8 == 0 && i + 1 != nL && !consumers[i + 1].empty() && !getenv("FLD_X3_OFF")) {
        bool all = true;
        for (int j : consumers[i + 1]) all = all && (geom_ok[j] || net->layers[j].d.op == FLD_OP_DENSE);
        out_split[i] = all;
      }
      geom_ok[i] = L.d.op == FLD_OP_CONV && tc_conv_supported(L.g) && (out_split[i] || !L.g.pool);
    }
    for (int i = 0; i < nL; ++i) {
      LayerRt& L = net->layers[i];
      net->tensors[i + 1].dtype = out_split[i] ? FLD_BF16X3 : FLD_F32;
      if (L.d.op == FLD_OP_CONV && geom_ok[i] && net->tensors[L.d.in0].dtype == FLD_BF16X3) {
        L.path = PATH_TC_TMA;
        L.x3 = true;
        L.cout_pad = (int)align_up(L.g.Cout, L.g.Cout > 256 ? 128 : 16);
      } else if (L.d.op == FLD_OP_CONV && out_split[i] && tc_conv_first_supported(L.g) && !getenv("FLD_X3_FIRST_OFF") &&
                 (net->tensors[L.d.in0].dtype == FLD_U8 || (net->tensors[L.d.in0].dtype == FLD_F32 && L.d.in0 == 0 && !getenv("FLD_X3_FIRST_F32_OFF")))) {
        L.path = PATH_TC_FIRST;   // uint8 is exact in bf16: only the weights are split (tc_conv_first.cu, K = 80); float pixels are split too (K = 128)
        L.x3 = true;
      }
    }
    // the final DECONV -> SOFTMAX pair: a class-map-only forward (fld_net_forward_classmap; argmax is invariant under the softmax)
    // runs the transposed conv on the tensor cores with split operands and takes the argmax in its epilogue; probabilities and
    // landmark decodes keep the fp32 kernels
    if (nL >= 2 && !getenv("FLD_X3_DECONV_OFF")) {
      LayerRt& L = net->layers[nL - 2];
      const LayerRt& S = net->layers[nL - 1];
      const TensorInfo& a = net->tensors[L.d.in0];
      if (L.d.op == FLD_OP_DECONV && S.d.op == FLD_OP_SOFTMAX && S.d.in0 == nL - 1 && L.d.kh == L.d.kw && a.dtype == FLD_F32 &&
          tc_deconv_x3_supported(L.d.kh, L.d.stride, a.c, L.d.cout))
        L.x3_cmap = true;
    }
  }
  return FLD_OK;
}

void free_layer(LayerRt& L) {
  if (L.d_w) cudaFree(L.d_w);
  if (L.d_bias) cudaFree(L.d_bias);
  if (L.d_wbf) cudaFree(L.d_wbf);
  if (L.d_zero) cudaFree(L.d_zero);
  L.d_zero = nullptr;
  if (L.d_wbf_hilo) cudaFree(L.d_wbf_hilo);
  if (L.d_wbf_hi) cudaFree(L.d_wbf_hi);
  L.d_wbf_hilo = nullptr; L.d_wbf_hi = nullptr;
  for (auto& pe : L.plans) tc_conv_plan_destroy(pe.plan);
  for (auto& pe : L.hplans) tc_halo_plan_destroy(pe.plan);
  L.hplans.clear();
  for (auto& pe : L.dplans) tc_deconv_plan_destroy(pe.plan);
  L.dplans.clear();
  for (auto& pe : L.splans) tc_conv_s2d_plan_destroy(pe.plan);
  L.splans.clear();
  L.d_w = nullptr; L.d_bias = nullptr; L.d_wbf = nullptr; L.plans.clear();
}

uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

float bf2f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace

// Bytes tensor t occupies in the workspace for batch B.  The logits of a transposed conv whose softmax / argmax / centroid is
// fused into its epilogue are never written (14.6 MB per 224x224 image for fcn_8): they get no space.
static size_t tensor_ws_bytes(const fld_net* net, int t, int B) {
  if (t >= 1 && net->layers[t - 1].dc_fuse_softmax) return 0;
  return align_up(net->tensors[t].elems() * net->tensors[t].esz() * (size_t)B, 1024);
}

extern "C" int fld_net_create(fld_handle* h, const fld_layer_desc* layers_h, int n_layers, int in_h, int in_w, int in_c, int in_dtype,
                              int compute, fld_net** out) {
  int rc = fld_enter(h);
  if (rc) return rc;
  FLD_REQUIRE(layers_h && out && n_layers > 0, "fld_net_create: null/empty layer list");
  FLD_REQUIRE(in_h > 0 && in_w > 0 && in_c > 0, "fld_net_create: bad input shape");
  FLD_REQUIRE(in_dtype == FLD_U8 || in_dtype == FLD_F32, "fld_net_create: input dtype must be FLD_U8 or FLD_F32");
  FLD_REQUIRE(compute == FLD_F32 || compute == FLD_BF16 || compute == FLD_BF16X3, "fld_net_create: compute must be FLD_F32, FLD_BF16 or FLD_BF16X3");
  fld_net* net = new fld_net();
  net->h = h;
  net->compute = compute;
  TensorInfo in;
  in.h = in_h; in.w = in_w; in.c = in_c; in.dtype = in_dtype;
  net->tensors.push_back(in);
  net->layers.resize(n_layers);
  for (int i = 0; i < n_layers; ++i) net->layers[i].d = layers_h[i];
  rc = infer_shapes(net);
  if (rc) { delete net; return rc; }
  *out = net;
  return FLD_OK;
}

extern "C" void fld_net_destroy(fld_net* net) {
  if (!net) return;
  cudaSetDevice(net->h->device);
  for (auto& L : net->layers) free_layer(L);
  for (auto e : net->events) cudaEventDestroy(e);
  delete net;
}

extern "C" int fld_net_set_weights(fld_net* net, int layer, const float* kernel_h, const float* bias_h, const float* bn_h, float eps) {
  FLD_REQUIRE(net, "fld_net_set_weights: null net");
  FLD_REQUIRE(layer >= 0 && layer < (int)net->layers.size(), "fld_net_set_weights: layer %d out of range", layer);
  LayerRt& L = net->layers[layer];
  FLD_REQUIRE(L.needs_weights, "fld_net_set_weights: layer %d takes no weights", layer);
  FLD_REQUIRE(kernel_h, "fld_net_set_weights: null kernel");
  const fld_layer_desc& d = L.d;
  const TensorInfo& a = net->tensors[d.in0];
  const int Cout = d.cout;
  std::vector<double> scale(Cout, 1.0), shift(Cout, 0.0);
  for (int o = 0; o < Cout; ++o) {
    double b = bias_h ? (double)bias_h[o] : 0.0;
    if (bn_h) {
      const double g = bn_h[o], be = bn_h[Cout + o], mu = bn_h[2 * Cout + o], var = bn_h[3 * Cout + o];
      const double s = g / sqrt(var + (double)eps);
      scale[o] = s;
      shift[o] = (b - mu) * s + be;
    } else {
      shift[o] = b;
    }
  }
  const bool any_bias = bias_h || bn_h;
  if (d.op == FLD_OP_CONV) {
    const size_t K = (size_t)d.kh * d.kw * a.c;
    const double isc = d.in_scale != 0.f ? (double)d.in_scale : 1.0;
    L.w_host.resize(K * Cout);
    for (size_t k = 0; k < K; ++k)
      for (int o = 0; o < Cout; ++o) L.w_host[k * Cout + o] = (float)((double)kernel_h[k * Cout + o] * scale[o] * isc);
  } else if (d.op == FLD_OP_DECONV) {
    FLD_REQUIRE(!bn_h && !bias_h, "fld_net_set_weights: Conv2DTranspose layers carry no bias/BN in the reference graphs");
    const int k = d.kh, Cin = a.c, st = d.stride;
    L.w_host.resize((size_t)k * k * Cin * Cout);
    if (k == 2 * st) {
      // phase layout for simt_deconv_phase: [a*s+b][u][v][Cin][Cout] = W[a + s(1-u)][b + s(1-v)][o][c]
      for (int pa = 0; pa < st; ++pa)
        for (int pb = 0; pb < st; ++pb)
          for (int u = 0; u < 2; ++u)
            for (int v = 0; v < 2; ++v) {
              const int ka = pa + st * (1 - u), kb = pb + st * (1 - v);
              float* dst = &L.w_host[((((size_t)(pa * st + pb) * 2 + u) * 2 + v) * Cin) * Cout];
              const float* src = kernel_h + ((size_t)(ka * k + kb) * Cout) * Cin;
              for (int c = 0; c < Cin; ++c)
                for (int o = 0; o < Cout; ++o) dst[(size_t)c * Cout + o] = src[(size_t)o * Cin + c];
            }
    } else {
      for (int t = 0; t < k * k; ++t)
        for (int o = 0; o < Cout; ++o)
          for (int c = 0; c < Cin; ++c)
            L.w_host[((size_t)t * Cin + c) * Cout + o] = kernel_h[((size_t)t * Cout + o) * Cin + c];
    }
  } else if (d.op == FLD_OP_DWCONV) {
    // Keras depthwise kernel [kh][kw][C][1] -> [tap][C], BN scale folded per channel
    const size_t T = (size_t)d.kh * d.kw;
    L.w_host.resize(T * Cout);
    for (size_t t = 0; t < T; ++t)
      for (int o = 0; o < Cout; ++o) L.w_host[t * Cout + o] = (float)((double)kernel_h[t * Cout + o] * scale[o]);
  } else if (d.op == FLD_OP_DENSE) {
    const size_t In = a.elems();
    L.w_host.resize(In * Cout);
    for (size_t k = 0; k < In; ++k)
      for (int o = 0; o < Cout; ++o) L.w_host[k * Cout + o] = (float)((double)kernel_h[k * Cout + o] * scale[o]);
  }
  L.b_host.clear();
  if (any_bias) {
    L.b_host.resize(Cout);
    for (int o = 0; o < Cout; ++o) L.b_host[o] = (float)shift[o];
  }
  L.has_weights = true;
  net->finalized = false;
  return FLD_OK;
}

extern "C" int fld_net_finalize(fld_net* net) {
  FLD_REQUIRE(net, "fld_net_finalize: null net");
  int rc = fld_enter(net->h);
  if (rc) return rc;
  for (size_t i = 0; i < net->layers.size(); ++i) {
    LayerRt& L = net->layers[i];
    if (!L.needs_weights) continue;
    if (!L.has_weights) { fld_set_error("fld_net_finalize: layer %zu has no weights", i); return FLD_ERR_STATE; }
    free_layer(L);
    L.s2d = false;
    const int Cout = L.d.cout;
    const TensorInfo& a = net->tensors[L.d.in0];
    // bias (padded so the epilogue can always read 32 floats per chunk)
    {
      const size_t nb = align_up((size_t)std::max(Cout, L.cout_pad), 32) + 32;
      std::vector<float> b(nb, 0.f);
      for (size_t o = 0; o < L.b_host.size(); ++o) b[o] = L.b_host[o];
      FLD_CUDA(cudaMalloc(&L.d_bias, nb * sizeof(float)));
      FLD_CUDA(cudaMemcpy(L.d_bias, b.data(), nb * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (L.d.op == FLD_OP_DENSE && L.path == PATH_TC_TMA) {
      // [cout_pad][In] K-major bf16  <-  w[In][Out]; the fp32 copy stays for batches too small for a 128-row tile
      const size_t In = a.elems();
      std::vector<uint16_t> pk((size_t)L.cout_pad * In, 0);
      for (size_t k = 0; k < In; ++k)
        for (int o = 0; o < Cout; ++o) pk[(size_t)o * In + k] = f2bf(L.w_host[k * Cout + o]);
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      FLD_CUDA(cudaMalloc(&L.d_zero, (size_t)(L.cout_pad + 64) * sizeof(float)));
      FLD_CUDA(cudaMemset(L.d_zero, 0, (size_t)(L.cout_pad + 64) * sizeof(float)));
      FLD_CUDA(cudaMalloc(&L.d_w, L.w_host.size() * sizeof(float)));
      FLD_CUDA(cudaMemcpy(L.d_w, L.w_host.data(), L.w_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    } else if (L.d.op == FLD_OP_DECONV && L.path == PATH_TC_TMA) {
      std::vector<uint16_t> pk;
      tc_deconv_pack_weights(L.w_host.data(), L.d.stride, a.c, Cout, f2bf, pk);
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (L.path == PATH_SIMT) {
      FLD_CUDA(cudaMalloc(&L.d_w, L.w_host.size() * sizeof(float)));
      FLD_CUDA(cudaMemcpy(L.d_w, L.w_host.data(), L.w_host.size() * sizeof(float), cudaMemcpyHostToDevice));
      if (L.d.op == FLD_OP_DECONV && L.x3_cmap) {
        std::vector<uint16_t> pk;
        tc_deconv_x3_pack_weights(L.w_host.data(), L.d.stride, a.c, Cout, f2bf, bf2f, pk);
        FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
        FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
        tc_deconv_x3_pack_weights_hilo(L.w_host.data(), L.d.stride, a.c, Cout, f2bf, bf2f, pk);
        FLD_CUDA(cudaMalloc(&L.d_wbf_hilo, pk.size() * 2));
        FLD_CUDA(cudaMemcpy(L.d_wbf_hilo, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
        tc_deconv_pack_weights(L.w_host.data(), L.d.stride, a.c, Cout, f2bf, pk);
        FLD_CUDA(cudaMalloc(&L.d_wbf_hi, pk.size() * 2));
        FLD_CUDA(cudaMemcpy(L.d_wbf_hi, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      }
    } else if (L.path == PATH_TC_FIRST && !tc_conv_first_supported(L.g)) {
      // strided stem (tc_conv_stem.cu): [KG][Cout/8][8][8], k' = 4*tap + c
      std::vector<uint16_t> pk((size_t)Cout * 8 * tc_conv_stem_kgroups(L.d.kh), 0);
      tc_conv_stem_pack(L.w_host.data(), L.b_host.empty() ? nullptr : L.b_host.data(), L.d.kh, Cout, f2bf, pk.data());
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (L.path == PATH_TC_FIRST && tc_conv_s2d_supported(L.g) && (a.dtype == FLD_U8 || !L.x3)) {
      // tc_conv_s2d.cu: 40 K groups ordered per (pool position, tap pair); x3: a second block with the weight remainders
      L.s2d = true;
      std::vector<uint16_t> pk((size_t)Cout * 8 * (L.x3 ? 80 : 40), 0);
      tc_conv_s2d_pack(L.w_host.data(), L.b_host.empty() ? nullptr : L.b_host.data(), Cout, f2bf, pk.data(), L.x3 ? 1 : 0);
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (L.path == PATH_TC_FIRST) {
      // core-matrix packed [6 kgroups][Cout/8][8][8], k' = kh*12 + kw*4 + c (see tc_conv_first.cu)
      const int kg = L.x3 ? (a.dtype == FLD_U8 ? 10 : 16) : 6;   // x3 with a float input splits the pixels too
      std::vector<uint16_t> pk((size_t)Cout * 8 * kg, 0);
      tc_conv_first_pack(L.w_host.data(), L.b_host.empty() ? nullptr : L.b_host.data(), Cout, f2bf, pk.data(), kg);
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (L.x3) {
      // FLD_BF16X3: [tap][cout_pad][3 Cin] = [w_hi | w_hi | w_lo] against the activation chunks [x_hi | x_lo | x_hi]
      const int taps = L.d.kh * L.d.kw, Cin = a.c, cp = L.cout_pad;
      std::vector<uint16_t> pk((size_t)taps * cp * 3 * Cin, 0);
      for (int t = 0; t < taps; ++t)
        for (int c = 0; c < Cin; ++c)
          for (int o = 0; o < Cout; ++o) {
            const float w = L.w_host[((size_t)t * Cin + c) * Cout + o];
            const uint16_t hb = f2bf(w);
            const uint32_t hu = (uint32_t)hb << 16;
            float hf;
            memcpy(&hf, &hu, 4);
            uint16_t* row = &pk[((size_t)t * cp + o) * 3 * Cin];
            row[c] = hb; row[Cin + c] = hb; row[2 * Cin + c] = f2bf(w - hf);
          }
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else {
      // [tap][cout_pad][Cin]  <-  w[(tap*Cin + c)][cout]
      const int taps = L.d.kh * L.d.kw, Cin = a.c, cp = L.cout_pad;
      std::vector<uint16_t> pk((size_t)taps * cp * Cin, 0);
      for (int t = 0; t < taps; ++t)
        for (int c = 0; c < Cin; ++c)
          for (int o = 0; o < Cout; ++o) pk[((size_t)t * cp + o) * Cin + c] = f2bf(L.w_host[((size_t)t * Cin + c) * Cout + o]);
      FLD_CUDA(cudaMalloc(&L.d_wbf, pk.size() * 2));
      FLD_CUDA(cudaMemcpy(L.d_wbf, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    }
  }
  net->finalized = true;
  return FLD_OK;
}

// split-K factor of the tensor-core dense layer: the largest divisor of the k-chunk count <= 32 that leaves at least two k-blocks
// per slice.  It must NOT depend on the batch: a face's result has to be bit-identical whatever batch (chunk, shard) it runs in,
// and the slice boundaries fix the summation order.  For the same reason every batch size takes this path.
static int dense_ksplit(const fld_net*, int, int kchunks) {
  int best = 1;
  for (int s = 1; s <= 32 && s <= kchunks; ++s)
    if (kchunks % s == 0 && kchunks / s >= 2) best = s;
  return best;
}
constexpr int kDenseTcMinBatch = 1;

static size_t dense_scratch_bytes(const fld_net* net, int B) {
  size_t m = 0;
  for (size_t i = 0; i < net->layers.size(); ++i)
  {
    const LayerRt& L = net->layers[i];
    const TensorInfo& a = net->tensors[L.d.in0];
    if (L.d.op == FLD_OP_DENSE) m = std::max(m, simt_dense_scratch_bytes(B, (int)a.elems(), L.d.cout));
    if (L.d.op == FLD_OP_DENSE && L.path == PATH_TC_TMA)
      m = std::max(m, (size_t)dense_ksplit(net, B, (int)a.elems() / 64) * B * L.d.cout * sizeof(float));
    if (L.d.op == FLD_OP_CONV && L.path == PATH_TC_FIRST && L.s2d) m = std::max(m, tc_conv_s2d_scratch_bytes(L.g, B));
    if (L.d.op == FLD_OP_DECONV && L.path == PATH_TC_TMA)
      m = std::max(m, align_up(tc_deconv_scratch_bytes(B, a.h, a.w, a.c), 256) + tc_deconv_acc_bytes(B, L.d.cout));
    if (L.d.op == FLD_OP_DECONV && L.x3_cmap) {
      m = std::max(m, tc_deconv_x3_scratch_bytes(B, a.h, a.w, a.c));
      m = std::max(m, 2 * align_up(tc_deconv_scratch_bytes(B, a.h, a.w, a.c), 256) + tc_deconv_acc_bytes(B, L.d.cout));   // two passes: x_hi, x_lo, sums
    }
  }
  return m;
}

extern "C" int fld_net_tensor_shape(const fld_net* net, int tensor, int32_t* hwc) {
  if (!net || !hwc || tensor < 0 || tensor >= (int)net->tensors.size()) { fld_set_error("fld_net_tensor_shape: bad argument"); return FLD_ERR_INVALID; }
  const TensorInfo& t = net->tensors[tensor];
  hwc[0] = t.h; hwc[1] = t.w; hwc[2] = t.c;
  return t.dtype;
}

extern "C" int64_t fld_net_tensor_offset(const fld_net* net, int tensor, int B) {
  if (!net || tensor < 1 || tensor >= (int)net->tensors.size() || B < 0) { fld_set_error("fld_net_tensor_offset: bad argument"); return FLD_ERR_INVALID; }
  size_t off = 0;
  for (int t = 1; t < tensor; ++t) off += tensor_ws_bytes(net, t, B);
  return (int64_t)off;
}

extern "C" size_t fld_net_workspace_bytes(const fld_net* net, int B) {
  if (!net || B < 0) return 0;
  size_t off = 0;
  for (size_t t = 1; t < net->tensors.size(); ++t) off += tensor_ws_bytes(net, (int)t, B);
  return off + align_up(dense_scratch_bytes(net, B), 1024) + 1024;
}

// fld_net_forward_landmarks: the plain forward's workspace followed by the partials of the stand-alone heat-map decode
// (used whenever the centroid is not fused into the last transposed conv's epilogue)
extern "C" size_t fld_net_landmarks_workspace_bytes(const fld_net* net, int B, int n_points) {
  if (!net || B < 0) return 0;
  const TensorInfo& o = net->tensors.back();
  return fld_net_workspace_bytes(net, B) + align_up(fld_decode_heatmap_scratch_bytes(net->h, B, o.h, o.w, o.c, n_points < 1 ? 0 : n_points), 1024);
}

extern "C" int fld_decode_classmap(fld_handle* h, const float* scores, int B, int hw, int L, int64_t* class_map, fld_stream stream);

// byte offset of the scratch region (dense partials / transposed-conv im2col / first-layer staging) inside the workspace
static size_t scratch_offset(const fld_net* net, int B) {
  size_t off = 0;
  for (size_t t = 1; t < net->tensors.size(); ++t) off += tensor_ws_bytes(net, (int)t, B);
  return off;
}

static int net_forward(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, float* out, int64_t* cmap_out, fld_stream stream,
                       double* xy_out = nullptr, double xy_thresh = 0.0, int xy_n = 0, bool staged = false) {
  FLD_REQUIRE(net, "fld_net_forward: null net");
  int rc = fld_enter(net->h);
  if (rc) return rc;
  if (!net->finalized) { fld_set_error("fld_net_forward: call fld_net_finalize first"); return FLD_ERR_STATE; }
  FLD_REQUIRE(B >= 0, "fld_net_forward: negative batch");
  if (B == 0) return FLD_OK;   // an empty batch is legal (empty tensors have null data pointers)
  FLD_REQUIRE(in && workspace, "fld_net_forward: null pointer");
  FLD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "fld_net_forward: workspace must be 1024-byte aligned");
  const size_t ws_need = xy_out ? fld_net_landmarks_workspace_bytes(net, B, xy_n) : fld_net_workspace_bytes(net, B);
  if (ws_bytes < ws_need) {
    fld_set_error("fld_net_forward: workspace %zu < required %zu", ws_bytes, ws_need);
    return FLD_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nT = (int)net->tensors.size();
  std::vector<void*> ptr(nT);
  ptr[0] = const_cast<void*>(in);
  float* dense_scratch = nullptr;
  bool cmap_done = false, xy_done = false, softmax_done = false;
  {
    size_t off = 0;
    for (int t = 1; t < nT; ++t) {
      ptr[t] = (char*)workspace + off;
      off += tensor_ws_bytes(net, t, B);
    }
    dense_scratch = (float*)((char*)workspace + off);
  }
  // the final tensor is produced straight into the caller's buffer (no device-to-device copy of the probabilities)
  const bool direct_out = out && net->tensors[nT - 1].dtype == FLD_F32;
  if (direct_out) ptr[nT - 1] = out;
  static const bool debug_sync = getenv("FLD_DEBUG_SYNC") != nullptr;
  if (net->profiling) FLD_CUDA(cudaEventRecord(net->events[0], st));
  for (size_t i = 0; i < net->layers.size(); ++i) {
    LayerRt& L = net->layers[i];
    const fld_layer_desc& d = L.d;
    const TensorInfo& a = net->tensors[d.in0];
    const TensorInfo& o = net->tensors[i + 1];
    const void* pin = ptr[d.in0];
    void* pout = ptr[i + 1];
    switch (d.op) {
      case FLD_OP_CONV:
        if (L.path == PATH_TC_FIRST && !tc_conv_first_supported(L.g)) {
          rc = tc_conv_stem(net->h, pin, a.dtype, L.d_wbf, (__nv_bfloat16*)pout, L.g, B, st);
        } else if (L.path == PATH_TC_FIRST && L.s2d) {
          TcS2dPlan* plan = nullptr;
          for (auto& pe : L.splans) if (pe.B == B && pe.scratch == (const void*)dense_scratch) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_conv_s2d_plan_create(net->h, dense_scratch, a.dtype, L.g, B, L.x3 ? 1 : 0, o.dtype == FLD_BF16X3 ? 1 : 0, &plan);
            if (rc) return rc;
            if (L.splans.size() >= 16 && net->retained == 0) { tc_conv_s2d_plan_destroy(L.splans.front().plan); L.splans.erase(L.splans.begin()); }
            L.splans.push_back({B, (const void*)dense_scratch, plan});
          }
          rc = tc_conv_s2d_run(plan, pin, L.d_wbf, pout, st, (staged && i == 0) ? 1 : 0);
        } else if (L.path == PATH_TC_FIRST) {
          rc = tc_conv_first(net->h, pin, a.dtype, L.d_wbf, L.d_bias, (__nv_bfloat16*)pout, L.g, B, st, L.x3 ? 1 : 0);
        } else if (L.path == PATH_TC_TMA && (o.dtype == FLD_BF16 || o.dtype == FLD_BF16X3) && tc_halo_supported(L.g, L.cout_pad)) {
          TcHaloPlan* plan = nullptr;
          for (auto& pe : L.hplans) if (pe.B == B && pe.in == pin) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_halo_plan_create(net->h, pin, L.d_wbf, L.cout_pad, L.g, B, &plan, L.x3 ? 1 : 0, o.dtype == FLD_BF16X3 ? 1 : 0);
            if (rc) return rc;
            if (L.hplans.size() >= 16 && net->retained == 0) { tc_halo_plan_destroy(L.hplans.front().plan); L.hplans.erase(L.hplans.begin()); }
            L.hplans.push_back({B, pin, plan});
          }
          rc = tc_halo_run(plan, L.d_bias, pout, st);
        } else if (L.path == PATH_TC_TMA) {
          TcConvPlan* plan = nullptr;
          for (auto& pe : L.plans) if (pe.B == B && pe.in == pin) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_conv_plan_create(net->h, pin, L.d_wbf, L.cout_pad, L.g, B, &plan, L.x3 ? 1 : 0, o.dtype == FLD_BF16X3 ? 1 : 0);
            if (rc) return rc;
            if (L.plans.size() >= 16 && net->retained == 0) { tc_conv_plan_destroy(L.plans.front().plan); L.plans.erase(L.plans.begin()); }
            L.plans.push_back({B, pin, plan});
          }
          rc = tc_conv_run(plan, L.d_bias, pout, o.dtype == FLD_BF16X3 ? FLD_BF16 : o.dtype, st);
        } else {
          rc = simt_conv(pin, a.dtype, L.d_w, L.b_host.empty() ? nullptr : L.d_bias, pout, o.dtype, L.g, B, st);
        }
        break;
      case FLD_OP_DECONV:
        if (L.path == PATH_TC_TMA) {
          // fused decode: 1 = softmax written into the (skipped) SOFTMAX layer's tensor, 2 = int64 class map to the caller
          const bool last_pair = L.dc_fuse_softmax && (int)i + 2 == (int)net->layers.size();
          // fused landmark decode: the soft centroid (n_points < 1); top-n decodes the materialised probabilities (a fused
          // variant measured slower, see tc_deconv.cu)
          const bool xy_fused = xy_out && last_pair && xy_n < 1;
          const int mode = L.dc_fuse_softmax ? (xy_fused ? 3 : (cmap_out && last_pair) ? 2 : 1) : 0;
          void* dst = mode >= 3 ? (void*)xy_out : mode == 2 ? (void*)cmap_out : (mode == 1 ? ptr[i + 2] : pout);
          float* acc = (float*)((char*)dense_scratch + align_up(tc_deconv_scratch_bytes(B, a.h, a.w, a.c), 256));
          TcDeconvPlan* plan = nullptr;
          for (auto& pe : L.dplans) if (pe.B == B && pe.scratch == (const void*)dense_scratch) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_deconv_plan_create(net->h, dense_scratch, L.d_wbf, B, a.h, a.w, a.c, d.cout, d.stride, &plan);
            if (rc) return rc;
            if (L.dplans.size() >= 16 && net->retained == 0) { tc_deconv_plan_destroy(L.dplans.front().plan); L.dplans.erase(L.dplans.begin()); }
            L.dplans.push_back({B, (const void*)dense_scratch, plan});
          }
          rc = tc_deconv_run(plan, (const float*)pin, dst, mode, st, acc, xy_thresh);
          if (mode == 2) cmap_done = true;
          if (mode >= 3) xy_done = true;
        } else if (L.x3_cmap && cmap_out && !out && !xy_out) {
          // class map only: split-operand tensor-core transposed conv, argmax in the epilogue; the SOFTMAX layer that follows is skipped
          TcDeconvPlan* plan = nullptr;
          for (auto& pe : L.dplans) if (pe.B == B && pe.scratch == (const void*)dense_scratch && pe.kind == 1) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_deconv_plan_create(net->h, dense_scratch, L.d_wbf, B, a.h, a.w, a.c, d.cout, d.stride, &plan, 1);
            if (rc) return rc;
            if (L.dplans.size() >= 16 && net->retained == 0) { tc_deconv_plan_destroy(L.dplans.front().plan); L.dplans.erase(L.dplans.begin()); }
            L.dplans.push_back({B, (const void*)dense_scratch, plan, 1});
          }
          rc = tc_deconv_run(plan, (const float*)pin, cmap_out, 2, st, nullptr, 0.0);
          cmap_done = true;
        } else if (L.x3_cmap && !getenv("FLD_X3_DECONV_2PASS_OFF")) {
          // probabilities / soft centroid in the fp32-accurate mode: two tensor-core passes (tc_deconv.cu).  Pass 1: x_hi against
          // [w_hi | w_lo] -> fp32 logits (this layer's tensor); pass 2: x_lo against w_hi, pass 1's logits added in the epilogue,
          // softmax (-> the SOFTMAX layer's tensor, that layer is then skipped) or the fused soft centroid.
          const size_t a_bytes = align_up(tc_deconv_scratch_bytes(B, a.h, a.w, a.c), 256);
          void* scr_hi = dense_scratch;
          void* scr_lo = (char*)dense_scratch + a_bytes;
          float* acc = (float*)((char*)dense_scratch + 2 * a_bytes);
          TcDeconvPlan *p1 = nullptr, *p2 = nullptr;
          for (auto& pe : L.dplans) {
            if (pe.B == B && pe.scratch == (const void*)scr_hi && pe.kind == 2) p1 = pe.plan;
            if (pe.B == B && pe.scratch == (const void*)scr_lo && pe.kind == 3) p2 = pe.plan;
          }
          if ((!p1 || !p2) && L.dplans.size() >= 24 && net->retained == 0) {        // bounded cache: drop everything for other shapes
            for (auto& pe : L.dplans) tc_deconv_plan_destroy(pe.plan);
            L.dplans.clear();
            p1 = p2 = nullptr;
          }
          if (!p1) {
            rc = tc_deconv_plan_create(net->h, scr_hi, L.d_wbf_hilo, B, a.h, a.w, a.c, d.cout, d.stride, &p1, 2);
            if (rc) return rc;
            L.dplans.push_back({B, (const void*)scr_hi, p1, 2});
          }
          if (!p2) {
            rc = tc_deconv_plan_create(net->h, scr_lo, L.d_wbf_hi, B, a.h, a.w, a.c, d.cout, d.stride, &p2, 3);
            if (rc) return rc;
            L.dplans.push_back({B, (const void*)scr_lo, p2, 3});
          }
          rc = tc_deconv_run(p1, (const float*)pin, pout, 0, st);
          if (rc) return rc;
          const bool xy_fused = xy_out && xy_n < 1;
          rc = tc_deconv_run(p2, (const float*)pin, xy_fused ? (void*)xy_out : ptr[i + 2], xy_fused ? 3 : 1, st, acc, xy_thresh, (const float*)pout);
          softmax_done = true;
          if (xy_fused) xy_done = true;
        } else if (d.kh == 2 * d.stride) rc = simt_deconv_phase(pin, a.dtype, L.d_w, (float*)pout, B, a.h, a.w, a.c, o.c, d.stride, st);
        else rc = simt_deconv(pin, a.dtype, L.d_w, (float*)pout, B, a.h, a.w, a.c, o.h, o.w, o.c, d.kh, d.stride, st);
        break;
      case FLD_OP_ADD: {
        const TensorInfo& b2 = net->tensors[d.in1];
        if (a.dtype == FLD_F32 && b2.dtype == FLD_F32 && o.dtype == FLD_F32 && d.act == FLD_ACT_NONE)
          rc = simt_add_crop((const float*)pin, a.h, a.w, (const float*)ptr[d.in1], b2.h, b2.w, (float*)pout, B, o.h, o.w, o.c, st);
        else
          rc = simt_add_act(pin, a.dtype, a.h, a.w, ptr[d.in1], b2.dtype, b2.h, b2.w, pout, o.dtype, B, o.h, o.w, o.c, d.act, st);
        break;
      }
      case FLD_OP_DENSE:
        if (L.path == PATH_TC_TMA && B >= kDenseTcMinBatch) {
          const int ks = dense_ksplit(net, B, (int)a.elems() / 64);
          TcConvPlan* plan = nullptr;
          for (auto& pe : L.plans) if (pe.B == B && pe.in == pin) { plan = pe.plan; break; }
          if (!plan) {
            rc = tc_conv_plan_create(net->h, pin, L.d_wbf, L.cout_pad, L.g, B, &plan, 0, 0, ks);
            if (rc) return rc;
            if (L.plans.size() >= 16 && net->retained == 0) { tc_conv_plan_destroy(L.plans.front().plan); L.plans.erase(L.plans.begin()); }
            L.plans.push_back({B, pin, plan});
          }
          rc = tc_conv_run(plan, L.d_zero, dense_scratch, FLD_F32, st);    // partial sums [ks][B][Out]
          if (rc) return rc;
          rc = simt_dense_reduce(dense_scratch, L.b_host.empty() ? nullptr : L.d_bias, (float*)pout, B, o.c, ks, d.act, st);
          break;
        }
        rc = simt_dense(pin, a.dtype, L.d_w, L.b_host.empty() ? nullptr : L.d_bias, (float*)pout, dense_scratch, B, (int)a.elems(), o.c, d.act,
                        st, a.c);
        break;
      case FLD_OP_SOFTMAX:
        if (L.skip) break;  // computed by the preceding transposed conv's epilogue
        if (cmap_done && !out && !xy_out && i + 1 == net->layers.size()) break;   // only the class map was asked for, and it is done
        if (softmax_done && i + 1 == net->layers.size()) break;                  // computed by the two-pass transposed conv's epilogue
        FLD_REQUIRE(a.dtype == FLD_F32, "layer %zu: SOFTMAX input must be fp32", i);
        rc = simt_softmax((const float*)pin, (float*)pout, (long long)B * a.h * a.w, a.c, st);
        break;
      case FLD_OP_MAXPOOL:
        if (a.dtype == FLD_F32 && o.dtype == FLD_F32) rc = simt_maxpool((const float*)pin, (float*)pout, B, a.h, a.w, a.c, o.h, o.w, d.kh, d.stride, st);
        else rc = simt_maxpool2d(pin, a.dtype, pout, o.dtype, B, a.h, a.w, a.c, o.h, o.w, d.kh, d.stride, st);
        break;
      case FLD_OP_DWCONV:
        rc = simt_dwconv(pin, a.dtype, L.d_w, L.b_host.empty() ? nullptr : L.d_bias, pout, o.dtype, B, a.h, a.w, a.c, o.h, o.w, d.kh, d.kw,
                         d.stride, d.pad_t, d.pad_l, d.act, st);
        break;
      default:
        fld_set_error("layer %zu: op %d not implemented", i, d.op);
        rc = FLD_ERR_INVALID;
    }
    if (rc) return rc;
    if (debug_sync) {  // FLD_DEBUG_SYNC=1: localise an asynchronous fault to its layer
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        fld_set_error("layer %zu (op %d, path %d, in %dx%dx%d -> out %dx%dx%d, B=%d) failed: %s", i, d.op, L.path, a.h, a.w, a.c, o.h,
                      o.w, o.c, B, cudaGetErrorString(e));
        return FLD_ERR_CUDA;
      }
    }
    if (net->profiling) FLD_CUDA(cudaEventRecord(net->events[i + 1], st));
  }
  if (net->profiling) net->profiled_once = true;
  if (cmap_out && !cmap_done) {  // no fused argmax available: class map from the final tensor
    const TensorInfo& o = net->tensors[nT - 1];
    FLD_REQUIRE(o.dtype == FLD_F32, "fld_net_forward_classmap: final tensor must be fp32");
    rc = fld_decode_classmap(net->h, (const float*)ptr[nT - 1], B, o.h * o.w, o.c, cmap_out, stream);
    if (rc) return rc;
  }
  if (xy_out && !xy_done) {  // no fused centroid available: soft-centroid decode of the final tensor
    const TensorInfo& o = net->tensors[nT - 1];
    FLD_REQUIRE(o.dtype == FLD_F32, "fld_net_forward_landmarks: final tensor must be fp32");
    const size_t base = fld_net_workspace_bytes(net, B);
    rc = fld_decode_heatmap_xy(net->h, (const float*)ptr[nT - 1], B, o.h, o.w, o.c, xy_n < 1 ? 0 : xy_n, xy_thresh, xy_out,
                               (char*)workspace + base, ws_bytes - base, stream);
    if (rc) return rc;
  }
  if (out && !direct_out) {
    const TensorInfo& o = net->tensors[nT - 1];
    const size_t n = o.elems() * (size_t)B;
    if (o.dtype == FLD_F32) FLD_CUDA(cudaMemcpyAsync(out, ptr[nT - 1], n * 4, cudaMemcpyDeviceToDevice, st));
    else { rc = simt_cvt_bf16_f32(ptr[nT - 1], out, (long long)n, st); if (rc) return rc; }
  }
  return FLD_OK;
}

extern "C" int fld_net_forward_landmarks(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, int n_points, double thresh,
                                         double* xy, fld_stream stream) {
  FLD_REQUIRE(B == 0 || xy, "fld_net_forward_landmarks: null output");
  FLD_REQUIRE(n_points <= FLD_MAX_TOPN, "fld_net_forward_landmarks: n_points must be <= %d", FLD_MAX_TOPN);
  return net_forward(net, in, B, workspace, ws_bytes, nullptr, nullptr, stream, xy, thresh, n_points);
}

extern "C" int fld_net_forward(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, float* out, fld_stream stream) {
  return net_forward(net, in, B, workspace, ws_bytes, out, nullptr, stream);
}

extern "C" int fld_net_input_staging(const fld_net* net, int B, void* workspace, void** staging) {
  FLD_REQUIRE(net && staging, "fld_net_input_staging: null pointer");
  *staging = nullptr;
  if (!net->finalized || B <= 0 || !workspace || net->layers.empty()) return FLD_OK;
  const LayerRt& L = net->layers[0];
  if (L.d.op == FLD_OP_CONV && L.path == PATH_TC_FIRST && L.s2d && L.d.in0 == 0 && net->tensors[0].dtype == FLD_U8)
    *staging = (char*)workspace + scratch_offset(net, B);
  return FLD_OK;
}

extern "C" int fld_net_forward_staged(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, float* out, fld_stream stream) {
  void* stg = nullptr;
  int rc = fld_net_input_staging(net, B, workspace, &stg);
  if (rc) return rc;
  FLD_REQUIRE(B == 0 || stg, "fld_net_forward_staged: this net's first layer takes no staged input (fld_net_input_staging returned NULL)");
  return net_forward(net, in, B, workspace, ws_bytes, out, nullptr, stream, nullptr, 0.0, 0, true);
}

extern "C" int fld_net_forward_classmap(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, int64_t* class_map,
                                        fld_stream stream) {
  if (!class_map) { fld_set_error("fld_net_forward_classmap: null class_map"); return FLD_ERR_INVALID; }
  return net_forward(net, in, B, workspace, ws_bytes, nullptr, class_map, stream);
}

// A captured CUDA graph bakes in device pointers owned by the kernel plans (tile schedules): while a net is retained no plan
// is evicted, so those pointers stay valid for the graph's lifetime.  Balanced by fld_net_release.
extern "C" int fld_net_retain(fld_net* net) {
  FLD_REQUIRE(net, "fld_net_retain: null net");
  return ++net->retained;
}
extern "C" int fld_net_release(fld_net* net) {
  FLD_REQUIRE(net, "fld_net_release: null net");
  if (net->retained > 0) --net->retained;
  return net->retained;
}

extern "C" int fld_net_set_profiling(fld_net* net, int enable) {
  FLD_REQUIRE(net, "fld_net_set_profiling: null net");
  int rc = fld_enter(net->h);
  if (rc) return rc;
  if (enable && net->events.empty()) {
    net->events.resize(net->layers.size() + 1);
    for (auto& e : net->events) FLD_CUDA(cudaEventCreate(&e));
  }
  net->profiling = enable != 0;
  net->profiled_once = false;
  return FLD_OK;
}

extern "C" int fld_net_layer_times(fld_net* net, float* ms_h, int n) {
  FLD_REQUIRE(net && ms_h, "fld_net_layer_times: null pointer");
  FLD_REQUIRE(n >= (int)net->layers.size(), "fld_net_layer_times: need room for %zu layers", net->layers.size());
  if (!net->profiled_once) { fld_set_error("fld_net_layer_times: no profiled forward yet"); return FLD_ERR_STATE; }
  int rc = fld_enter(net->h);
  if (rc) return rc;
  FLD_CUDA(cudaEventSynchronize(net->events.back()));
  for (size_t i = 0; i < net->layers.size(); ++i) FLD_CUDA(cudaEventElapsedTime(&ms_h[i], net->events[i], net->events[i + 1]));
  return (int)net->layers.size();
}
