/* fld.h — C-ABI of libfld_sm100.so: the B200 (sm_100a) hot path of
 * sandyz1000/face-landmark-detector (landmark-CNN forward -> landmark decode -> similarity
 * align warp).
 *
 * The reference has no FFI: its boundary is the Python API of keypoints_detector.prediction
 * (SURVEY.md §8b).  Every entry point below names the reference call site whose arithmetic it
 * replaces (paths relative to the reference root).  The Python host package
 * (face-landmark-detector_b200/keypoints_detector) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; return 0 on success, negative fld_status on error, message via
 *     fld_last_error() (thread-local).
 *   - every pointer is a DEVICE pointer on the handle's device unless its name ends in _h (host).
 *   - the caller owns all buffers; nothing is allocated per call except inside fld_net (weights,
 *     tensor maps) which fld_net_destroy releases.
 *   - every compute call takes a cudaStream_t (as void*) and is asynchronous on that stream.
 *   - one handle per GPU; a handle is not thread-safe; distinct handles may be driven from
 *     distinct host threads (multi-GPU sharding).
 *   - there is NO CPU fallback: on a machine without a usable sm_100 device fld_create fails.
 */
#ifndef FLD_H_
#define FLD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLD_ABI_VERSION 2

#if defined(__GNUC__)
#define FLD_API __attribute__((visibility("default")))
#else
#define FLD_API
#endif

typedef enum {
  FLD_OK = 0,
  FLD_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  FLD_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  FLD_ERR_NODEVICE = -3,  /* no sm_100 device */
  FLD_ERR_WORKSPACE = -4, /* workspace too small */
  FLD_ERR_STATE = -5      /* call order (e.g. forward before finalize) */
} fld_status;

typedef enum { FLD_U8 = 0, FLD_F32 = 1, FLD_BF16 = 2, FLD_BF16X3 = 3 } fld_dtype;

/* compute modes of a network:
 *   FLD_F32    fp32 CUDA-core kernels — the reference's arithmetic (TensorFlow fp32, prediction.py:84,208); <=0.05 px bar;
 *   FLD_BF16   tcgen05 tensor-core kernels, bf16 operands / fp32 accumulate; <=0.5 px bar;
 *   FLD_BF16X3 fp32-ACCURATE tensor-core mode: every conv operand is split into bf16 hi + lo parts and the product is formed as
 *              x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation in TMEM (relative error ~2^-17 per product, 250x below
 *              plain bf16); activations travel between tensor-core convs as SPLIT tensors [hi(C) | lo(C)] bf16 per pixel
 *              (fld_net_tensor_shape reports them as FLD_BF16X3).  Layers the tensor-core kernels do not cover run the FLD_F32
 *              kernels.  Meets the <=0.05 px bar; the drop-in package's default.  Segmentation nets: a float32 first-layer input is
 *              split like the weights, and fld_net_forward_classmap runs the final transposed conv on the tensor cores with split
 *              operands and takes the argmax in its epilogue (the softmax is skipped: argmax is invariant under it); probability and
 *              landmark outputs take two tensor-core passes (x_hi against [w_hi | w_lo], then x_lo against w_hi with the first
 *              pass's logits added before the softmax / soft-centroid epilogue). */

typedef struct fld_handle fld_handle;
typedef struct fld_net fld_net;
typedef void* fld_stream; /* cudaStream_t */

FLD_API int fld_abi_version(void);
FLD_API const char* fld_last_error(void);
FLD_API int fld_create(int device, fld_handle** out);
FLD_API void fld_destroy(fld_handle* h);

/* ------------------------------------------------------------------------------------------
 * Pre-processing
 * ------------------------------------------------------------------------------------------ */

/* Replaces prediction.py:76-83 (move_box, get_square_box, crop, cv2.resize(128,128), BGR2RGB) for a
 * batch of faces.  frames: uint8 [F,H,W,3] BGR contiguous.  boxes: int32 [B,4] raw detector boxes
 * (x0,y0,x1,y1).  face2frame: int32 [B].  out: uint8 [B,S,S,3] (RGB when swap_rb).  faceboxes:
 * int32 [B,4] squared boxes (what prediction.py:91-93 scales the landmarks by).  The crop is the
 * intersection of the squared box with the frame (numpy slicing semantics for boxes leaving the
 * bottom/right; boxes leaving the top/left, where the reference crashes, are clipped the same way).
 * Resize is bit-exact cv2.resize(INTER_LINEAR) of OpenCV 4.13 incl. its exact-2x area shortcut. */
FLD_API int fld_preprocess_faces(fld_handle* h, const uint8_t* frames, int F, int H, int W,
                         const int32_t* boxes, const int32_t* face2frame, int B, int S, int swap_rb,
                         uint8_t* out, int32_t* faceboxes, fld_stream stream);

/* Same, and additionally writes every resized pixel into the first conv layer's operand staging buffer of a network
 * (fld_net_input_staging: space-to-depth planes of the crop widened to 8 bf16 channels), so that fld_net_forward_staged can skip
 * its own widening pass over the crops.  `out` is still written (it is what prediction.py:84 hands the model).  S % 4 == 0. */
FLD_API int fld_preprocess_faces_staged(fld_handle* h, const uint8_t* frames, int F, int H, int W,
                         const int32_t* boxes, const int32_t* face2frame, int B, int S, int swap_rb,
                         uint8_t* out, int32_t* faceboxes, void* staging, fld_stream stream);

/* Replaces data/generator.py:50-69 get_image_array on a batch of equally sized images.
 * images: uint8 [B,H,W,3] BGR.  norm: 0 = sub_mean (:53-61; subtract [103.939,116.779,123.68] per
 * BGR channel, then reverse channels), 1 = sub_and_divide (:51; /127.5-1), 2 = divide (:63-65; /255).
 * out: float32 [B,oh,ow,3] channels_last. */
FLD_API int fld_image_array(fld_handle* h, const uint8_t* images, int B, int H, int W, int ow, int oh, int norm,
                    float* out, fld_stream stream);

/* ------------------------------------------------------------------------------------------
 * CNN forward: replaces model.signatures["predict"] (prediction.py:84) and
 * model.predict (prediction.py:208) for graphs built from networks/*.py vocabulary.
 * ------------------------------------------------------------------------------------------ */

typedef enum {
  FLD_OP_CONV = 0,    /* Conv2D (+ZeroPadding2D) [+BatchNormalization folded] [+ReLU/ReLU6] [+MaxPool 2x2] */
  FLD_OP_DECONV = 1,  /* Conv2DTranspose padding='valid', no bias (fcn.py:104,114,121,145) */
  FLD_OP_ADD = 2,     /* fcn.py:55-86 crop + KL.Add: both inputs cropped (bottom/right) to the smaller; with `act`:
                         the residual add -> ReLU of resnet50.py:68-69,117-118 */
  FLD_OP_DENSE = 3,   /* Flatten (H,W,C row-major) + Dense */
  FLD_OP_SOFTMAX = 4, /* networks/utils.py:28-30: softmax over channels, output [B, oh*ow, C] */
  FLD_OP_DWCONV = 5,  /* ZeroPadding2D + DepthwiseConv2D(depth_multiplier 1) [+BN folded] [+ReLU6] (mobilenet.py:37-47);
                         kh, kw, stride, pad_*, act as for CONV; cout = 0 or the input channel count */
  FLD_OP_MAXPOOL = 6  /* stand-alone MaxPooling2D((kh,kh), strides=stride) 'valid' (resnet50.py:149) */
} fld_op;

typedef enum { FLD_ACT_NONE = 0, FLD_ACT_RELU = 1, FLD_ACT_RELU6 = 2 } fld_act;

/* Tensor 0 is the network input; layer i produces tensor i+1. */
typedef struct {
  int32_t op;        /* fld_op */
  int32_t in0, in1;  /* input tensor ids (in1 only for ADD, else -1) */
  int32_t kh, kw, stride;
  int32_t pad_t, pad_b, pad_l, pad_r; /* explicit zero padding (ZeroPadding2D or resolved 'same') */
  int32_t cout;      /* output channels (DENSE: units) */
  int32_t act;       /* fld_act */
  int32_t pool;      /* 2 = fused MaxPooling2D(2,2) after the activation, 0 = none */
  int32_t has_bias, has_bn;
  float in_scale;    /* CONV only: input is multiplied by this before the conv (1/255 for the uint8
                        regression input); folded into the weights */
} fld_layer_desc;

FLD_API int fld_net_create(fld_handle* h, const fld_layer_desc* layers_h, int n_layers, int in_h, int in_w, int in_c,
                   int in_dtype /* FLD_U8 | FLD_F32 */, int compute /* FLD_F32 | FLD_BF16 | FLD_BF16X3 */, fld_net** out);
FLD_API void fld_net_destroy(fld_net* net);
/* Keras layouts, host fp32: Conv2D kernel [kh,kw,Cin,Cout]; Conv2DTranspose [kh,kw,Cout,Cin]; Dense
 * [In,Out]; depthwise [kh,kw,C,1]; bias [Cout] or NULL; bn = gamma|beta|moving_mean|moving_variance
 * (4*Cout floats) or NULL.  BN is folded: w' = w*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta. */
FLD_API int fld_net_set_weights(fld_net* net, int layer, const float* kernel_h, const float* bias_h, const float* bn_h, float eps);
FLD_API int fld_net_finalize(fld_net* net);
/* shape of tensor id for batch 1: hwc[3]; returns dtype (fld_dtype) or negative error */
FLD_API int fld_net_tensor_shape(const fld_net* net, int tensor, int32_t* hwc);
FLD_API size_t fld_net_workspace_bytes(const fld_net* net, int B);
/* Offset (bytes) of a tensor inside the workspace for batch B (debug / parity of intermediate levels) */
FLD_API int64_t fld_net_tensor_offset(const fld_net* net, int tensor, int B);
/* in: uint8 or float32 NHWC [B,in_h,in_w,in_c].  out: float32, final tensor [B, ...]; the last layer writes straight into it
 * (no device-to-device copy).  NULL keeps the result in the workspace only (fld_net_tensor_offset).  The workspace holds
 * the activations of ONE forward: forwards that overlap on different streams need a workspace each; the net object (weights,
 * kernel plans) is shared. */
FLD_API int fld_net_forward(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, float* out, fld_stream stream);

/* Producer-side fusion of the first layer's operand staging.  fld_net_input_staging returns in *staging the address INSIDE
 * `workspace` (for batch B) where the first conv layer keeps its widened copy of a uint8 input — or NULL when the net has no such
 * buffer (other first layers, float inputs).  A producer that writes it (fld_preprocess_faces_staged, same B) lets
 * fld_net_forward_staged run without the widening pass; `in` must still be the uint8 tensor the staging was derived from. */
FLD_API int fld_net_input_staging(const fld_net* net, int B, void* workspace, void** staging);
FLD_API int fld_net_forward_staged(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, float* out, fld_stream stream);

/* Forward of a segmentation graph followed by prediction.py:209 (argmax over classes, first max wins): class_map int64
 * [B, oh, ow].  In FLD_BF16 mode, when the graph ends in Conv2DTranspose(k = 2*stride) + softmax (fcn_8), the argmax runs
 * in the transposed conv's epilogue and neither logits nor probabilities are written to HBM. */
FLD_API int fld_net_forward_classmap(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, int64_t* class_map,
                                     fld_stream stream);

/* Forward of a segmentation graph followed by the landmark decode of every class channel of its softmax output
 * (utils/metrics.py:46-80 get_average_xy applied per channel as transfer_target does, :83-109): n_points < 1 = soft centroid over
 * the whole map, n_points >= 1 = weighted centroid of the n_points largest values.  xy double [B, 2L] = (x0, y0, x1, y1, ...) in
 * heat-map pixels, (-1, -1) where sum / n <= thresh.  In FLD_BF16 mode, when the graph ends in Conv2DTranspose(k = 2*stride) +
 * softmax (fcn_8 / fcn_32), the SOFT centroid is accumulated in the transposed conv's epilogue (fp32 atomics: summation order,
 * hence the last bits, vary from run to run) and the probabilities are never written to HBM.  In every other case (top-n,
 * fp32 mode, other graphs) the final tensor is materialised in the workspace and decoded by fld_decode_heatmap_xy, whose
 * partials follow the activations: the workspace must hold fld_net_landmarks_workspace_bytes(net, B, n_points). */
FLD_API size_t fld_net_landmarks_workspace_bytes(const fld_net* net, int B, int n_points);
FLD_API int fld_net_forward_landmarks(fld_net* net, const void* in, int B, void* workspace, size_t ws_bytes, int n_points, double thresh,
                                      double* xy, fld_stream stream);

/* CUDA-graph lifetime: a graph captured around fld_net_forward* bakes in device pointers owned by the net's cached kernel
 * plans.  While the retain count is > 0 no plan is evicted (the plan caches otherwise keep the 16 most recent (batch, buffer)
 * combinations per layer).  Both return the new count.  The caller keeps workspace and I/O buffers alive itself. */
FLD_API int fld_net_retain(fld_net* net);
FLD_API int fld_net_release(fld_net* net);

/* Per-layer device timing (bench.py's live roofline measurement): when enabled, fld_net_forward brackets
 * every layer with CUDA events on the launching stream; fld_net_layer_times waits for the last profiled
 * forward and writes one duration (ms) per layer, returning the layer count. */
FLD_API int fld_net_set_profiling(fld_net* net, int enable);
FLD_API int fld_net_layer_times(fld_net* net, float* ms_h, int n);

/* ------------------------------------------------------------------------------------------
 * Landmark decode
 * ------------------------------------------------------------------------------------------ */

/* Replaces prediction.py:88-94.  out136: float32 [B, stride] (first 136 used); faceboxes int32 [B,4].
 * marks_f32 [B,68,2] = out*side + (x0,y0) in float32 (pre-cast values); marks_u64 [B,68,2] (nullable) =
 * astype(np.uint) with negatives clamped to 0. */
FLD_API int fld_decode_regress(fld_handle* h, const float* out136, int stride, const int32_t* faceboxes, int B,
                       float* marks_f32, uint64_t* marks_u64, fld_stream stream);

/* Replaces prediction.py:209 (argmax over classes, first max wins).  scores: float32 [B,oh*ow,L]
 * (probabilities or logits — argmax is invariant under softmax).  class_map: int64 [B,oh,ow]. */
FLD_API int fld_decode_classmap(fld_handle* h, const float* scores, int B, int hw, int L, int64_t* class_map, fld_stream stream);

/* Replaces utils/metrics.py:46-109 (get_average_xy / transfer_xy_coord / transfer_target).
 * hm: float32 [B,H,W,L].  n_points < 1: full soft-centroid (:58-64); else top-n weighted centroid
 * (:66-77, ties -> higher flat index).  xy: float64 [B, 2L] = (x0,y0,x1,y1,...), (-1,-1) where
 * sum/n <= thresh (:78-79).  n_points <= FLD_MAX_TOPN.  The decode is two-pass; its partials live in a CALLER-provided
 * scratch buffer of at least fld_decode_heatmap_scratch_bytes(...) bytes, so decodes in flight on different streams (each
 * with its own scratch) never interfere and nothing is allocated, freed or synchronised inside the call (graph-capture safe). */
#define FLD_MAX_TOPN 128
FLD_API size_t fld_decode_heatmap_scratch_bytes(fld_handle* h, int B, int H, int W, int L, int n_points);
FLD_API int fld_decode_heatmap_xy(fld_handle* h, const float* hm, int B, int H, int W, int L, int n_points, double thresh,
                          double* xy, void* scratch, size_t scratch_bytes, fld_stream stream);

/* ------------------------------------------------------------------------------------------
 * Alignment (build-defined — nothing in the reference to replace; semantics =
 * fp64 Umeyama similarity fit + cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of OpenCV 4.13)
 * ------------------------------------------------------------------------------------------ */

/* frames uint8 [F,H,W,C] (C = 1,3,4); face2frame int32 [B]; marks float32 [B,N,2] in frame pixels;
 * tmpl float64 [Nt,2].  five_point != 0: N must be 68 and Nt 5 — the iBUG-68 landmarks are reduced to
 * (left-eye centre 36..41, right-eye centre 42..47, nose 30, mouth corners 48, 54) before the fit;
 * otherwise N == Nt.  M_out float64 [B,2,3] (nullable), crops uint8 [B,out_h,out_w,C].  A degenerate
 * fit gives NaN in M_out and an all-zero crop. */
FLD_API int fld_align(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
              const float* marks, int N, const double* tmpl, int Nt, int five_point, int B, int out_h, int out_w,
              double* M_out, uint8_t* crops, fld_stream stream);

/* Warp only, with caller-supplied forward matrices M float64 [B,2,3] (cv2.warpAffine semantics). */
FLD_API int fld_warp_affine(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                    const double* M, int B, int out_h, int out_w, uint8_t* crops, fld_stream stream);

/* Large batches: the same two calls with a CALLER-provided scratch buffer of at least fld_align_scratch_bytes(h, B) bytes
 * (16-byte aligned).  The fit then runs in a kernel of its own (one warp per face) and the faces are warped in order of decreasing
 * source-box size, which removes the scheduling tail of a mixed batch (config C4: 0.187 -> 0.168 ms per 4096 faces).  Results are
 * bit-identical to fld_align / fld_warp_affine; below 2048 faces the calls behave exactly like them.  Nothing is allocated or
 * synchronised inside the call (graph-capture safe); one scratch buffer per stream in flight. */
FLD_API size_t fld_align_scratch_bytes(fld_handle* h, int B);
FLD_API int fld_align_ordered(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                      const float* marks, int N, const double* tmpl, int Nt, int five_point, int B, int out_h, int out_w,
                      double* M_out, uint8_t* crops, void* scratch, size_t scratch_bytes, fld_stream stream);
FLD_API int fld_warp_affine_ordered(fld_handle* h, const uint8_t* frames, int F, int H, int W, int C, const int32_t* face2frame,
                            const double* M, int B, int out_h, int out_w, uint8_t* crops, void* scratch, size_t scratch_bytes,
                            fld_stream stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches evidence) */
FLD_API uint64_t fld_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLD_H_ */
