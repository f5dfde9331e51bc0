"""Oracle: Keras-semantics CNN forward.  Test infrastructure only (oracle/__init__.py).

PARITY UNPINNED for this file: the arithmetic lives in TensorFlow/Keras (unpinned in the
reference — requirements.txt is empty, setup.py:6,25 — and not installable here), and the
reference holds no tests or golden vectors.  The graphs are restated from the reference's
model definitions with Keras layer semantics:

  vanilla_encoder      networks/fcn.py:10-51
  crop                 networks/fcn.py:55-86
  fcn_8 / fcn_32       networks/fcn.py:89-150
  segmentation softmax networks/utils.py:22-30
  regression net       build-defined: vanilla_encoder@128 -> Flatten -> Dense(136)
  VGG16 encoder        networks/vgg16.py:17-81
  MobileNet-v1 encoder networks/mobilenet.py:12-104
  ResNet50 encoder     networks/resnet50.py:23-173
                       standing in for the opaque SavedModel at prediction.py:84

Two independent restatements are kept for EVERY graph (vanilla / VGG16 / MobileNet / ResNet50 encoders, fcn_8, fcn_32) and must
agree: torch functional ops (``*_t``) and a numpy version (``*_np``) written without torch — sliding windows + einsum for
dense convs (stride by slicing), an explicit tap loop for depth-wise convs, window maxima for pools, a scatter-add loop for
transposed convs.  Tests run the numpy side on small inputs only.

Weights are a dict name -> ndarray in Keras layouts: Conv2D kernel [kh,kw,Cin,Cout], bias [Cout];
BatchNormalization gamma/beta/moving_mean/moving_variance [C]; Conv2DTranspose kernel
[kh,kw,Cout,Cin]; Dense kernel [In,Out].
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon


# ----------------------------------------------------------------------------- torch restatement
def _t(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype)


def conv2d_t(x, kernel, bias=None, stride=1, pad=(0, 0, 0, 0)):
    """x NCHW; kernel HWIO; pad = (top, bottom, left, right) explicit zero padding; 'valid' conv."""
    w = kernel.permute(3, 2, 0, 1).contiguous()
    if any(pad):
        x = F.pad(x, (pad[2], pad[3], pad[0], pad[1]))
    return F.conv2d(x, w, bias, stride=stride)


def same_pad(k):
    """Keras/TF 'same' for stride 1: total k-1, extra on the bottom/right."""
    tot = k - 1
    lo = tot // 2
    return (lo, tot - lo, lo, tot - lo)


def bn_t(x, w, name):
    g, b, m, v = w[name + "/gamma"], w[name + "/beta"], w[name + "/moving_mean"], w[name + "/moving_variance"]
    sh = (1, -1, 1, 1)
    return g.view(sh) * (x - m.view(sh)) / torch.sqrt(v.view(sh) + BN_EPS) + b.view(sh)


def deconv_t(x, kernel, stride):
    """Conv2DTranspose 'valid', no bias: out[n,o,i*s+a,j*s+b] += x[n,c,i,j]*K[a,b,o,c] (no flip)."""
    w = kernel.permute(3, 2, 0, 1).contiguous()  # [Cin, Cout, kh, kw]
    return F.conv_transpose2d(x, w, stride=stride)


def vanilla_encoder_t(x, w, prefix=""):
    """fcn.py:25-49: 5 x (ZeroPad1, Conv3x3 valid + bias, BN, ReLU, MaxPool2).  x NCHW."""
    levels = []
    for i in range(1, 6):
        x = conv2d_t(x, w[f"{prefix}conv{i}/kernel"], w[f"{prefix}conv{i}/bias"], pad=(1, 1, 1, 1))
        x = bn_t(x, w, f"{prefix}bn{i}")
        x = F.relu(x)
        x = F.max_pool2d(x, 2, 2)
        levels.append(x)
    return levels


def _crop_pair(o1, o2):
    """fcn.py:55-86 crop(o1, o2): crop the wider one's right columns and the taller one's bottom rows."""
    h1, w1 = o1.shape[2:]
    h2, w2 = o2.shape[2:]
    cx = abs(w1 - w2); cy = abs(h2 - h1)
    if w1 > w2:
        o1 = o1[:, :, :, : w1 - cx]
    else:
        o2 = o2[:, :, :, : w2 - cx]
    if h1 > h2:
        o1 = o1[:, :, : h1 - cy, :]
    else:
        o2 = o2[:, :, : h2 - cy, :]
    return o1, o2


def fcn_head_t(f5, w):
    """fcn.py:98-103: 7x7 same 4096 relu, (dropout=id), 1x1 4096 relu, 1x1 n_classes."""
    o = F.relu(conv2d_t(f5, w["head7/kernel"], w["head7/bias"], pad=same_pad(7)))
    o = F.relu(conv2d_t(o, w["head1/kernel"], w["head1/bias"]))
    return conv2d_t(o, w["score5/kernel"], w["score5/bias"])


def fcn_8_logits_t(levels, w):
    """fcn.py:96-122 -> pre-softmax logits NCHW [B, n_classes, 8*h3+8, 8*w3+8]."""
    f3, f4, f5 = levels[2], levels[3], levels[4]
    o = fcn_head_t(f5, w)
    o = deconv_t(o, w["up2a/kernel"], 2)                                   # :104
    o2 = conv2d_t(f4, w["score4/kernel"], w["score4/bias"])                # :107-108
    o, o2 = _crop_pair(o, o2)                                              # :110
    o = o + o2                                                             # :112
    o = deconv_t(o, w["up2b/kernel"], 2)                                   # :114
    o2 = conv2d_t(f3, w["score3/kernel"], w["score3/bias"])                # :116-117
    o2, o = _crop_pair(o2, o)                                              # :118
    o = o2 + o                                                             # :119
    return deconv_t(o, w["up8/kernel"], 8)                                 # :121


def fcn_32_logits_t(levels, w):
    """fcn.py:137-146."""
    o = fcn_head_t(levels[4], w)
    return deconv_t(o, w["up32/kernel"], 32)


def segmentation_probs_t(logits):
    """networks/utils.py:28-30: reshape (oh*ow, n) row-major over (y,x), softmax over classes."""
    B, n, oh, ow = logits.shape
    return torch.softmax(logits.permute(0, 2, 3, 1).reshape(B, oh * ow, n), dim=-1)


def _prep(weights, dtype):
    return {k: _t(v, dtype) for k, v in weights.items()}


def regression_forward(x_u8_nhwc, weights, dtype=torch.float64, threads=None):
    """uint8 RGB [B,128,128,3] -> [B,136].  Input scaled by 1/255 (build-defined, App. A.8)."""
    if threads:
        torch.set_num_threads(threads)
    w = _prep(weights, dtype)
    x = _t(x_u8_nhwc, dtype).permute(0, 3, 1, 2) / 255.0
    f5 = vanilla_encoder_t(x, w)[4]
    flat = f5.permute(0, 2, 3, 1).reshape(f5.shape[0], -1)                 # Flatten over (H,W,C)
    out = flat @ w["fc/kernel"] + w["fc/bias"]
    return out.numpy()


def trunk_forward(x_nhwc, weights, dtype=torch.float64, scale=1.0):
    """float NHWC -> list of 5 NHWC level arrays."""
    w = _prep(weights, dtype)
    x = _t(x_nhwc, dtype).permute(0, 3, 1, 2) * scale
    return [l.permute(0, 2, 3, 1).contiguous().numpy() for l in vanilla_encoder_t(x, w)]


def fcn_forward(x_nhwc, weights, variant="fcn_8", dtype=torch.float64, return_logits=False):
    """float NHWC (output of get_image_array) -> probs [B, oh*ow, n] (and logits NHWC)."""
    w = _prep(weights, dtype)
    x = _t(x_nhwc, dtype).permute(0, 3, 1, 2)
    levels = vanilla_encoder_t(x, w)
    logits = fcn_8_logits_t(levels, w) if variant == "fcn_8" else fcn_32_logits_t(levels, w)
    probs = segmentation_probs_t(logits).numpy()
    if return_logits:
        return probs, logits.permute(0, 2, 3, 1).contiguous().numpy()
    return probs


# ----------------------------------------------------------------------------- numpy restatement
def conv2d_np(x, kernel, bias=None, pad=(0, 0, 0, 0)):
    """x NHWC fp64, kernel HWIO; stride 1."""
    x = np.pad(x, ((0, 0), (pad[0], pad[1]), (pad[2], pad[3]), (0, 0)))
    kh, kw = kernel.shape[:2]
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))  # [B,oh,ow,C,kh,kw]
    out = np.einsum("bhwcij,ijco->bhwo", win, kernel, optimize=True)
    return out + bias if bias is not None else out


def bn_np(x, w, name):
    return w[name + "/gamma"] * (x - w[name + "/moving_mean"]) / np.sqrt(w[name + "/moving_variance"] + BN_EPS) \
        + w[name + "/beta"]


def maxpool2_np(x):
    B, H, W, C = x.shape
    x = x[:, : H // 2 * 2, : W // 2 * 2]
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def deconv_np(x, kernel, s):
    B, H, W, C = x.shape
    kh, kw, Co, Ci = kernel.shape
    out = np.zeros((B, (H - 1) * s + kh, (W - 1) * s + kw, Co), dtype=x.dtype)
    for a in range(kh):
        for b in range(kw):
            out[:, a: a + (H - 1) * s + 1: s, b: b + (W - 1) * s + 1: s] += np.einsum("bhwc,oc->bhwo", x, kernel[a, b])
    return out


def vanilla_encoder_np(x, w):
    levels = []
    for i in range(1, 6):
        x = conv2d_np(x, w[f"conv{i}/kernel"], w[f"conv{i}/bias"], pad=(1, 1, 1, 1))
        x = np.maximum(bn_np(x, w, f"bn{i}"), 0)
        x = maxpool2_np(x)
        levels.append(x)
    return levels


def fcn_8_logits_np(levels, w):
    f3, f4, f5 = levels[2], levels[3], levels[4]
    o = np.maximum(conv2d_np(f5, w["head7/kernel"], w["head7/bias"], pad=same_pad(7)), 0)
    o = np.maximum(conv2d_np(o, w["head1/kernel"], w["head1/bias"]), 0)
    o = conv2d_np(o, w["score5/kernel"], w["score5/bias"])
    o = deconv_np(o, w["up2a/kernel"], 2)
    o2 = conv2d_np(f4, w["score4/kernel"], w["score4/bias"])
    h = min(o.shape[1], o2.shape[1]); ww = min(o.shape[2], o2.shape[2])
    o = o[:, :h, :ww] + o2[:, :h, :ww]
    o = deconv_np(o, w["up2b/kernel"], 2)
    o2 = conv2d_np(f3, w["score3/kernel"], w["score3/bias"])
    h = min(o.shape[1], o2.shape[1]); ww = min(o.shape[2], o2.shape[2])
    o = o2[:, :h, :ww] + o[:, :h, :ww]
    return deconv_np(o, w["up8/kernel"], 8)


def softmax_np(logits_nhwc):
    B, oh, ow, n = logits_nhwc.shape
    z = logits_nhwc.reshape(B, oh * ow, n)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


# ----------------------------------------------------------------------------- MobileNet-v1 / ResNet50 encoders
def relu6_t(x):
    """mobilenet.py:12-13: K.relu(x, max_value=6)."""
    return torch.clamp(x, 0.0, 6.0)


def mobilenet_encoder_t(x, w):
    """mobilenet.py:59-104 (alpha = 1, depth_multiplier = 1).  x NCHW.  Layer names as in the reference:
    conv1 / conv1_bn, conv_dw_%d / conv_dw_%d_bn, conv_pw_%d / conv_pw_%d_bn.  Returns [f1..f5]."""
    # _conv_block (:16-28): ZeroPadding2D((1,1)) (symmetric!) -> Conv 3x3 valid stride 2 no-bias -> BN -> relu6
    x = conv2d_t(x, w["conv1/kernel"], None, stride=2, pad=(1, 1, 1, 1))
    x = relu6_t(bn_t(x, w, "conv1_bn"))
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1), (512, 1),
           (1024, 2), (1024, 1)]
    levels = []
    for i, (f, s) in enumerate(cfg, start=1):
        # _depthwise_conv_block (:31-56): ZeroPad(1,1) -> DepthwiseConv 3x3 valid stride s no-bias -> BN -> relu6
        #                                 -> Conv 1x1 same no-bias -> BN -> relu6
        k = w["conv_dw_%d/depthwise_kernel" % i]                     # [3,3,C,1]
        C = k.shape[2]
        wt = k.permute(2, 3, 0, 1).contiguous()              # [C,1,3,3]
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), wt, None, stride=s, groups=C)
        x = relu6_t(bn_t(x, w, "conv_dw_%d_bn" % i))
        x = conv2d_t(x, w["conv_pw_%d/kernel" % i], None)
        x = relu6_t(bn_t(x, w, "conv_pw_%d_bn" % i))
        if i in (1, 3, 5, 11, 13):
            levels.append(x)
    return levels


def _res_block_t(x, w, stage, block, stride, with_shortcut):
    """resnet50.py:32-119 identity_block / conv_block (all convs have bias; BN after every conv)."""
    cb = "res%d%s_branch" % (stage, block)
    bb = "bn%d%s_branch" % (stage, block)
    y = F.relu(bn_t(conv2d_t(x, w[cb + "2a/kernel"], w[cb + "2a/bias"], stride=stride), w, bb + "2a"))
    y = F.relu(bn_t(conv2d_t(y, w[cb + "2b/kernel"], w[cb + "2b/bias"], pad=same_pad(3)), w, bb + "2b"))
    y = bn_t(conv2d_t(y, w[cb + "2c/kernel"], w[cb + "2c/bias"]), w, bb + "2c")
    sc = x
    if with_shortcut:
        sc = bn_t(conv2d_t(x, w[cb + "1/kernel"], w[cb + "1/bias"], stride=stride), w, bb + "1")
    return F.relu(y + sc)


def resnet50_encoder_t(x, w):
    """resnet50.py:122-173.  Returns [f1 (conv1 output BEFORE BN/ReLU), f2 (one_side_pad), f3, f4, f5]."""
    x = conv2d_t(x, w["conv1/kernel"], w["conv1/bias"], stride=2, pad=(3, 3, 3, 3))
    f1 = x
    x = F.relu(bn_t(x, w, "bn_conv1"))
    x = F.max_pool2d(x, 3, 2)
    levels = [f1]
    for stage, blocks, stride in ((2, "abc", 1), (3, "abcd", 2), (4, "abcdef", 2), (5, "abc", 2)):
        for bi, b in enumerate(blocks):
            x = _res_block_t(x, w, stage, b, stride if bi == 0 else 1, bi == 0)
        if stage == 2:
            levels.append(F.pad(x, (1, 0, 1, 0)))   # one_side_pad (:23-29): pad 1 all round, drop last row/col = pad top/left
        else:
            levels.append(x)
    return levels


def vgg_encoder_t(x, w):
    """vgg16.py:17-81: blocks of 2-2-3-3-3 Conv2D(3x3, 'same', bias, relu), channels 64/128/256/512/512, MaxPool 2x2/2 after
    each block (no BatchNormalization); levels f1..f5 are the pooled block outputs.  x NCHW."""
    levels = []
    for b, (n, _f) in enumerate(((2, 64), (2, 128), (3, 256), (3, 512), (3, 512)), start=1):
        for i in range(1, n + 1):
            name = "block%d_conv%d" % (b, i)
            x = F.relu(conv2d_t(x, w[name + "/kernel"], w[name + "/bias"], pad=same_pad(3)))
        x = F.max_pool2d(x, 2, 2)
        levels.append(x)
    return levels


def fcn_forward_encoder(x_nhwc, weights, encoder="mobilenet", dtype=torch.float64, return_levels=False):
    """fcn_8 over the MobileNet / ResNet50 / VGG16 encoder: float NHWC -> probs [B, oh*ow, n]."""
    w = _prep(weights, dtype)
    x = _t(x_nhwc, dtype).permute(0, 3, 1, 2)
    levels = {"mobilenet": mobilenet_encoder_t, "resnet50": resnet50_encoder_t, "vgg": vgg_encoder_t}[encoder](x, w)
    probs = segmentation_probs_t(fcn_8_logits_t(levels, w)).numpy()
    if return_levels:
        return probs, [l.permute(0, 2, 3, 1).contiguous().numpy() for l in levels]
    return probs


# ----------------------------------------------------------------------------- numpy restatement of the other encoders
def conv2d_np_s(x, kernel, bias=None, pad=(0, 0, 0, 0), stride=1):
    """conv2d_np with a stride: the 'valid' conv evaluated at every position, then sub-sampled (Keras strides start at 0)."""
    y = conv2d_np(x, kernel, None, pad)[:, ::stride, ::stride]
    return y + bias if bias is not None else y


def dwconv_np(x, kernel, pad=(0, 0, 0, 0), stride=1):
    """DepthwiseConv2D 'valid', depth_multiplier 1: x NHWC, kernel [kh,kw,C,1]; explicit loop over the taps."""
    x = np.pad(x, ((0, 0), (pad[0], pad[1]), (pad[2], pad[3]), (0, 0)))
    kh, kw = kernel.shape[:2]
    oh, ow = x.shape[1] - kh + 1, x.shape[2] - kw + 1
    out = np.zeros((x.shape[0], oh, ow, x.shape[3]), dtype=x.dtype)
    for a in range(kh):
        for b in range(kw):
            out += x[:, a: a + oh, b: b + ow, :] * kernel[a, b, :, 0]
    return out[:, ::stride, ::stride]


def maxpool_np(x, k, s):
    """MaxPooling2D((k,k), strides=s) 'valid' (floor)."""
    B, H, W, C = x.shape
    oh, ow = (H - k) // s + 1, (W - k) // s + 1
    out = np.full((B, oh, ow, C), -np.inf, dtype=x.dtype)
    for a in range(k):
        for b in range(k):
            out = np.maximum(out, x[:, a: a + (oh - 1) * s + 1: s, b: b + (ow - 1) * s + 1: s])
    return out


def vgg_encoder_np(x, w):
    """vgg16.py:27-74 (x NHWC)."""
    levels = []
    for b, n in enumerate((2, 2, 3, 3, 3), start=1):
        for i in range(1, n + 1):
            name = "block%d_conv%d" % (b, i)
            x = np.maximum(conv2d_np(x, w[name + "/kernel"], w[name + "/bias"], pad=same_pad(3)), 0)
        x = maxpool_np(x, 2, 2)
        levels.append(x)
    return levels


def mobilenet_encoder_np(x, w):
    """mobilenet.py:16-104 (x NHWC)."""
    r6 = lambda t: np.minimum(np.maximum(t, 0), 6)
    x = r6(bn_np(conv2d_np_s(x, w["conv1/kernel"], None, pad=(1, 1, 1, 1), stride=2), w, "conv1_bn"))
    strides = (1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1)
    levels = []
    for i, s in enumerate(strides, start=1):
        x = r6(bn_np(dwconv_np(x, w["conv_dw_%d/depthwise_kernel" % i], pad=(1, 1, 1, 1), stride=s), w, "conv_dw_%d_bn" % i))
        x = r6(bn_np(conv2d_np(x, w["conv_pw_%d/kernel" % i]), w, "conv_pw_%d_bn" % i))
        if i in (1, 3, 5, 11, 13):
            levels.append(x)
    return levels


def resnet50_encoder_np(x, w):
    """resnet50.py:23-173 (x NHWC): f1 is conv1's raw output, f2 the top/left-padded stage-2 output."""
    relu = lambda t: np.maximum(t, 0)

    def block(x, stage, blk, stride, shortcut):
        cb, bb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        y = relu(bn_np(conv2d_np_s(x, w[cb + "2a/kernel"], w[cb + "2a/bias"], stride=stride), w, bb + "2a"))
        y = relu(bn_np(conv2d_np(y, w[cb + "2b/kernel"], w[cb + "2b/bias"], pad=same_pad(3)), w, bb + "2b"))
        y = bn_np(conv2d_np(y, w[cb + "2c/kernel"], w[cb + "2c/bias"]), w, bb + "2c")
        sc = bn_np(conv2d_np_s(x, w[cb + "1/kernel"], w[cb + "1/bias"], stride=stride), w, bb + "1") if shortcut else x
        return relu(y + sc)

    x = conv2d_np_s(x, w["conv1/kernel"], w["conv1/bias"], pad=(3, 3, 3, 3), stride=2)
    levels = [x]
    x = maxpool_np(relu(bn_np(x, w, "bn_conv1")), 3, 2)
    for stage, blocks, stride in ((2, "abc", 1), (3, "abcd", 2), (4, "abcdef", 2), (5, "abc", 2)):
        for bi, b in enumerate(blocks):
            x = block(x, stage, b, stride if bi == 0 else 1, bi == 0)
        levels.append(np.pad(x, ((0, 0), (1, 0), (1, 0), (0, 0))) if stage == 2 else x)
    return levels


def fcn_32_logits_np(levels, w):
    """fcn.py:137-146 (numpy)."""
    f5 = levels[4]
    o = np.maximum(conv2d_np(f5, w["head7/kernel"], w["head7/bias"], pad=same_pad(7)), 0)
    o = np.maximum(conv2d_np(o, w["head1/kernel"], w["head1/bias"]), 0)
    o = conv2d_np(o, w["score5/kernel"], w["score5/bias"])
    return deconv_np(o, w["up32/kernel"], 32)
