"""GPU parity tests (pytest -m gpu) for the MobileNet-v1 / ResNet50 encoders behind fcn_8 / fcn_32
(SURVEY §8 rows a5 / f3): depth-wise convs, strided convs, 3x3/2 max-pool, residual add + ReLU, checked layer by
layer against torch fp32 references and end to end against the fp64 oracle restatement (oracle/cnn.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_lib):
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).contiguous()


def rel(a, b):
    return float(np.abs(a - b).mean() / (np.abs(b).mean() + 1e-12))


# ------------------------------------------------------------------------------------------------ single layers
@pytest.mark.parametrize("dtype,tol", [("float32", 2e-6), ("bfloat16", 8e-3)])
def test_dwconv_maxpool_add_layers(dev, dtype, tol):
    """One graph exercising DWCONV (stride 1 and 2, odd map sizes), MAXPOOL 3/2 and ADD+ReLU with both dtypes flowing
    between them; every intermediate tensor is compared with a torch fp32 evaluation of the same folded graph."""
    from keypoints_detector import _native as N
    from keypoints_detector.networks.model import Graph, Model, BN_EPS
    g = Graph(37, 29, 3)
    c0 = g.conv(0, "stem", 64, 3, pad=(1, 1, 1, 1), stride=2, act=N.ACT_RELU6, bias=False, bn=True, bn_name="stem_bn")
    d1 = g.dwconv(c0, "dw1", 3, stride=1, act=N.ACT_RELU6)
    p1 = g.conv(d1, "pw1", 64, 1, act=N.ACT_NONE, bias=True, bn=True, bn_name="pw1_bn")
    a1 = g.add(p1, c0, "res1", act=N.ACT_RELU)
    mp = g.maxpool(a1, "pool", 3, 2)
    d2 = g.dwconv(mp, "dw2", 3, stride=2, act=N.ACT_RELU6)
    p2 = g.conv(d2, "pw2", 128, 1, act=N.ACT_RELU6, bias=False, bn=True, bn_name="pw2_bn")
    s2 = g.conv(p2, "strided", 64, 1, stride=2, act=N.ACT_RELU, bias=True, bn=True, bn_name="strided_bn")
    out = g.dense(s2, "fc", 10)
    m = Model(g, "regression", "layers").init_weights(11)
    w = {k: torch.from_numpy(v).double() for k, v in m.weights.items()}
    x = np.random.default_rng(3).uniform(0, 1, (5, 37, 29, 3)).astype(np.float32)

    def bn(t, name):
        s = w[name + "/gamma"] / torch.sqrt(w[name + "/moving_variance"] + BN_EPS)
        return (t - w[name + "/moving_mean"].view(1, -1, 1, 1)) * s.view(1, -1, 1, 1) + w[name + "/beta"].view(1, -1, 1, 1)

    def conv(t, name, stride=1, pad=0, bias=False):
        k = w[name + "/kernel"].permute(3, 2, 0, 1)
        return F.conv2d(F.pad(t, (pad,) * 4), k, w[name + "/bias"] if bias else None, stride=stride)

    def dw(t, name, stride):
        k = w[name + "/depthwise_kernel"].permute(2, 3, 0, 1)
        return F.conv2d(F.pad(t, (1, 1, 1, 1)), k, None, stride=stride, groups=k.shape[0])

    t = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    r = {}
    r[c0] = torch.clamp(bn(conv(t, "stem", 2, 1), "stem_bn"), 0, 6)
    r[d1] = torch.clamp(bn(dw(r[c0], "dw1", 1), "dw1_bn"), 0, 6)
    r[p1] = bn(conv(r[d1], "pw1", bias=True), "pw1_bn")
    r[a1] = F.relu(r[p1] + r[c0])
    r[mp] = F.max_pool2d(r[a1], 3, 2)
    r[d2] = torch.clamp(bn(dw(r[mp], "dw2", 2), "dw2_bn"), 0, 6)
    r[p2] = torch.clamp(bn(conv(r[d2], "pw2"), "pw2_bn"), 0, 6)
    r[s2] = F.relu(bn(conv(r[p2], "strided", 2, 0, True), "strided_bn"))
    xt = T(x, dev)
    for tid in (c0, d1, p1, a1, mp, d2, p2, s2):
        got = m.intermediate(xt, tid, dtype).cpu().numpy()
        ref = r[tid].permute(0, 2, 3, 1).numpy()
        assert got.shape == ref.shape, (tid, got.shape, ref.shape)
        assert rel(got, ref) < tol, (tid, rel(got, ref))
    ref_out = r[s2].permute(0, 2, 3, 1).reshape(5, -1) @ w["fc/kernel"] + w["fc/bias"]
    got_out = m.forward_device(xt, dtype).cpu().numpy()
    assert rel(got_out, ref_out.numpy()) < tol * 2


# ------------------------------------------------------------------------------------------------ full encoders
def _encoder_case(dev, name, dtype, H, W, B=3, seed=21):
    from keypoints_detector.networks import fcn, mobilenet, resnet50
    from oracle import cnn as o_cnn
    if name == "mobilenet":
        m = fcn.fcn_8_mobilenet(68, H, W).init_weights(seed)
        _, levels = mobilenet.get_mobilenet_encoder(H, W)
    elif name == "vgg":
        from keypoints_detector.networks import vgg16
        m = fcn.fcn_8_vgg(68, H, W).init_weights(seed)
        _, levels = vgg16.get_vgg_encoder(H, W)
    else:
        m = fcn.fcn_8_resnet50(68, H, W).init_weights(seed)
        _, levels = resnet50.get_resnet50_encoder(H, W)
    x = np.random.default_rng(seed).uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    probs_ref, lv_ref = o_cnn.fcn_forward_encoder(x.astype(np.float64), m.weights, name, torch.float64, return_levels=True)
    xt = T(x, dev)
    probs = m.forward_device(xt, dtype).cpu().numpy()
    lv = [m.intermediate(xt, levels[i], dtype).cpu().numpy() for i in (2, 3, 4)]
    return m, probs, probs_ref, lv, [lv_ref[i] for i in (2, 3, 4)]


@pytest.mark.parametrize("name,H,W", [("mobilenet", 64, 96), ("resnet50", 96, 64), ("vgg", 64, 96)])
def test_fcn8_encoder_fp32(dev, name, H, W):
    m, probs, probs_ref, lv, lv_ref = _encoder_case(dev, name, "float32", H, W)
    assert m.model_name == "fcn_8_" + name
    for got, ref in zip(lv, lv_ref):
        assert got.shape == ref.shape
        assert rel(got, ref) < 1e-5, rel(got, ref)
    assert probs.shape == probs_ref.shape
    assert np.abs(probs - probs_ref).max() < 1e-4
    oh, ow = m.output_height, m.output_width
    ref = probs_ref.reshape(-1, oh, ow, 68).argmax(-1)
    assert (probs.reshape(-1, oh, ow, 68).argmax(-1) == ref).mean() > 0.999


@pytest.mark.parametrize("name,H,W", [("mobilenet", 64, 96), ("resnet50", 96, 64), ("vgg", 64, 96)])
def test_fcn8_encoder_bf16(dev, name, H, W):
    """bf16 operands / fp32 accumulation through 27 (MobileNet) / 53 (ResNet50) stacked layers on random-init weights:
    the levels stay within a few percent of the fp64 oracle, the class map agrees on the bulk of the pixels."""
    m, probs, probs_ref, lv, lv_ref = _encoder_case(dev, name, "bfloat16", H, W)
    for got, ref in zip(lv, lv_ref):
        assert rel(got, ref) < 0.05, rel(got, ref)
    assert np.abs(probs - probs_ref).mean() < 2e-3
    oh, ow = m.output_height, m.output_width
    ref = probs_ref.reshape(-1, oh, ow, 68).argmax(-1)
    assert (probs.reshape(-1, oh, ow, 68).argmax(-1) == ref).mean() > 0.9


def test_fcn32_mobilenet_and_registry(dev):
    """LANDMARKS_MODELS builds every reference entry (basic_models.py) and fcn_32 over an encoder with f5 only."""
    from keypoints_detector.networks.basic_models import LANDMARKS_MODELS
    from oracle import cnn as o_cnn
    for k in ("fcn_8_mobilenet", "fcn_8_resnet50", "fcn_32_mobilenet", "fcn_32_resnet50"):
        assert k in LANDMARKS_MODELS
    m = LANDMARKS_MODELS["fcn_32_mobilenet"](68, input_height=64, input_width=64).init_weights(4)
    x = np.random.default_rng(4).uniform(0, 1, (2, 64, 64, 3)).astype(np.float32)
    w = o_cnn._prep(m.weights, torch.float64)
    lv = o_cnn.mobilenet_encoder_t(torch.from_numpy(x).double().permute(0, 3, 1, 2), w)
    ref = o_cnn.segmentation_probs_t(o_cnn.fcn_32_logits_t(lv, w)).numpy()
    probs = m.forward_device(T(x, dev), "float32").cpu().numpy()
    assert probs.shape == ref.shape
    assert np.abs(probs - ref).max() < 1e-4
    # bf16: the 64x64 / stride-32 transposed conv runs as a tensor-core GEMM with 1024 phases
    probs16 = m.forward_device(T(x, dev), "bfloat16").cpu().numpy()
    assert np.abs(probs16 - ref).mean() < 2e-3
    oh, ow = m.output_height, m.output_width
    assert (probs16.reshape(2, oh, ow, 68).argmax(-1) == ref.reshape(2, oh, ow, 68).argmax(-1)).mean() > 0.9
