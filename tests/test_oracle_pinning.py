"""CPU tests: the oracle restatements against (1) the committed golden vectors minted from the reference's own
functions + OpenCV 4.13 (tests/make_golden.py), (2) cv2 directly, (3) each other (two CNN restatements)."""
import hashlib

import cv2
import numpy as np
import pytest

import golden_inputs as gi
from oracle import align, cnn, decode, preprocess


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_opencv_version_matches_golden(golden):
    assert str(golden["opencv_version"]) == cv2.__version__ == "4.13.0"


@pytest.mark.parametrize("i", range(len(gi.DETECT_CASES)))
def test_detect_marks_pre_and_post(golden, i):
    seed, h, w, face = gi.DETECT_CASES[i]
    img = gi.image(seed, h, w)
    rgb, fb = preprocess.crop_resize_rgb(img, face)
    assert (sha(rgb[None]) == golden["detect_input_sha_%d" % i]).all()
    if i < 2:
        assert (rgb[None] == golden["detect_input_%d" % i]).all()
    _, marks_u = decode.regression_decode(gi.fake_outputs(seed)[0], fb)
    assert marks_u.dtype == np.uint64
    assert (marks_u == golden["detect_marks_%d" % i]).all()


def test_square_box_python_semantics():
    assert preprocess.square_box([200, 120, 400, 360]) == [180, 144, 420, 384]
    # odd positive diff -> extra pixel on the right; odd negative diff -> extra pixel at the bottom
    b = preprocess.square_box([100, 100, 301, 333])
    assert b[2] - b[0] == b[3] - b[1]
    b = preprocess.square_box([50, 60, 351, 260])
    assert b[2] - b[0] == b[3] - b[1] == 301


def test_get_image_array(golden):
    img = gi.image(21, 45, 60)
    for norm in ("sub_mean", "sub_and_divide", "divide"):
        a = preprocess.get_image_array(img, 48, 32, imgNorm=norm)
        g = golden["image_array_" + norm]
        assert a.dtype == np.float32 and a.shape == g.shape
        assert np.array_equal(np.ascontiguousarray(a), g), norm
    a = preprocess.get_image_array(img, 48, 32, ordering="channels_first")
    assert np.array_equal(np.ascontiguousarray(a), golden["image_array_cf"])


def test_class_map(golden):
    for i, (oh, ow, n) in enumerate([(12, 12, 5), (9, 14, 68)]):
        p = gi.probs(31 + i, oh * ow, n)
        assert np.array_equal(decode.class_map(p[0], oh, ow, n), golden["class_map_%d" % i])


def test_average_xy_family(golden):
    hm = gi.heatmaps(41, 2, 24, 20, 3)
    for ci, (npnt, th) in enumerate(golden["average_xy_cases"]):
        for b in range(2):
            for l in range(3):
                got = decode.average_xy(hm[b, :, :, l], int(npnt), float(th))
                np.testing.assert_allclose(got, golden["average_xy"][ci, b, l], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(decode.transfer_xy_coord(hm[0], 9, 0.5), golden["transfer_xy_coord"], rtol=1e-6)
    np.testing.assert_allclose(decode.transfer_target(hm, 0.4, 16), golden["transfer_target"], rtol=1e-6)
    # the positional-argument slip: whatever is passed, the reference decodes with n_points=4, thresh=0
    np.testing.assert_allclose(decode.transfer_xy_coord(hm[0], 64, 0.2), decode.transfer_xy_coord(hm[0], 9, 0.5))


def test_resize_bit_exact(golden):
    for i, (sh, sw, dh, dw) in enumerate(gi.RESIZE_SHAPES):
        img = gi.image(50 + i, sh, sw)
        got = preprocess.resize_linear_u8(img, dw, dh)
        assert (sha(got) == golden["resize_sha_%d" % i]).all(), (sh, sw, dh, dw)
        assert np.array_equal(got, cv2.resize(img, (dw, dh)))


def test_resize_bit_exact_random_shapes():
    rng = np.random.default_rng(7)
    for _ in range(25):
        sh, sw = int(rng.integers(8, 500)), int(rng.integers(8, 500))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(preprocess.resize_linear_u8(img, 128, 128), cv2.resize(img, (128, 128)))


def test_warp_bit_exact(golden):
    frame = gi.image(60, 1080, 1920)
    for i in range(6):
        M = gi.similarity(i)
        got = align.warp_affine_u8(frame, M, 112, 112)
        assert (sha(got) == golden["warp_sha_%d" % i]).all()
        if i == 0:
            assert np.array_equal(got, golden["warp_crop_0"])


def test_umeyama_closed_form_vs_svd_and_recovery():
    rng = np.random.default_rng(3)
    for t in range(300):
        n = 5 if t % 2 else 68
        p = rng.normal(0, 60, (n, 2)) + rng.uniform(0, 1500, 2)
        s, th = rng.uniform(0.2, 2), rng.uniform(-np.pi, np.pi)
        L = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        tv = rng.uniform(-80, 80, 2)
        q = p @ L.T + tv
        M = align.umeyama(p, q)
        np.testing.assert_allclose(M, np.concatenate([L, tv[:, None]], 1), rtol=1e-9, atol=1e-8)   # exact recovery
        qn = q + rng.normal(0, 1, q.shape)
        np.testing.assert_allclose(align.umeyama(p, qn), align.umeyama_svd(p, qn), rtol=1e-9, atol=1e-9)
    assert np.isnan(align.umeyama(np.ones((5, 2)), np.ones((5, 2)))).all()                          # degenerate -> NaN


def _small_weights(rng, n_classes=5):
    w = {}
    cin = 3
    for i, f in enumerate((8, 8, 16, 16, 16), start=1):
        w[f"conv{i}/kernel"] = rng.normal(0, 0.3, (3, 3, cin, f))
        w[f"conv{i}/bias"] = rng.normal(0, 0.1, f)
        w[f"bn{i}/gamma"] = rng.uniform(0.5, 1.5, f); w[f"bn{i}/beta"] = rng.normal(0, 0.1, f)
        w[f"bn{i}/moving_mean"] = rng.normal(0, 0.1, f); w[f"bn{i}/moving_variance"] = rng.uniform(0.5, 1.5, f)
        cin = f
    w["head7/kernel"] = rng.normal(0, 0.05, (7, 7, 16, 32)); w["head7/bias"] = rng.normal(0, 0.1, 32)
    w["head1/kernel"] = rng.normal(0, 0.2, (1, 1, 32, 32)); w["head1/bias"] = rng.normal(0, 0.1, 32)
    for s, c in (("score5", 32), ("score4", 16), ("score3", 16)):
        w[s + "/kernel"] = rng.normal(0, 0.2, (1, 1, c, n_classes)); w[s + "/bias"] = rng.normal(0, 0.1, n_classes)
    w["up2a/kernel"] = rng.normal(0, 0.2, (4, 4, n_classes, n_classes))
    w["up2b/kernel"] = rng.normal(0, 0.2, (4, 4, n_classes, n_classes))
    w["up8/kernel"] = rng.normal(0, 0.1, (16, 16, n_classes, n_classes))
    return w


def test_cnn_two_restatements_agree():
    """torch-functional vs numpy-einsum restatement of vanilla_encoder + fcn_8 + softmax (fcn.py:10-126)."""
    import torch
    rng = np.random.default_rng(5)
    w = _small_weights(rng)
    x = rng.normal(0, 1, (2, 64, 96, 3))
    probs, logits = cnn.fcn_forward(x, w, "fcn_8", torch.float64, return_logits=True)
    levels = cnn.vanilla_encoder_np(x, w)
    logits_np = cnn.fcn_8_logits_np(levels, w)
    assert logits.shape == logits_np.shape == (2, 8 * 8 + 8, 8 * 12 + 8, 5)     # oh = 8*(H/8)+8
    np.testing.assert_allclose(logits, logits_np, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(probs, cnn.softmax_np(logits_np), rtol=1e-9, atol=1e-12)
    lv_t = cnn.trunk_forward(x, w, torch.float64)
    for a, b in zip(lv_t, levels):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", ["vgg", "mobilenet", "resnet50", "fcn_32"])
def test_cnn_two_restatements_agree_other_encoders(name):
    """The second, torch-free restatement of the graphs the first cross-check does not reach: depth-wise and
    strided convs, ReLU6, residual blocks with strided 1x1 shortcuts, the 3x3/2 pool, one_side_pad, the 64x64 / stride-32
    transposed conv.  reference networks/vgg16.py:27-74, mobilenet.py:16-104, resnet50.py:23-173, fcn.py:129-150.  Weights come
    from the builders' own seeded initialiser (non-trivial BN statistics)."""
    import torch
    from keypoints_detector.networks import fcn, mobilenet, resnet50, vgg16
    from keypoints_detector.networks.model import Model
    rng = np.random.default_rng(11)
    H, W = (32, 64) if name != "fcn_32" else (32, 32)
    x = rng.normal(0, 1, (2, H, W, 3))
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    if name == "fcn_32":
        m = fcn.fcn_32(5, input_height=H, input_width=W).init_weights(3)
        wn = {k: v.astype(np.float64) for k, v in m.weights.items()}
        wt = cnn._prep(m.weights, torch.float64)
        lv_t = cnn.vanilla_encoder_t(xt, wt)
        lo_t = cnn.fcn_32_logits_t(lv_t, wt).permute(0, 2, 3, 1).numpy()
        lo_n = cnn.fcn_32_logits_np(cnn.vanilla_encoder_np(x, wn), wn)
        assert lo_t.shape == lo_n.shape == (2, 64, 64, 5)                             # 32 * h5 + 32
        np.testing.assert_allclose(lo_t, lo_n, rtol=1e-9, atol=1e-9)
        return
    getter, enc_t, enc_n = {"vgg": (vgg16.get_vgg_encoder, cnn.vgg_encoder_t, cnn.vgg_encoder_np),
                            "mobilenet": (mobilenet.get_mobilenet_encoder, cnn.mobilenet_encoder_t, cnn.mobilenet_encoder_np),
                            "resnet50": (resnet50.get_resnet50_encoder, cnn.resnet50_encoder_t, cnn.resnet50_encoder_np)}[name]
    g, _ = getter(H, W)
    wf = Model(g, "segmentation").init_weights(0).weights
    wn = {k: v.astype(np.float64) for k, v in wf.items()}
    lv_t = enc_t(xt, cnn._prep(wf, torch.float64))
    lv_n = enc_n(x, wn)
    assert len(lv_t) == len(lv_n) == 5
    for i, (a, b) in enumerate(zip(lv_t, lv_n)):
        a = a.permute(0, 2, 3, 1).numpy()
        assert a.shape == b.shape, (i, a.shape, b.shape)
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(a).max()), err_msg="level %d" % (i + 1))


def test_keras_semantics_spot_checks():
    import torch
    # fresh BatchNormalization is NOT the identity: y = x / sqrt(1 + 1e-3)
    w = {"bn/gamma": torch.ones(2, dtype=torch.float64), "bn/beta": torch.zeros(2, dtype=torch.float64),
         "bn/moving_mean": torch.zeros(2, dtype=torch.float64), "bn/moving_variance": torch.ones(2, dtype=torch.float64)}
    x = torch.ones(1, 2, 1, 1, dtype=torch.float64)
    assert abs(float(cnn.bn_t(x, w, "bn")[0, 0, 0, 0]) - 1 / np.sqrt(1.001)) < 1e-15
    # Conv2DTranspose valid: out = (in-1)*s + k, no kernel flip
    k = np.zeros((4, 4, 1, 1)); k[0, 1, 0, 0] = 1.0
    o = cnn.deconv_np(np.ones((1, 2, 2, 1)), k, 2)
    assert o.shape == (1, 6, 6, 1) and o[0, 0, 1, 0] == 1 and o[0, 2, 3, 0] == 1 and o[0, 1, 1, 0] == 0
    # 'same' with a 7x7 kernel pads 3/3
    assert cnn.same_pad(7) == (3, 3, 3, 3) and cnn.same_pad(4) == (1, 2, 1, 2)
