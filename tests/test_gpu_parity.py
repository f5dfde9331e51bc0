"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C-ABI (ctypes shim of the drop-in
package), against the CPU oracle, the committed golden vectors minted from the reference's own functions,
and OpenCV 4.13.  Bars (BASELINE.json north_star): integer / byte / index work bit-exact; landmarks within
0.05 px (fp32 mode) / 0.5 px (bf16 mode); aligned crops bit-exact given the same matrix."""
import hashlib

import cv2
import numpy as np
import pytest
import torch

import golden_inputs as gi

pytestmark = pytest.mark.gpu


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def dev(cuda_lib):
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def T(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev).contiguous()


# ------------------------------------------------------------------------------------------------ a1 pre-processing
def test_preprocess_faces_golden_and_oracle(dev, golden):
    from keypoints_detector import prediction
    from oracle import preprocess as o_pre
    for i, (seed, h, w, face) in enumerate(gi.DETECT_CASES):
        img = gi.image(seed, h, w)
        crops, fb = prediction.preprocess_faces_device(T(img[None], dev), T(np.array([face], np.int32), dev),
                                                       T(np.zeros(1, np.int32), dev))
        ref, rfb = o_pre.crop_resize_rgb(img, face)
        assert fb.cpu().numpy()[0].tolist() == rfb
        assert np.array_equal(crops.cpu().numpy()[0], ref), "case %d" % i
        assert (sha(crops.cpu().numpy()) == golden["detect_input_sha_%d" % i]).all()      # what the reference fed its model


def test_preprocess_faces_batched_random_boxes(dev):
    from keypoints_detector import prediction
    from oracle import preprocess as o_pre
    rng = np.random.default_rng(5)
    frames = np.stack([gi.image(70 + i, 360, 500) for i in range(3)])
    boxes, f2f = [], []
    for i in range(96):
        w, h = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        x0, y0 = int(rng.integers(0, 480)), int(rng.integers(0, 330))        # many leave the frame bottom/right
        boxes.append([x0, y0, x0 + w, y0 + h]); f2f.append(i % 3)
    boxes.append([100, 100, 356, 356]); f2f.append(0)                          # exact 2x (area shortcut)
    boxes.append([50, 40, 178, 168]); f2f.append(1)                            # same size (copy shortcut)
    boxes = np.array(boxes, np.int32); f2f = np.array(f2f, np.int32)
    crops, fb = prediction.preprocess_faces_device(T(frames, dev), T(boxes, dev), T(f2f, dev))
    crops, fb = crops.cpu().numpy(), fb.cpu().numpy()
    n_checked = 0
    for i in range(len(boxes)):
        sq = o_pre.square_box(boxes[i])
        assert fb[i].tolist() == sq
        if sq[0] < 0 or sq[1] < 0 or sq[0] >= 500 or sq[1] >= 360:
            continue                                                           # reference undefined (crashes)
        ref, _ = o_pre.crop_resize_rgb(frames[f2f[i]], boxes[i])
        assert np.array_equal(crops[i], ref), (i, boxes[i])
        n_checked += 1
    assert n_checked > 40


def test_image_array_golden(dev, golden):
    from keypoints_detector.data.generator import get_image_array
    img = gi.image(21, 45, 60)
    for norm in ("sub_mean", "sub_and_divide", "divide"):
        a = get_image_array(img, 48, 32, imgNorm=norm, ordering="channels_last")
        assert a.dtype == np.float32 and np.array_equal(a, golden["image_array_" + norm]), norm
    assert np.array_equal(get_image_array(img, 48, 32), golden["image_array_cf"])                # default channels_first
    big = gi.image(22, 480, 640)
    from oracle import preprocess as o_pre
    assert np.array_equal(get_image_array(big, 224, 224, ordering="channels_last"), np.ascontiguousarray(o_pre.get_image_array(big, 224, 224)))


# ------------------------------------------------------------------------------------------------ a3 / a7 / a8 decodes
def test_decode_regress_golden(dev, golden):
    from keypoints_detector import prediction
    from oracle import decode as o_dec, preprocess as o_pre
    outs = np.concatenate([gi.fake_outputs(c[0]) for c in gi.DETECT_CASES])
    fbs = np.array([o_pre.square_box(c[3]) for c in gi.DETECT_CASES], np.int32)
    marks, marks_u = prediction.decode_regress_device(T(outs, dev), T(fbs, dev))
    marks, marks_u = marks.cpu().numpy(), marks_u.cpu().numpy()
    for i in range(len(gi.DETECT_CASES)):
        mf, mu = o_dec.regression_decode(outs[i], fbs[i])
        assert np.array_equal(marks[i], mf)                                    # pre-cast floats bit-exact
        assert np.array_equal(marks_u[i].astype(np.uint64), golden["detect_marks_%d" % i])


def test_classmap_golden_and_large(dev, golden):
    from keypoints_detector import prediction
    for i, (oh, ow, n) in enumerate([(12, 12, 5), (9, 14, 68)]):
        p = gi.probs(31 + i, oh * ow, n)
        cm = prediction.class_map_device(T(p, dev), oh, ow).cpu().numpy()[0]
        assert cm.dtype == np.int64 and np.array_equal(cm, golden["class_map_%d" % i])
    rng = np.random.default_rng(9)
    p = rng.normal(0, 1, (3, 232 * 232, 68)).astype(np.float32)
    p[0, 5, 7] = p[0, 5, 40] = 9.0                                             # tie: first maximum wins
    cm = prediction.class_map_device(T(p, dev), 232, 232).cpu().numpy()
    assert np.array_equal(cm, p.reshape(3, 232, 232, 68).argmax(-1)) and cm[0, 0, 5] == 7


def test_heatmap_xy_golden_and_oracle(dev, golden):
    from keypoints_detector.utils import metrics
    from oracle import decode as o_dec
    hm = gi.heatmaps(41, 2, 24, 20, 3)
    for ci, (npnt, th) in enumerate(golden["average_xy_cases"]):
        xy = metrics.heatmap_xy_device(T(hm, dev), int(npnt), float(th)).cpu().numpy().reshape(2, 3, 2)
        np.testing.assert_allclose(xy, golden["average_xy"][ci], rtol=1e-6, atol=1e-9)
    # the reference's wrappers (positional-argument slip reproduced)
    np.testing.assert_allclose(metrics.transfer_xy_coord(hm[0], n_points=9, thresh=0.5), golden["transfer_xy_coord"], rtol=1e-6)
    np.testing.assert_allclose(metrics.transfer_target(hm, thresh=0.4, n_points=16), golden["transfer_target"], rtol=1e-6)
    got = metrics.get_average_xy(hm[1, :, :, 2], 24, 20, 4, 0)
    np.testing.assert_allclose(got, golden["average_xy"][0, 1, 2], rtol=1e-6)
    assert metrics.get_average_xy(np.zeros((8, 8), np.float32), 8, 8, 4, 0.5) == [-1, -1]
    # larger maps, more channels and top-n sizes, all slabs / lanes exercised; top-n path agrees to fp64 rounding
    big = gi.heatmaps(42, 3, 136, 136, 68)
    for n in (4, 1, 9, 40):
        xy = metrics.heatmap_xy_device(T(big, dev), n, 0.0).cpu().numpy()
        ref = o_dec.transfer_target(big, 0.0, n, reproduce_slip=False)
        np.testing.assert_allclose(xy, ref, rtol=1e-12, atol=1e-12)
    xy = metrics.heatmap_xy_device(T(big, dev), 0, 0.0).cpu().numpy()
    np.testing.assert_allclose(xy, o_dec.transfer_target(big, 0.0, 0, reproduce_slip=False), rtol=1e-5)
    # tie-heavy maps (saturated softmax outputs are mostly exact zeros / ones): the higher flat index wins among equal values
    ties = np.round(gi.heatmaps(43, 2, 72, 104, 68) * 3).astype(np.float32) / 3
    ties[0, :, :, :20] = 0.0
    ties[1, 10:40, :, 20:30] = 1.0
    for n in (4, 2, 9):
        xy = metrics.heatmap_xy_device(T(ties, dev), n, -1.0).cpu().numpy()
        ref = o_dec.transfer_target(ties, -1.0, n, reproduce_slip=False)
        np.testing.assert_allclose(xy, ref, rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------------ a10 alignment
def test_warp_affine_golden(dev, golden):
    from keypoints_detector import prediction
    frame = gi.image(60, 1080, 1920)
    Ms = np.stack([gi.similarity(i) for i in range(6)])
    crops = prediction.warp_affine_device(T(frame[None], dev), T(np.zeros(6, np.int32), dev), T(Ms, dev)).cpu().numpy()
    for i in range(6):
        assert (sha(crops[i]) == golden["warp_sha_%d" % i]).all(), i
    assert np.array_equal(crops[0], golden["warp_crop_0"])


def test_align_c4_full_size(dev):
    """config C4: 4096 faces from 64 1080p frames, 5-point fit + warp to 112x112."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    from oracle import align as o_al
    F, B = 64, 4096
    g = torch.Generator().manual_seed(1)
    frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, generator=g)
    pts, _ = synthetic.make_similarity_landmarks(B, 1080, 1920, o_al.TEMPLATE_112, seed=4)
    f2f = (np.arange(B) // 64).astype(np.int32)
    crops, M = prediction.align_device(frames.to(dev), T(f2f, dev), T(pts, dev), None, (112, 112), five_point=False)
    crops, M = crops.cpu().numpy(), M.cpu().numpy()
    fr = frames.numpy()
    for i in range(0, B, 7):
        Mo = o_al.umeyama(pts[i], o_al.TEMPLATE_112)
        np.testing.assert_allclose(M[i], Mo, rtol=1e-9, atol=1e-9)
        assert np.array_equal(M[i], Mo), "fit is expected to be bit-identical (same summation order, no FMA)"
        ref = cv2.warpAffine(fr[f2f[i]], M[i], (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(crops[i], ref), i
    for i in range(0, B, 512):
        assert np.array_equal(crops[i], o_al.warp_affine_u8(fr[f2f[i]], M[i], 112, 112))


def test_align_68_to_5_and_edge_cases(dev):
    from keypoints_detector import prediction
    from keypoints_detector.networks.init import canonical_face68
    from oracle import align as o_al
    rng = np.random.default_rng(8)
    frames = np.stack([gi.image(80 + i, 300, 420) for i in range(2)])
    B = 24
    marks = np.zeros((B, 68, 2), np.float32)
    for i in range(B):
        s = rng.uniform(60, 400)
        marks[i] = (canonical_face68() * s + rng.uniform(-80, 300, 2) + rng.normal(0, 1.5, (68, 2))).astype(np.float32)
    marks[5] = 7.0                                                           # degenerate: all points coincide
    f2f = (np.arange(B) % 2).astype(np.int32)
    crops, M = prediction.align_device(T(frames, dev), T(f2f, dev), T(marks, dev))
    crops, M = crops.cpu().numpy(), M.cpu().numpy()
    Mo, co = o_al.align_faces(frames, f2f, marks)
    assert np.isnan(M[5]).all() and np.isnan(Mo[5]).all() and not crops[5].any()
    ok = [i for i in range(B) if i != 5]
    assert np.array_equal(M[ok], Mo[ok])
    assert np.array_equal(crops[ok], co[ok])                                  # faces partly / fully outside the frame included
    # generic path: odd output width, single channel, caller template
    gray = np.ascontiguousarray(frames[..., :1])
    tm = o_al.TEMPLATE_112 * (96 / 112.0)
    p5 = o_al.five_points(marks[ok]).astype(np.float32)
    c2, M2 = prediction.align_device(T(gray, dev), T(f2f[ok], dev), T(p5, dev), tm, (90, 97), five_point=False)
    c2, M2 = c2.cpu().numpy(), M2.cpu().numpy()
    for j, i in enumerate(ok):
        ref = cv2.warpAffine(gray[f2f[i]], M2[j], (97, 90), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(c2[j, :, :, 0], ref)


# ------------------------------------------------------------------------------------------------ a2 / a5 CNN
def _regressor(seed=3):
    from keypoints_detector.networks.regression import landmark_regressor
    return landmark_regressor().init_weights(seed)


def _crops(n, seed):
    rng = np.random.default_rng(seed)
    base = np.stack([gi.image(seed * 100 + i, 128, 128) for i in range(min(n, 8))])
    x = base[np.arange(n) % len(base)].copy()
    x ^= rng.integers(0, 32, x.shape, dtype=np.uint8)
    return x


@pytest.mark.parametrize("B", [1, 5, 33])
def test_regression_net_fp32(dev, B):
    from oracle import cnn as o_cnn
    m = _regressor()
    x = _crops(B, 2)
    ref = o_cnn.regression_forward(x, m.weights, torch.float64)
    out = m.forward_device(T(x, dev), "float32").cpu().numpy()
    # bar: 0.05 px on a 400 px box  ->  1.25e-4 on the normalised output
    assert np.abs(out - ref).max() < 1.25e-4, np.abs(out - ref).max()
    # every trunk level
    levels = o_cnn.trunk_forward(x.astype(np.float64), m.weights, torch.float64, scale=1 / 255.0)
    for t in range(1, 6):
        got = m.intermediate(T(x, dev), t, "float32").cpu().numpy()
        scale = np.abs(levels[t - 1]).max()
        assert np.abs(got - levels[t - 1]).max() < 2e-5 * scale, t


@pytest.mark.parametrize("B", [1, 6, 33])
def test_regression_net_bf16_tensor_cores(dev, B):
    from oracle import cnn as o_cnn
    m = _regressor()
    x = _crops(B, 4)
    ref = o_cnn.regression_forward(x, m.weights, torch.float64)
    out = m.forward_device(T(x, dev), "bfloat16").cpu().numpy()
    # bar: 0.5 px on a 400 px box  ->  1.25e-3 on the normalised output
    err = np.abs(out - ref).max()
    assert err < 1.25e-3, err
    levels = o_cnn.trunk_forward(x.astype(np.float64), m.weights, torch.float64, scale=1 / 255.0)
    for t in range(1, 6):
        got = m.intermediate(T(x, dev), t, "bfloat16").cpu().numpy()
        scale = np.abs(levels[t - 1]).max()
        assert np.abs(got - levels[t - 1]).max() < 3e-2 * scale, (t, np.abs(got - levels[t - 1]).max(), scale)


def test_tc_conv_layers_against_torch(dev):
    """Each tensor-core conv against torch.nn.functional on the SAME bf16-rounded operands (tight: layout bugs
    cannot hide behind bf16 tolerance): first-layer kernel and the TMA implicit-GEMM kernel, pooled and not."""
    import torch.nn.functional as F
    from keypoints_detector import _native as N
    from keypoints_detector.networks.model import Graph, Model
    rng = np.random.default_rng(12)

    def run(h, w, c1, layers, B):
        g = Graph(h, w, 3)
        x = g.conv(0, "conv1", c1, 3, pad=(1, 1, 1, 1), act=N.ACT_RELU, pool=2, bias=True, in_scale=1 / 255.0)
        for (name, cout, k, pad, act, pool) in layers:
            x = g.conv(x, name, cout, k, pad=pad, act=act, pool=pool, bias=True)
        g.dense(x, "fc", 8)
        m = Model(g, "regression", in_dtype="uint8")
        wts = {}
        for k_, shp in m.weight_specs().items():
            wts[k_] = (rng.normal(0, 1.0 / np.sqrt(max(np.prod(shp[:-1]), 1)), shp) * 2).astype(np.float32) if k_.endswith("kernel") \
                else rng.normal(0, 0.1, shp).astype(np.float32)
        m.set_weights(wts)
        xin = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
        xt = T(xin, dev)
        m.forward_device(xt, "bfloat16")
        # torch reference on bf16-rounded operands, fp32 math
        cur = xt.float().permute(0, 3, 1, 2)
        wk = torch.from_numpy(wts["conv1/kernel"] / 255.0).to(dev).bfloat16().float().permute(3, 2, 0, 1)
        cur = F.max_pool2d(F.relu(F.conv2d(F.pad(cur, (1, 1, 1, 1)), wk, torch.from_numpy(wts["conv1/bias"]).to(dev))), 2)
        cur = cur.bfloat16().float()
        got = m.intermediate(xt, 1, "bfloat16").permute(0, 3, 1, 2)
        assert (got - cur).abs().max() <= 2e-2 * cur.abs().max(), ("conv1", float((got - cur).abs().max()), float(cur.abs().max()))
        cur = got  # continue from the kernel's own (bf16) activations
        for i, (name, cout, k, pad, act, pool) in enumerate(layers):
            kh, kw = (k, k) if isinstance(k, int) else k
            if pad == "same":
                th_, tw_ = kh - 1, kw - 1
                pad = (th_ // 2, th_ - th_ // 2, tw_ // 2, tw_ - tw_ // 2)
            wk = torch.from_numpy(wts[name + "/kernel"]).to(dev).bfloat16().float().permute(3, 2, 0, 1)
            y = F.conv2d(F.pad(cur, (pad[2], pad[3], pad[0], pad[1])), wk, torch.from_numpy(wts[name + "/bias"]).to(dev))
            if act == N.ACT_RELU:
                y = F.relu(y)
            if pool:
                y = F.max_pool2d(y, 2)
            got = m.intermediate(xt, 2 + i, "bfloat16").permute(0, 3, 1, 2)
            tol = 1e-2 * y.abs().max()
            assert got.shape == y.shape, (name, got.shape, y.shape)
            assert (got - y).abs().max() <= tol, (name, float((got - y).abs().max()), float(y.abs().max()))
            cur = got

    R = N.ACT_RELU
    run(32, 32, 64, [("conv2", 128, 3, (1, 1, 1, 1), R, 2), ("conv3", 256, 3, (1, 1, 1, 1), R, 2), ("conv4", 256, 3, (1, 1, 1, 1), R, 2)], 3)
    run(64, 48, 64, [("conv2", 64, 3, (1, 1, 1, 1), R, 0), ("conv3", 128, 3, (1, 1, 1, 1), R, 2)], 2)       # un-pooled, ragged tiles
    run(28, 28, 64, [("head7", 512, 7, "same", R, 0), ("head1", 512, 1, (0, 0, 0, 0), R, 0), ("score", 68, 1, (0, 0, 0, 0), 0, 0)], 2)
    run(16, 16, 128, [("conv2", 256, 3, (1, 1, 1, 1), R, 2), ("conv3", 256, 3, (1, 1, 1, 1), R, 2)], 5)     # 4x4 / 2x2 maps, NB > 1
    run(7, 7, 256, [("head7", 512, 7, "same", R, 0), ("head1", 256, 1, (0, 0, 0, 0), R, 0)], 19)           # one-row tiles, skipped kernel rows, ragged batch
    run(7, 7, 64, [("head7", 256, 7, "same", R, 0)], 130)                                                  # one-pixel tiles of 128 images: kernel rows AND columns skipped
    run(3, 5, 64, [("head7", 128, 5, "same", R, 0)], 3)                                                    # TW = 8 on a 5-wide map, 3 rows


# ------------------------------------------------------------------------------------------------ a6 FCN
def _fcn_case(dev, dtype):
    from keypoints_detector.networks.fcn import fcn_8
    from keypoints_detector.data.generator import get_image_array
    from oracle import cnn as o_cnn
    m = fcn_8(68, input_height=64, input_width=96).init_weights(5)
    imgs = [gi.image(90 + i, 120, 160) for i in range(2)]
    x = np.stack([get_image_array(im, 96, 64, ordering="channels_last") for im in imgs])
    probs_ref, logits_ref = o_cnn.fcn_forward(x.astype(np.float64), m.weights, "fcn_8", torch.float64, return_logits=True)
    probs = m.forward_device(T(x, dev), dtype).cpu().numpy()
    return m, probs, probs_ref, logits_ref


def test_fcn8_fp32(dev):
    from keypoints_detector import prediction
    m, probs, probs_ref, logits_ref = _fcn_case(dev, "float32")
    assert probs.shape == probs_ref.shape == (2, 72 * 104, 68)
    assert np.abs(probs - probs_ref).max() < 1e-4
    cm = prediction.class_map_device(T(probs, dev), 72, 104).cpu().numpy()
    ref = probs_ref.reshape(2, 72, 104, 68).argmax(-1)
    assert (cm == ref).mean() > 0.999


def test_fcn8_bf16(dev):
    m, probs, probs_ref, logits_ref = _fcn_case(dev, "bfloat16")
    # bf16 operands through two 4096-wide layers: per-pixel class probabilities move by up to ~0.07 on random-init
    # weights; the bar for this config is the class map, checked below
    assert np.abs(probs - probs_ref).max() < 0.15
    assert np.abs(probs - probs_ref).mean() < 2e-3
    ref = probs_ref.reshape(2, 72, 104, 68).argmax(-1)
    assert (probs.reshape(2, 72, 104, 68).argmax(-1) == ref).mean() > 0.97


def test_fcn8_fused_classmap(dev):
    """fld_net_forward_classmap: argmax fused into the tensor-core up8 epilogue (bf16) / classmap kernel (fp32) must equal
    argmax over the same mode's materialised probabilities, and agree with the oracle."""
    from keypoints_detector import prediction
    from keypoints_detector.networks.fcn import fcn_8
    from keypoints_detector.data.generator import get_image_array
    from oracle import cnn as o_cnn
    m = fcn_8(68, input_height=64, input_width=96).init_weights(5)
    imgs = [gi.image(90 + i, 120, 160) for i in range(3)]
    x = np.stack([get_image_array(im, 96, 64, ordering="channels_last") for im in imgs])
    ref = o_cnn.fcn_forward(x.astype(np.float64), m.weights, "fcn_8", torch.float64).reshape(3, 72, 104, 68).argmax(-1)
    for dtype, rate in (("float32", 0.999), ("bfloat16", 0.97)):
        xt = T(x, dev)
        cm = m.forward_classmap_device(xt, dtype)
        assert cm.dtype == torch.int64 and tuple(cm.shape) == (3, 72, 104)
        probs = m.forward_device(xt, dtype)
        same_mode = prediction.class_map_device(probs, 72, 104)
        assert (cm == same_mode).float().mean().item() > 0.9995          # identical up to exact ties in rounded probabilities
        np.testing.assert_allclose(probs.sum(-1).cpu().numpy(), 1.0, atol=1e-4)
        assert (cm.cpu().numpy() == ref).mean() > rate


def test_fcn8_fused_soft_centroid(dev):
    """fld_net_forward_landmarks (config C3's fused soft-argmax): per-class soft centroid of the softmax output.  fp32 mode
    = forward + fld_decode_heatmap_xy (bit-identical); bf16 mode accumulates the sums inside the tensor-core up8 epilogue:
    equal to the stand-alone decode of the same mode's probabilities to 1e-3 px, and within the 0.5 px bf16 bar of the
    reference formula (utils/metrics.py:56-64) evaluated on the fp64 oracle's probabilities."""
    from keypoints_detector.networks.fcn import fcn_8
    from keypoints_detector.data.generator import get_image_array
    from keypoints_detector.utils import metrics
    from oracle import cnn as o_cnn
    m = fcn_8(68, input_height=64, input_width=96).init_weights(5)
    imgs = [gi.image(90 + i, 120, 160) for i in range(3)]
    x = np.stack([get_image_array(im, 96, 64, ordering="channels_last") for im in imgs])
    pr = o_cnn.fcn_forward(x.astype(np.float64), m.weights, "fcn_8", torch.float64).reshape(3, 72, 104, 68)
    ref = np.empty((3, 68, 2))
    ref[..., 0] = (pr * np.arange(104)[None, None, :, None]).sum((1, 2)) / pr.sum((1, 2))
    ref[..., 1] = (pr * np.arange(72)[None, :, None, None]).sum((1, 2)) / pr.sum((1, 2))
    xt = T(x, dev)
    for dtype, tol in (("float32", 0.05), ("bfloat16", 0.5)):
        xy = m.forward_landmarks_device(xt, dtype)
        assert xy.dtype == torch.float64 and tuple(xy.shape) == (3, 136)
        probs = m.forward_device(xt, dtype).view(3, 72, 104, 68)
        alone = metrics.heatmap_xy_device(probs.contiguous(), 0, 0.0)
        if dtype == "float32":
            assert torch.equal(xy, alone)
        else:
            # the fused decode sums the probabilities on the tensor cores in tf32: each pixel's weight is cut to 11 bits, the SAME
            # cut weight in numerator and denominator, so the centroid moves by <= 2^-10 x (spread of the map) — measured 3e-3 px
            # on these near-uniform random-weight maps, against the 0.5 px bar of the bf16 mode just below
            assert (xy - alone).abs().max().item() < 1e-2
        assert np.abs(xy.cpu().numpy().reshape(3, 68, 2) - ref).max() < tol
    # sentinel: mean probability 1/68 <= thresh -> (-1, -1) for every class
    assert (m.forward_landmarks_device(xt, "bfloat16", thresh=0.5) == -1).all()
    # top-n centroid (the reference wrappers' effective default is n_points = 4) behind the same entry point: probabilities are
    # materialised in the workspace and decoded by the stand-alone kernel, bit-identical to doing the two steps by hand
    for dtype in ("bfloat16", "float32"):
        probs = m.forward_device(xt, dtype).view(3, 72, 104, 68).contiguous()
        for n in (1, 3, 4, 9):
            xy = m.forward_landmarks_device(xt, dtype, n_points=n)
            assert torch.equal(xy, metrics.heatmap_xy_device(probs, n, 0.0)), (dtype, n)
    assert (m.forward_landmarks_device(xt, "bfloat16", thresh=1.5, n_points=4) == -1).all()   # probabilities never exceed 1


def test_prediction_dropin_fcn(dev, tmp_path):
    """keypts_predict / _prediction / model_from_checkpoint_path round trip (reference prediction.py:116-222)."""
    from keypoints_detector import prediction
    from keypoints_detector.networks.basic_models import LANDMARKS_MODELS
    from oracle import cnn as o_cnn, decode as o_dec, preprocess as o_pre
    m = LANDMARKS_MODELS["fcn_8_vanilla"](68, input_height=64, input_width=64).init_weights(6)
    ck = str(tmp_path / "w")
    m.save_weights(ck + ".00003")
    m.save_config(ck, "fcn_8_vanilla")
    img = gi.image(95, 200, 260)
    cv2.imwrite(str(tmp_path / "in.png"), img)
    pr = prediction.keypts_predict(inp=str(tmp_path / "in.png"), checkpoints_path=ck, out_fname=str(tmp_path / "out.png"))
    assert pr.shape == (72, 72) and pr.dtype == np.int64
    x = np.ascontiguousarray(o_pre.get_image_array(cv2.imread(str(tmp_path / "in.png")), 64, 64))
    ref = o_dec.class_map(o_cnn.fcn_forward(x[None].astype(np.float64), m.weights, "fcn_8", torch.float64)[0], 72, 72, 68)
    assert (pr == ref).mean() > 0.999
    assert cv2.imread(str(tmp_path / "out.png")).shape[:2] == (200, 260)


# ------------------------------------------------------------------------------------------------ whole path
def test_pipeline_end_to_end(dev):
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    from oracle import align as o_al, cnn as o_cnn, decode as o_dec, preprocess as o_pre
    frames = synthetic.make_frames(3, 480, 640, seed=21)
    boxes = synthetic.make_boxes(20, 480, 640, seed=22, max_side=300)
    f2f = (np.arange(20) % 3).astype(np.int32)
    m = _regressor(9)
    crops, fbs = zip(*[o_pre.crop_resize_rgb(frames[f2f[i]], boxes[i]) for i in range(20)])
    out = o_cnn.regression_forward(np.stack(crops), m.weights, torch.float64)
    marks_ref = np.stack([o_dec.regression_decode(out[i], fbs[i])[0] for i in range(20)])
    _, aligned_ref = o_al.align_faces(frames, f2f, marks_ref)
    for dtype, tol in (("float32", 0.05), ("bfloat16", 0.5)):
        r = prediction.LandmarkPipeline(m, dtype=dtype)(frames, boxes, f2f)
        assert np.array_equal(r["crops"], np.stack(crops))
        assert np.abs(r["marks"] - marks_ref).max() <= tol, (dtype, np.abs(r["marks"] - marks_ref).max())
        for i in range(20):                                                     # crops bit-exact for the GPU's own landmarks
            M = o_al.umeyama(o_al.five_points(r["marks"][i]), o_al.TEMPLATE_112)
            assert np.array_equal(M, r["M"][i])
            assert np.array_equal(r["aligned"][i], o_al.warp_affine_u8(frames[f2f[i]], M, 112, 112))
        if dtype == "float32":
            d = np.abs(r["aligned"].astype(int) - aligned_ref.astype(int))
            assert (d <= 1).mean() > 0.97                                       # chained against the fp64-oracle landmarks
    # drop-in single-face API: (68,2) np.uint, equal to the oracle chain up to the float->uint boundary
    marks = prediction.detect_marks(frames[0], m, boxes[0].tolist())
    assert marks.shape == (68, 2) and marks.dtype == np.uint
    assert np.abs(marks.astype(np.int64) - np.maximum(marks_ref[0], 0).astype(np.int64)).max() <= 1
    with pytest.raises(ValueError):
        prediction.detect_marks(frames[0], m, [2, 2, 60, 200])                 # squared box leaves the image on the left


def test_pipeline_chunks_lanes_and_graph_bit_identical(dev):
    """Size-independent properties of the batched pipeline (config C5): chunked execution (max_batch), a second lane on its
    own stream and a replayed CUDA graph all give results bit-identical to one plain run."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    frames = T(synthetic.make_frames(2, 480, 640, seed=31), dev)
    boxes = T(synthetic.make_boxes(70, 480, 640, seed=32, max_side=300), dev)
    f2f = T((np.arange(70) % 2).astype(np.int32), dev)
    pipe = prediction.LandmarkPipeline(_regressor(9), dtype="bfloat16")
    keys = ("marks", "aligned", "M", "faceboxes", "crops")
    ref = {k: v.clone() for k, v in pipe.run_device(frames, boxes, f2f).items() if k in keys}
    pipe.max_batch = 32                                                          # 32 + 32 + 6
    got = pipe.run_device(frames, boxes, f2f)
    for k in keys:
        assert torch.equal(got[k], ref[k]), k
    pipe.max_batch = 4096
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        got1 = pipe.run_device(frames, boxes, f2f, lane=1)
    torch.cuda.current_stream(dev).wait_stream(side)
    for k in keys:
        assert torch.equal(got1[k], ref[k]), k
    g, res = pipe.capture(frames, boxes, f2f, lane=2)
    for k in keys:
        res[k].zero_()
    g.replay()
    torch.cuda.synchronize(dev)
    for k in keys:
        assert torch.equal(res[k], ref[k]), k


def test_full_size_properties_c3_c5(dev):
    """Size-independent properties at BASELINE.json's sizes (no oracle can run these in seconds):
    C3  fcn_8/vanilla @224x224, 68 classes, batch 64: probabilities sum to 1, the fused class map equals the argmax of the
        materialised probabilities (exact ties aside), the fused soft centroid equals the stand-alone decode, and a permuted
        batch gives the permuted result bit for bit (no cross-image coupling anywhere in the path);
    C5  5000 faces through the regression pipeline (crosses the 4096-face chunk boundary) equal the same faces run in
        two halves, bit for bit."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    from keypoints_detector.networks.fcn import fcn_8
    from keypoints_detector.utils import metrics
    m = fcn_8(68, input_height=224, input_width=224).init_weights(3)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn((64, 224, 224, 3), generator=g) * 50).to(dev)
    probs = m.forward_device(x, "bfloat16")
    assert tuple(probs.shape) == (64, 232 * 232, 68)
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-4
    cm = m.forward_classmap_device(x, "bfloat16")
    assert (cm.view(64, -1) == probs.argmax(-1)).float().mean().item() > 0.9995
    xy = m.forward_landmarks_device(x, "bfloat16")
    alone = metrics.heatmap_xy_device(probs.view(64, 232, 232, 68), 0, 0.0)
    assert (xy - alone).abs().max().item() < 1e-2   # tf32 pixel weights in the fused reduction, see test_fcn8_fused_soft_centroid
    perm = torch.randperm(64, generator=g).to(dev)
    cm_p = m.forward_classmap_device(x[perm].contiguous(), "bfloat16")
    assert torch.equal(cm_p, cm[perm])
    del probs, cm, cm_p, alone
    torch.cuda.empty_cache()

    frames = T(synthetic.make_frames(8, 480, 640, seed=41), dev)
    boxes = T(synthetic.make_boxes(5000, 480, 640, seed=42, max_side=300), dev)
    f2f = T((np.arange(5000) % 8).astype(np.int32), dev)
    pipe = prediction.LandmarkPipeline(_regressor(9), dtype="bfloat16")
    keys = ("marks", "aligned", "M")
    full = {k: v.clone() for k, v in pipe.run_device(frames, boxes, f2f).items() if k in keys}
    for lo, hi in ((0, 2500), (2500, 5000)):
        part = pipe.run_device(frames, boxes[lo:hi].contiguous(), f2f[lo:hi].contiguous(), lane=1)
        for k in keys:
            assert torch.equal(part[k], full[k][lo:hi]), (k, lo)


def test_multi_gpu_shards_bit_identical(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    frames = synthetic.make_frames(2, 480, 640, seed=31)
    boxes = synthetic.make_boxes(16, 480, 640, seed=32, max_side=300)
    f2f = (np.arange(16) % 2).astype(np.int32)
    m = _regressor(10)
    full = prediction.LandmarkPipeline(m, dtype="bfloat16", device=0)(frames, boxes, f2f)
    parts = []
    for g, (lo, hi) in enumerate(prediction.shard_faces(16, 2)):
        parts.append(prediction.LandmarkPipeline(m, dtype="bfloat16", device=g)(frames, boxes[lo:hi], f2f[lo:hi]))
    for k in ("marks", "aligned"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k])


@pytest.mark.parametrize("n_classes", [12, 21])
def test_fcn8_other_class_counts_bf16(dev, n_classes):
    """The tensor-core transposed conv has a compile-time variant for 68 classes; other counts take the generic variant
    (12: 16-byte vector stores, 21: scalar stores).  Same bars as test_fcn8_bf16 / test_fcn8_fused_classmap."""
    from keypoints_detector.networks.fcn import fcn_8
    from oracle import cnn as o_cnn
    m = fcn_8(n_classes, input_height=64, input_width=96).init_weights(9)
    x = np.random.default_rng(9).normal(0, 40, (3, 64, 96, 3)).astype(np.float32)
    probs_ref = o_cnn.fcn_forward(x.astype(np.float64), m.weights, "fcn_8", torch.float64)
    xt = T(x, dev)
    probs = m.forward_device(xt, "bfloat16").cpu().numpy()
    assert probs.shape == probs_ref.shape == (3, 72 * 104, n_classes)
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-4)
    assert np.abs(probs - probs_ref).mean() < 4e-3
    ref = probs_ref.reshape(3, 72, 104, n_classes).argmax(-1)
    assert (probs.reshape(3, 72, 104, n_classes).argmax(-1) == ref).mean() > 0.97
    cm = m.forward_classmap_device(xt, "bfloat16").cpu().numpy()
    assert (cm == ref).mean() > 0.97


def test_error_behaviour_through_the_abi(dev):
    """Errors surface as exceptions with the library's message (no silent fallback): wrong shapes / dtypes, an undersized
    workspace, an unfinalised net, unsupported arguments; empty batches are legal and return empty results."""
    import ctypes
    from keypoints_detector import _native as N, prediction
    from keypoints_detector.utils import metrics
    m = _regressor(3)
    with pytest.raises(ValueError):
        m.forward_device(torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device=dev))
    with pytest.raises(TypeError):
        m.forward_device(torch.zeros((1, 128, 128, 3), dtype=torch.float32, device=dev))
    out = m.forward_device(torch.zeros((0, 128, 128, 3), dtype=torch.uint8, device=dev))
    assert tuple(out.shape) == (0, 136)
    lib = N.load_library()
    net = m.compiled(dev.index, "bfloat16")
    x = torch.zeros((2, 128, 128, 3), dtype=torch.uint8, device=dev)
    ws = torch.empty(4096 + 1024, dtype=torch.uint8, device=dev)
    al = (-ws.data_ptr()) % 1024
    o = torch.empty((2, 136), dtype=torch.float32, device=dev)
    rc = lib.fld_net_forward(net, N.ptr(x), 2, N._vp(ws.data_ptr() + al), 4096, N.ptr(o), N.stream_ptr(dev.index))
    assert rc != 0 and b"workspace" in lib.fld_last_error()
    with pytest.raises(N.FldError):
        N.check(rc)
    with pytest.raises(ValueError):
        metrics.heatmap_xy_device(torch.zeros((1, 8, 8, 2), device=dev), n_points=N.MAX_TOPN + 1)
    frames = torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device=dev)
    crops, M = prediction.align_device(frames, torch.zeros(0, dtype=torch.int32, device=dev),
                                       torch.zeros((0, 68, 2), dtype=torch.float32, device=dev))
    assert tuple(crops.shape) == (0, 112, 112, 3) and tuple(M.shape) == (0, 2, 3)
    # degenerate landmarks (all points equal): the fit has no solution -> NaN matrix, zero crop, no crash
    marks = torch.full((1, 68, 2), 10.0, device=dev)
    crops, M = prediction.align_device(frames, torch.zeros(1, dtype=torch.int32, device=dev), marks)
    assert torch.isnan(M).all() and int(crops.sum()) == 0


def test_no_out_of_bounds_writes(dev):
    """Guard-band check (compute-sanitizer is not available on the pool): every caller-owned output lives inside a larger
    sentinel-filled buffer; after the calls the bands on both sides, and the workspace tail, must be untouched."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    from keypoints_detector.networks.fcn import fcn_8

    def guarded(shape, dtype, pad=4096):
        n = int(np.prod(shape))
        es = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((n * es + 2 * pad,), 0xA5, dtype=torch.uint8, device=dev)
        view = raw[pad:pad + n * es].view(dtype).view(shape)
        return raw, view, pad

    def intact(raw, pad):
        return bool((raw[:pad] == 0xA5).all()) and bool((raw[-pad:] == 0xA5).all())

    frames = T(synthetic.make_frames(2, 240, 320, seed=51), dev)
    boxes = T(synthetic.make_boxes(5, 240, 320, seed=52, max_side=150), dev)
    f2f = T((np.arange(5) % 2).astype(np.int32), dev)
    m = _regressor(4)
    for dtype in ("bfloat16", "float32"):
        bufs = {}
        rc, crops, p = guarded((5, 128, 128, 3), torch.uint8); bufs["crops"] = (rc, p)
        rb, fb, p = guarded((5, 4), torch.int32); bufs["fb"] = (rb, p)
        prediction.preprocess_faces_device(frames, boxes, f2f, 128, True, out=crops, out_boxes=fb)
        ro, out, p = guarded((5, 136), torch.float32); bufs["out"] = (ro, p)
        _, ws, al = m.forward_device(crops, dtype, out=out, return_workspace=True)
        lib = N_lib()
        need = lib.fld_net_workspace_bytes(m.compiled(dev.index, dtype), 5)
        ws[al + need:].fill_(0x5A)
        m.forward_device(crops, dtype, out=out)
        assert bool((ws[al + need:] == 0x5A).all()), "workspace tail overwritten (%s)" % dtype
        rm, marks, p = guarded((5, 68, 2), torch.float32); bufs["marks"] = (rm, p)
        ru, marks_u, p = guarded((5, 68, 2), torch.int64); bufs["marks_u"] = (ru, p)
        prediction.decode_regress_device(out, fb, True, out=marks, out_uint=marks_u)
        ra, aligned, p = guarded((5, 112, 112, 3), torch.uint8); bufs["aligned"] = (ra, p)
        rM, M, p = guarded((5, 2, 3), torch.float64); bufs["M"] = (rM, p)
        prediction.align_device(frames, f2f, marks, None, (112, 112), True, True, out=aligned, out_matrix=M)
        torch.cuda.synchronize(dev)
        for k, (raw, pad) in bufs.items():
            assert intact(raw, pad), (k, dtype)
    net = fcn_8(68, input_height=64, input_width=96).init_weights(5)
    x = torch.randn((3, 64, 96, 3), device=dev) * 40
    for dtype in ("bfloat16", "float32"):
        raw, out, pad = guarded((3, 72 * 104, 68), torch.float32)
        net.forward_device(x, dtype, out=out)
        torch.cuda.synchronize(dev)
        assert intact(raw, pad), dtype
        assert (out.sum(-1) - 1).abs().max().item() < 1e-4
    # ordered alignment (fit kernel + counting sort + tile kernel): crops, matrices and the caller's scratch, exactly sized
    B = prediction.ALIGN_ORDER_MIN_FACES
    pts, _ = synthetic.make_similarity_landmarks(B, 240, 320, prediction.TEMPLATE_112, seed=53, scale=(0.9, 1.4))
    f2b = T((np.arange(B) % 2).astype(np.int32), dev)
    lib = N_lib()
    need = int(lib.fld_align_scratch_bytes(prediction.N.handle(dev), B))
    rs, scratch, ps = guarded((need,), torch.uint8)
    ra, aligned, pa = guarded((B, 112, 112, 3), torch.uint8)
    rM, M, pM = guarded((B, 2, 3), torch.float64)
    prediction.align_device(frames, f2b, T(pts, dev), None, (112, 112), False, True, out=aligned, out_matrix=M, scratch=scratch)
    torch.cuda.synchronize(dev)
    assert intact(rs, ps) and intact(ra, pa) and intact(rM, pM)
    # top-4 heat-map decode: partial lists in the caller's scratch, exactly sized
    from keypoints_detector.utils import metrics
    hm = torch.rand((3, 40, 56, 68), device=dev)
    need = int(lib.fld_decode_heatmap_scratch_bytes(prediction.N.handle(dev), 3, 40, 56, 68, 4))
    rs, scratch, ps = guarded((need,), torch.uint8)
    rx, xy, px = guarded((3, 136), torch.float64)
    prediction.N.check(lib.fld_decode_heatmap_xy(prediction.N.handle(dev), prediction.N.ptr(hm), 3, 40, 56, 68, 4, 0.0, prediction.N.ptr(xy),
                                                 prediction.N.ptr(scratch), need, prediction.N.stream_ptr(dev)))
    torch.cuda.synchronize(dev)
    assert intact(rs, ps) and intact(rx, px)
    ref = metrics.heatmap_xy_device(hm, 4, 0.0)
    assert torch.equal(ref.view(-1), xy.view(-1))


def N_lib():
    from keypoints_detector import _native as N
    return N.load_library()


def test_cli_predict_end_to_end(dev, tmp_path):
    """`face_landmark_detect predict --checkpoints_path P --inp IMG` (reference scripts/cli.py:68-74) through
    a saved checkpoint + `_config.json` sidecar; the coloured class map it writes decodes back to the model's own class map."""
    import cv2
    import sys
    from click.testing import CliRunner
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1] / "face-landmark-detector_b200"))
    from scripts import cli
    from keypoints_detector import prediction
    from keypoints_detector.networks.basic_models import LANDMARKS_MODELS
    m = LANDMARKS_MODELS["fcn_8_vanilla"](68, input_height=64, input_width=64).init_weights(8)
    ck = str(tmp_path / "w")
    m.save_weights(ck + ".00001")
    m.save_config(ck, "fcn_8_vanilla")
    img = gi.image(77, 90, 120)
    inp = str(tmp_path / "face.png")
    cv2.imwrite(inp, img)
    out = str(tmp_path / "seg.png")
    res = CliRunner().invoke(cli.main, ["predict", "--checkpoints_path", ck, "--inp", inp, "--out_fname", out])
    assert res.exit_code == 0, res.output
    seg = cv2.imread(out)
    assert seg is not None and seg.shape[2] == 3
    cm = prediction.keypts_predict(checkpoints_path=ck, inp=inp)
    assert cm.shape == (m.output_height, m.output_width) and cm.dtype == np.int64
