"""Property tests: facts that hold for every input, checked on the CPU oracle with hypothesis and on the
CUDA path (pytest -m gpu) with seeded sweeps:
  * the Umeyama fit recovers a known similarity (scale, rotation, translation);
  * warping with the identity matrix is a crop of the frame;
  * decoding a one-hot heat map returns the coordinate of its hot pixel, the class map of one-hot scores its class."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import align as o_al, decode as o_dec


def _similarity(s, deg, tx, ty):
    th = np.deg2rad(deg)
    return np.array([[s * np.cos(th), -s * np.sin(th), tx], [s * np.sin(th), s * np.cos(th), ty]])


@settings(max_examples=60, deadline=None)
@given(st.floats(0.2, 3.0), st.floats(-179, 179), st.floats(-300, 300), st.floats(-300, 300), st.integers(0, 2 ** 31 - 1))
def test_oracle_umeyama_recovers_known_similarity(s, deg, tx, ty, seed):
    M = _similarity(s, deg, tx, ty)
    src = np.random.default_rng(seed).uniform(0, 500, (7, 2))
    dst = src @ M[:, :2].T + M[:, 2]
    np.testing.assert_allclose(o_al.umeyama(src, dst), M, rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(o_al.umeyama_svd(src, dst), M, rtol=1e-9, atol=1e-7)


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(120, 200), st.integers(120, 200))
def test_oracle_identity_warp_is_a_crop(seed, h, w):
    frame = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    crop = o_al.warp_affine_u8(frame, np.array([[1.0, 0, 0], [0, 1.0, 0]]), 112, 112)
    assert np.array_equal(crop, frame[:112, :112])
    shifted = o_al.warp_affine_u8(frame, np.array([[1.0, 0, -5], [0, 1.0, -3]]), 100, 100)   # crop = frame shifted by (5, 3)
    assert np.array_equal(shifted, frame[3:103, 5:105])


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 40), st.integers(2, 40), st.integers(0, 10 ** 6), st.integers(-1, 4))
def test_oracle_one_hot_heatmap_decodes_to_its_coordinate(h, w, pos, n):
    y, x = (pos // w) % h, pos % w
    hm = np.zeros((h, w), dtype=np.float32)
    hm[y, x] = 0.75
    xy = o_dec.average_xy(hm, n_points=min(n, h * w) if n > 0 else n, thresh=0.0) if n != 0 else o_dec.average_xy(hm, 0, 0.0)
    assert list(xy) == [x, y]
    probs = np.zeros((h * w, 5), dtype=np.float32)
    probs[:, 1] = 0.2
    probs[y * w + x, 3] = 0.9
    cm = o_dec.class_map(probs, h, w, 5)
    assert cm[y, x] == 3 and (np.delete(cm.reshape(-1), y * w + x) == 1).all()


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.fixture(scope="module")
def dev(cuda_lib):
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def _T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).contiguous()


def _bf16(a):
    """round-to-nearest-even bf16 of a float32 array (what the library's f2bf / __float2bfloat16_rn do)"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10_000), scale=st.sampled_from([1e-3, 1.0, 50.0, 4e3]))
def test_bf16x3_split_product_error_bound(seed, scale):
    """The arithmetic FLD_BF16X3 rests on: with hi = bf16(v), lo = bf16(v - hi), the three-term product
    x_hi w_hi + x_lo w_hi + x_hi w_lo (exact products of bf16 values, fp32 / here fp64 accumulation) misses x w by the dropped
    x_lo w_lo term and the remainders' own rounding: every product within 3 * 2^-16 |x w| (worst case: operands just above a power
    of two), half of them within 2^-17, two orders of magnitude tighter than plain bf16 operands — and a K = 2304 dot product (conv3's
    depth) stays within 2^-16 of the sum of |terms|."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(2304) * scale).astype(np.float32)
    w = (rng.standard_normal(2304) * 0.05).astype(np.float32)
    xh, wh = _bf16(x), _bf16(w)
    xl, wl = _bf16(x - xh), _bf16(w - wh)
    exact = x.astype(np.float64) * w.astype(np.float64)
    three = xh.astype(np.float64) * wh + xl.astype(np.float64) * wh + xh.astype(np.float64) * wl
    rel = np.abs(three - exact) / np.maximum(np.abs(exact), 1e-300)
    assert rel.max() <= 3 * 2.0 ** -16 and np.median(rel) <= 2.0 ** -17
    one = xh.astype(np.float64) * wh
    assert np.abs(one - exact).max() > 20 * np.abs(three - exact).max()          # what plain bf16 operands would lose
    assert abs(three.sum() - exact.sum()) <= 2.0 ** -16 * np.abs(exact).sum()


@pytest.mark.gpu
def test_gpu_fit_recovers_known_similarity_and_identity_warp_is_a_crop(dev):
    from keypoints_detector import prediction
    rng = np.random.default_rng(7)
    n = 64
    frames = rng.integers(0, 256, (2, 160, 200, 3), dtype=np.uint8)
    Ms = np.stack([_similarity(rng.uniform(0.3, 2.5), rng.uniform(-170, 170), rng.uniform(-100, 100), rng.uniform(-100, 100))
                   for _ in range(n)])
    tmpl = rng.uniform(0, 112, (68, 2))
    # landmarks = template pulled back through the inverse similarity: the fit must return Ms (frame -> crop)
    marks = np.stack([(tmpl - M[:, 2]) @ np.linalg.inv(M[:, :2]).T for M in Ms])
    crops, M = prediction.align_device(_T(frames, dev), _T((np.arange(n) % 2).astype(np.int32), dev), _T(marks.astype(np.float32), dev),
                                       tmpl, (112, 112), five_point=False)
    # float32 landmarks carry ~1e-7 relative error of their own
    np.testing.assert_allclose(M.cpu().numpy(), Ms, rtol=2e-4, atol=2e-3)
    eye = np.tile(np.array([[1.0, 0, 0], [0, 1.0, 0]]), (2, 1, 1))
    out = prediction.warp_affine_device(_T(frames, dev), _T(np.arange(2, dtype=np.int32), dev), _T(eye, dev), (112, 112)).cpu().numpy()
    assert np.array_equal(out, frames[:, :112, :112])


@pytest.mark.gpu
def test_gpu_one_hot_heatmaps_decode_to_their_coordinates(dev):
    from keypoints_detector import prediction
    from keypoints_detector.utils import metrics
    rng = np.random.default_rng(11)
    B, H, W, L = 3, 37, 52, 68
    ys, xs = rng.integers(0, H, (B, L)), rng.integers(0, W, (B, L))
    hm = np.zeros((B, H, W, L), dtype=np.float32)
    for b in range(B):
        for l in range(L):
            hm[b, ys[b, l], xs[b, l], l] = rng.uniform(0.1, 1.0)
    expect = np.stack([xs, ys], -1).reshape(B, 2 * L).astype(np.float64)
    for n in (1, 2, 4, 0):
        xy = metrics.heatmap_xy_device(_T(hm, dev), n, 0.0).cpu().numpy()
        np.testing.assert_allclose(xy, expect, rtol=0, atol=1e-12 if n else 1e-4)
    scores = np.full((B, H * W, L), 0.01, dtype=np.float32)
    cls = rng.integers(0, L, (B, H * W))
    np.put_along_axis(scores, cls[..., None], 0.9, axis=2)
    cm = prediction.class_map_device(_T(scores, dev), H, W).cpu().numpy()
    assert np.array_equal(cm.reshape(B, -1), cls)
