"""Seeded inputs shared by tests/make_golden.py (which runs the reference / cv2 on them) and the tests
(which run the oracle and the CUDA path on the same bytes).  numpy Generator streams are stable across
numpy versions for integers / uniform / normal with PCG64."""
import numpy as np


def image(seed, h, w):
    """uint8 BGR noise blended 50/50 with a smooth sinusoid field."""
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    ph = rng.uniform(0, 2 * np.pi, 3)
    smooth = np.stack([127.5 + 127.5 * np.sin(xx * 0.021 + yy * 0.013 + p) for p in ph], -1)
    return (0.5 * noise + 0.5 * smooth).astype(np.uint8)


# (seed, frame h, frame w, face box) — boxes exercise: square, tall (even/odd diff), wide (even/odd diff),
# exact-2x crop (256 -> 128), same-size crop (128), box leaving the frame bottom/right (numpy slice clipping)
DETECT_CASES = [
    (11, 480, 640, [200, 120, 400, 360]),
    (12, 480, 640, [100, 100, 301, 333]),
    (13, 480, 640, [50, 60, 351, 260]),
    (14, 480, 640, [320, 40, 470, 249]),
    (15, 600, 800, [100, 100, 356, 356]),
    (16, 300, 400, [130, 80, 258, 208]),
    (17, 480, 640, [450, 250, 630, 470]),
    (18, 480, 640, [10, 5, 107, 120]),
]


def fake_outputs(seed, n=1):
    """Stand-in CNN outputs: normalised landmark coordinates in [0,1) (float32 like a TF model returns)."""
    rng = np.random.default_rng(1000 + seed)
    return rng.uniform(0.05, 0.95, (n, 136)).astype(np.float32)


def heatmaps(seed, n, h, w, l):
    """Peaky positive heat-maps (a few Gaussian blobs + noise floor), float32."""
    rng = np.random.default_rng(2000 + seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    hm = rng.uniform(0, 0.02, (n, h, w, l))
    for i in range(n):
        for j in range(l):
            cy, cx = rng.uniform(2, h - 3), rng.uniform(2, w - 3)
            hm[i, :, :, j] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * rng.uniform(0.8, 2.5) ** 2))
    return hm.astype(np.float32)


def probs(seed, hw, n_classes):
    rng = np.random.default_rng(3000 + seed)
    z = rng.normal(0, 2, (1, hw, n_classes))
    e = np.exp(z - z.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def similarity(seed, w=1920, h=1080, out=112):
    rng = np.random.default_rng(4000 + seed)
    s = rng.uniform(0.28, 1.4)
    th = np.deg2rad(rng.uniform(-30, 30))
    L = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cx, cy = rng.uniform(0, w), rng.uniform(0, h)
    t = np.array([out / 2, out / 2]) - L @ np.array([cx, cy])
    return np.concatenate([L, t[:, None]], 1)


RESIZE_SHAPES = [(96, 96, 128, 128), (97, 131, 128, 128), (256, 256, 128, 128), (128, 128, 128, 128), (300, 211, 128, 128),
                 (480, 640, 224, 224), (57, 33, 96, 64)]


# a .pts landmark file exactly as the reference's writer emits it (scripts/prepare_dataset.py:46-52)
PTS_POINTS = [(66.03356, 39.00227), (30.22701, 36.42168), (59.58208, 39.6474), (73.13035, 39.96999), (36.35657, 37.38940),
              (23.45287, 37.38940), (56.95326, 29.03365), (80.22713, 32.22814), (40.22761, 29.00232), (16.35638, 29.64747)]
PTS_TEXT = "version: 1\nn_points: %d\n{\n" % len(PTS_POINTS) + "".join("%s %s\n" % (str(x), str(y)) for x, y in PTS_POINTS) + "}"
