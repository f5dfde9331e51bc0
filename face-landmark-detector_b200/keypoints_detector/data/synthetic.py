"""Seeded synthetic inputs for example.py, the tests and bench.py (there are no datasets offline):
frames = uniform noise blended 50/50 with a smooth sinusoid field, detector boxes, and
similarity-transformed five-point landmark sets."""
import numpy as np


def make_frames(n, height, width, seed=0):
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (n, height, width, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    out = np.empty_like(noise)
    for i in range(n):
        ph = rng.uniform(0, 2 * np.pi, 3)
        fx, fy = rng.uniform(0.005, 0.03, 2)
        smooth = np.stack([127.5 + 127.5 * np.sin(xx * fx + yy * fy + p) for p in ph], -1)
        out[i] = (0.5 * noise[i] + 0.5 * smooth).astype(np.uint8)
    return out


def make_boxes(n, height, width, seed=0, min_side=96, max_side=400):
    """Detector-style boxes fully inside the frame after the reference's shift-down + squaring."""
    rng = np.random.default_rng(seed)
    max_side = min(max_side, int(min(height, width) * 0.8))
    boxes = np.zeros((n, 4), dtype=np.int32)
    for i in range(n):
        w = int(rng.integers(min_side, max_side))
        h = int(np.clip(w + rng.integers(-w // 6, w // 6 + 1), min_side, max_side))
        side = max(w, h)
        margin = side // 2 + side // 8 + 2
        cx = int(rng.integers(margin, width - margin))
        cy = int(rng.integers(margin, height - margin - h // 10))
        boxes[i] = (cx - w // 2, cy - h // 2, cx - w // 2 + w, cy - h // 2 + h)
    return boxes


def make_similarity_landmarks(n, height, width, template, seed=0, scale=(0.28, 1.4), rot_deg=30.0, jitter=1.0, out_size=112):
    """Landmarks = template mapped by a random inverse similarity (config C4): returns (pts [n,K,2] float32,
    M_true [n,2,3] frame->crop)."""
    rng = np.random.default_rng(seed)
    template = np.asarray(template, dtype=np.float64)
    pts = np.zeros((n,) + template.shape, dtype=np.float32)
    Ms = np.zeros((n, 2, 3))
    for i in range(n):
        s = rng.uniform(*scale)
        th = np.deg2rad(rng.uniform(-rot_deg, rot_deg))
        L = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        half = out_size / 2 / s
        cx = rng.uniform(half, width - half)
        cy = rng.uniform(half, height - half)
        t = np.array([out_size / 2, out_size / 2]) - L @ np.array([cx, cy])
        Li = np.linalg.inv(L)
        p = (template - t) @ Li.T + rng.normal(0, jitter, template.shape)
        pts[i] = p.astype(np.float32)
        Ms[i] = np.concatenate([L, t[:, None]], 1)
    return pts, Ms
