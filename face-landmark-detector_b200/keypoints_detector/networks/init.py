"""Seeded random initialisation in Keras layouts (what Keras does when a model is built without a
checkpoint): glorot_uniform kernels (Keras default), he_normal for the FCN score convs
(reference fcn.py:103,108,117), biases ~ N(0, 0.05), and — so that BN folding is actually exercised by
parity tests — non-trivial BatchNormalization statistics.

The regression model's Dense layer is scaled down and biased with a canonical 68-point face layout so
that synthetic outputs land inside the face box (normalised coordinates in ~[0.2, 0.8]).
"""
import numpy as np

from .. import _native as N


def canonical_face68():
    """A fixed, plausible normalised 68-point layout (jaw arc, brows, nose, eyes, mouth)."""
    pts = np.zeros((68, 2), dtype=np.float64)
    t = np.linspace(np.pi * 0.95, np.pi * 0.05, 17)                      # jaw 0..16
    pts[0:17] = np.stack([0.5 + 0.36 * np.cos(t), 0.42 + 0.42 * np.sin(t)], 1)
    pts[17:22] = np.stack([np.linspace(0.22, 0.42, 5), 0.30 - 0.02 * np.sin(np.linspace(0, np.pi, 5))], 1)
    pts[22:27] = np.stack([np.linspace(0.58, 0.78, 5), 0.30 - 0.02 * np.sin(np.linspace(0, np.pi, 5))], 1)
    pts[27:31] = np.stack([np.full(4, 0.5), np.linspace(0.38, 0.54, 4)], 1)   # nose bridge, 30 = tip
    pts[31:36] = np.stack([np.linspace(0.43, 0.57, 5), np.full(5, 0.60)], 1)
    for k, cx in ((36, 0.33), (42, 0.67)):                                # eyes 36..41, 42..47
        a = np.linspace(0, 2 * np.pi, 7)[:6]
        pts[k:k + 6] = np.stack([cx - 0.055 * np.cos(a), 0.40 - 0.025 * np.sin(a)], 1)
    a = np.linspace(0, 2 * np.pi, 13)[:12]                                # outer mouth 48..59 (48 left, 54 right corner)
    pts[48:60] = np.stack([0.5 - 0.12 * np.cos(a), 0.74 - 0.05 * np.sin(a)], 1)
    a = np.linspace(0, 2 * np.pi, 9)[:8]                                  # inner mouth 60..67
    pts[60:68] = np.stack([0.5 - 0.07 * np.cos(a), 0.74 - 0.02 * np.sin(a)], 1)
    return pts


def _bn_stats(rng, w, bn, c, nontrivial_bn):
    if nontrivial_bn:
        w[bn + "/gamma"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        w[bn + "/beta"] = rng.normal(0, 0.1, c).astype(np.float32)
        w[bn + "/moving_mean"] = rng.normal(0, 0.1, c).astype(np.float32)
        w[bn + "/moving_variance"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    else:  # fresh Keras BN: gamma 1, beta 0, mean 0, var 1  (y = x / sqrt(1.001), not identity)
        w[bn + "/gamma"] = np.ones(c, np.float32)
        w[bn + "/beta"] = np.zeros(c, np.float32)
        w[bn + "/moving_mean"] = np.zeros(c, np.float32)
        w[bn + "/moving_variance"] = np.ones(c, np.float32)


def random_weights(model, seed=0, nontrivial_bn=True):
    rng = np.random.default_rng(seed)
    g = model.graph
    w = {}
    for L in g.layers:
        n = L["name"]
        cin = g.shapes[L["in0"]][2]
        if L["op"] == N.OP_CONV:
            kh, kw, cout = L["kh"], L["kw"], L["cout"]
            fan_in, fan_out = kh * kw * cin, kh * kw * cout
            if n.startswith("score"):
                k = rng.normal(0.0, np.sqrt(2.0 / fan_in), (kh, kw, cin, cout))            # he_normal
            else:
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                k = rng.uniform(-lim, lim, (kh, kw, cin, cout))                             # glorot_uniform
                if L["act"] != N.ACT_NONE:
                    k *= 1.6  # keep activation scale roughly constant through the ReLU/pool stages
            w[n + "/kernel"] = k.astype(np.float32)
            if L["has_bias"]:
                w[n + "/bias"] = rng.normal(0, 0.05, cout).astype(np.float32)
            if L["has_bn"]:
                _bn_stats(rng, w, L["bn_name"], cout, nontrivial_bn)
        elif L["op"] == N.OP_DWCONV:
            kh, kw = L["kh"], L["kw"]
            lim = np.sqrt(6.0 / (2 * kh * kw))
            k = rng.uniform(-lim, lim, (kh, kw, cin, 1))
            if L["act"] != N.ACT_NONE:
                k *= 1.6
            w[n + "/depthwise_kernel"] = k.astype(np.float32)
            if L["has_bias"]:
                w[n + "/bias"] = rng.normal(0, 0.05, cin).astype(np.float32)
            if L["has_bn"]:
                _bn_stats(rng, w, L["bn_name"], cin, nontrivial_bn)
        elif L["op"] == N.OP_DECONV:
            k, cout = L["kh"], L["cout"]
            lim = np.sqrt(6.0 / (k * k * cin + k * k * cout))
            w[n + "/kernel"] = rng.uniform(-lim, lim, (k, k, cout, cin)).astype(np.float32)
        elif L["op"] == N.OP_DENSE:
            h, ww, c = g.shapes[L["in0"]]
            fin, cout = h * ww * c, L["cout"]
            lim = np.sqrt(6.0 / (fin + cout))
            k = rng.uniform(-lim, lim, (fin, cout))
            b = rng.normal(0, 0.05, cout)
            if model.kind == "regression" and cout == 136:
                k *= 0.05
                b = canonical_face68().reshape(-1) + rng.normal(0, 0.004, 136)
            w[n + "/kernel"] = k.astype(np.float32)
            w[n + "/bias"] = b.astype(np.float32)
    return w
