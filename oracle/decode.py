"""Oracle: landmark decodes.  Test infrastructure only (oracle/__init__.py).

PINNED: each function here is checked against the reference's own function run from
/root/reference under stubbed imports (tests/make_golden.py, tests/test_oracle_pinning.py).

* ``regression_decode``  prediction.py:88-94
* ``class_map``          prediction.py:209
* ``average_xy``         utils/metrics.py:46-80   (get_average_xy)
* ``transfer_xy_coord``  utils/metrics.py:83-99   (incl. the positional-argument slip at :98)
* ``transfer_target``    utils/metrics.py:102-109
"""
import numpy as np


def regression_decode(out, facebox):
    """out [>=136] normalised (x0,y0,x1,...) -> (marks float32 [68,2] pre-cast, marks uint64 [68,2]).

    prediction.py:88-93 works in the network's float32; :94 casts with astype(np.uint)
    (truncation toward zero; negatives are clamped to 0 here)."""
    m = np.array(out, dtype=np.float32).flatten()[:136].reshape(-1, 2).copy()
    side = facebox[2] - facebox[0]
    m *= side
    m[:, 0] += facebox[0]
    m[:, 1] += facebox[1]
    return m, np.maximum(m, 0).astype(np.uint64)


def class_map(probs, oh, ow, n_classes):
    """prediction.py:209: per-pixel argmax over classes, first max wins; int64 [oh,ow]."""
    return np.asarray(probs).reshape(oh, ow, n_classes).argmax(axis=2)


def average_xy(hmi, n_points=4, thresh=0.0):
    """utils/metrics.py:46-80 on one heat-map [H,W] -> [x, y] (python floats, float64 accumulation).

    Top-n ties: the reference uses numpy's unstable argsort (:66) so ties are unspecified;
    this restatement (and the CUDA kernel) prefer the HIGHER flat index among equal values
    (what a stable ascending sort followed by [-n:] yields)."""
    hmi = np.asarray(hmi)
    H, W = hmi.shape
    if n_points < 1:
        hsum = float(np.sum(hmi.astype(np.float64)))
        n = H * W
        i1 = float(np.sum(np.arange(W)[None, :] * hmi.astype(np.float64))) / hsum      # :60-61
        i0 = float(np.sum(np.arange(H)[:, None] * hmi.astype(np.float64))) / hsum      # :62-63
    else:
        n = n_points
        flat = hmi.reshape(-1)
        ind = np.argsort(flat, kind="stable")[-n_points:]                              # :66
        # :68-77 with the reference's scalar dtypes: `hsum` accumulates in the heat-map's dtype
        # (int 0 + np.float32 -> np.float32), i0/i1 in float64 (np.int64 * np.float32 -> float64)
        i0, i1, hsum = 0, 0, 0
        for k in ind:                                                                   # :70-74 ascending order
            h = flat[k]
            r, c = np.int64(int(k) // W), np.int64(int(k) % W)
            hsum += h
            i0 += r * h
            i1 += c * h
        i0 /= hsum
        i1 /= hsum
        i0, i1, hsum = float(i0), float(i1), float(hsum)
    if hsum / n <= thresh:                                                              # :78-79
        i0, i1 = -1, -1
    return [i1, i0]


def transfer_xy_coord(hm, n_points=64, thresh=0.2, reproduce_slip=True):
    """utils/metrics.py:83-99.  With reproduce_slip the (n_points, thresh) arguments land in
    get_average_xy's unused (height, width) slots (:98), so the effective decode is n_points=4,
    thresh=0 regardless of what the caller passed."""
    hm = np.asarray(hm)
    assert hm.ndim == 3
    out = []
    for i in range(hm.shape[-1]):
        if reproduce_slip:
            out.extend(average_xy(hm[:, :, i], 4, 0))
        else:
            out.extend(average_xy(hm[:, :, i], n_points, thresh))
    return out


def transfer_target(y_pred, thresh=0, n_points=64, reproduce_slip=True):
    """utils/metrics.py:102-109: [N,H,W,L] -> [N, 2L]."""
    return np.array([transfer_xy_coord(y_pred[i], n_points, thresh, reproduce_slip)
                     for i in range(y_pred.shape[0])])
