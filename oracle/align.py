"""Oracle: similarity alignment (build-defined; NOT in the reference — SURVEY §0, App. B).

Test infrastructure only (see oracle/__init__.py).

* ``umeyama``          fp64 closed-form 2-D least-squares similarity (the algorithm
                       skimage.transform.SimilarityTransform.estimate implements;
                       skimage is a reference dependency, data/generator.py:17).
* ``umeyama_svd``      independent SVD formulation used to cross-check the closed form.
* ``warp_affine_u8``   integer restatement of cv2.warpAffine(u8, INTER_LINEAR,
                       BORDER_CONSTANT 0) as OpenCV 4.13 computes it (1/32-px
                       coordinates, 15-bit weights).  Pinned against cv2 in tests.
* ``five_points``      iBUG-68 -> 5-point reduction used by prediction.align_faces.
"""
import numpy as np

# widely used 112x112 five-point template
TEMPLATE_112 = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                         [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float64)


def five_points(marks68):
    """marks68 [...,68,2] -> [...,5,2]: eye centres (36..41, 42..47), nose tip 30, mouth corners 48, 54."""
    m = np.asarray(marks68, dtype=np.float64)
    le = m[..., 36:42, :].sum(axis=-2) / 6.0
    re = m[..., 42:48, :].sum(axis=-2) / 6.0
    return np.stack([le, re, m[..., 30, :], m[..., 48, :], m[..., 54, :]], axis=-2)


def _tree_sum(vals):
    """Summation tree of the CUDA fit (csrc/align.cu fit_similarity_warp): lane l of a 32-lane warp adds elements
    l, l+32, ... in index order, then an xor butterfly (16, 8, 4, 2, 1) combines the lanes.  IEEE addition is
    commutative, so every lane ends with the same value; lane 0's is returned."""
    p = [0.0] * 32
    for i, v in enumerate(vals):
        p[i % 32] = p[i % 32] + float(v)
    for off in (16, 8, 4, 2, 1):
        p = [p[l] + p[l ^ off] for l in range(32)]
    return p[0]


def umeyama(src, dst):
    """Closed-form 2-D similarity fit src->dst.  Returns M [2,3] fp64 (NaN if degenerate).

    Every sum uses `_tree_sum`, the fixed warp-parallel order of the CUDA kernel, and no fused multiply-add, so the
    fp64 results agree bit for bit."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    n = src.shape[0]
    px_, py_ = [float(v) for v in src[:, 0]], [float(v) for v in src[:, 1]]
    qx_, qy_ = [float(v) for v in dst[:, 0]], [float(v) for v in dst[:, 1]]
    mpx = _tree_sum(px_) / n; mpy = _tree_sum(py_) / n
    mqx = _tree_sum(qx_) / n; mqy = _tree_sum(qy_) / n
    dx = [v - mpx for v in px_]; dy = [v - mpy for v in py_]
    ex = [v - mqx for v in qx_]; ey = [v - mqy for v in qy_]
    var = _tree_sum([x * x + y * y for x, y in zip(dx, dy)])
    a = _tree_sum([q * x for q, x in zip(ex, dx)])
    b = _tree_sum([q * y for q, y in zip(ex, dy)])
    c = _tree_sum([q * x for q, x in zip(ey, dx)])
    d = _tree_sum([q * y for q, y in zip(ey, dy)])
    P = a + d
    Q = c - b
    if var == 0.0 or (P == 0.0 and Q == 0.0) or not np.isfinite(var) or not np.isfinite(P) or not np.isfinite(Q):
        return np.full((2, 3), np.nan)
    # (1/n) factors of covariance and variance cancel
    l00 = P / var; l01 = -Q / var
    l10 = Q / var; l11 = P / var
    t0 = mqx - (l00 * mpx + l01 * mpy)
    t1 = mqy - (l10 * mpx + l11 * mpy)
    return np.array([[l00, l01, t0], [l10, l11, t1]], dtype=np.float64)


def umeyama_svd(src, dst):
    """Umeyama 1991 via SVD with reflection handling (cross-check only)."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    n = src.shape[0]
    mp = src.mean(0); mq = dst.mean(0)
    sp = src - mp; sq = dst - mq
    A = sq.T @ sp / n
    d = np.ones(2)
    if np.linalg.det(A) < 0:
        d[1] = -1
    U, S, Vt = np.linalg.svd(A)
    R = U @ np.diag(d) @ Vt
    var = (sp ** 2).sum() / n
    scale = (S * d).sum() / var
    L = scale * R
    t = mq - L @ mp
    return np.concatenate([L, t[:, None]], axis=1)


def invert_affine(M):
    """cv2.invertAffineTransform in fp64."""
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    i00 = M[1, 1] * D; i01 = -M[0, 1] * D
    i10 = -M[1, 0] * D; i11 = M[0, 0] * D
    i02 = -i00 * M[0, 2] - i01 * M[1, 2]
    i12 = -i10 * M[0, 2] - i11 * M[1, 2]
    return np.array([[i00, i01, i02], [i10, i11, i12]], dtype=np.float64)


def bilinear_tab():
    """32x32 table of 2x2 int16 weights, sum fixed to 32768 (OpenCV initInterTab2D, INTER_LINEAR)."""
    tab = np.zeros((32, 32, 2, 2), dtype=np.int32)
    for fy in range(32):
        wy = np.array([1.0 - fy / 32.0, fy / 32.0], dtype=np.float32)
        for fx in range(32):
            wx = np.array([1.0 - fx / 32.0, fx / 32.0], dtype=np.float32)
            w = (wy[:, None] * wx[None, :]).astype(np.float32)
            iw = np.rint(w * np.float32(32768)).astype(np.int32)
            s = int(iw.sum())
            if s != 32768:
                diff = s - 32768
                flat = iw.reshape(-1)
                if diff < 0:
                    # add the deficit to the max element (OpenCV: ksize2.. search mk for max)
                    k = int(np.argmax(flat))
                    flat[k] -= diff
                else:
                    k = int(np.argmin(flat))
                    flat[k] -= diff
                iw = flat.reshape(2, 2)
            tab[fy, fx] = iw
    return tab


_TAB = None


def warp_affine_u8(src, M, out_w, out_h):
    """Integer model of cv2.warpAffine(src, M, (out_w,out_h), INTER_LINEAR, BORDER_CONSTANT, 0).

    src uint8 [H,W,C]; M [2,3] forward (src->dst) transform, inverted here in fp64."""
    global _TAB
    if _TAB is None:
        _TAB = bilinear_tab()
    src = np.asarray(src)
    H, W = src.shape[:2]
    C = src.shape[2] if src.ndim == 3 else 1
    s3 = src.reshape(H, W, C).astype(np.int64)
    iM = invert_affine(M)
    AB = 1024.0
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(iM[0, 0] * xs * AB).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * AB).astype(np.int64)
    X0 = np.rint((iM[0, 1] * ys + iM[0, 2]) * AB).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * ys + iM[1, 2]) * AB).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    # OpenCV stores sx, sy as saturate_cast<short>
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx = X & 31
    fy = Y & 31
    w = _TAB[fy, fx]  # [h,w,2,2]
    acc = np.zeros((out_h, out_w, C), dtype=np.int64)
    for ky in range(2):
        for kx in range(2):
            yy = sy + ky
            xx = sx + kx
            inb = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            pix = s3[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
            pix = np.where(inb[..., None], pix, 0)
            acc += pix * w[..., ky, kx][..., None]
    out = (acc + 16384) >> 15
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[..., 0]


def align_faces(frames, face2frame, marks, template=TEMPLATE_112, out_hw=(112, 112), mode5=True):
    """Oracle for prediction.align_faces: per-face fit + warp.  Returns (M [B,2,3] f64, crops u8)."""
    marks = np.asarray(marks, dtype=np.float64)
    B = marks.shape[0]
    oh, ow = out_hw
    Ms = np.zeros((B, 2, 3), dtype=np.float64)
    C = frames[0].shape[2]
    crops = np.zeros((B, oh, ow, C), dtype=np.uint8)
    for i in range(B):
        pts = five_points(marks[i]) if mode5 else marks[i]
        M = umeyama(pts, template)
        Ms[i] = M
        if np.isfinite(M).all():
            crops[i] = warp_affine_u8(frames[int(face2frame[i])], M, ow, oh)
    return Ms, crops
