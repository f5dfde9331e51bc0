"""Oracle: pre-processing ahead of the CNN.  Test infrastructure only (oracle/__init__.py).

* ``square_box``        reference prediction.py:36-78 (move_box + get_square_box).
* ``resize_linear_u8``  integer restatement of cv2.resize(u8, INTER_LINEAR) as OpenCV
                        4.13 computes it (11-bit coefficients, two-pass), incl. the exact-2x
                        area shortcut.  Pinned against cv2 in tests.
* ``crop_resize_rgb``   reference prediction.py:80-83 (crop, resize to 128x128, BGR->RGB).
* ``get_image_array``   reference data/generator.py:29-69.
"""
import numpy as np


def square_box(face):
    """prediction.py:67-78 then :36-65.  Python int semantics (incl. `%` on negatives)."""
    x0, y0, x1, y1 = [int(v) for v in face]
    offset_y = int(abs((y1 - y0) * 0.1))          # :76
    y0 += offset_y; y1 += offset_y                # :67-74 move_box([0, offset_y])
    bw = x1 - x0; bh = y1 - y0                    # :40-41
    diff = bh - bw                                # :44
    delta = int(abs(diff) / 2)                    # :45
    if diff == 0:
        return [x0, y0, x1, y1]
    elif diff > 0:                                # :50-54
        x0 -= delta; x1 += delta
        if diff % 2 == 1:
            x1 += 1
    else:                                         # :56-60
        y0 -= delta; y1 += delta
        if diff % 2 == 1:                         # true for odd negative diff in Python
            y1 += 1
    assert (x1 - x0) == (y1 - y0)                 # :63
    return [x0, y0, x1, y1]


def _axis_coeffs(src, dst, clamp_frac=True):
    """OpenCV resize.cpp linear coefficient set-up: returns (ofs int32[dst], a0, a1 int32[dst]).

    x axis (clamp_frac=True): at the borders the index is clamped AND the fraction forced to 0.
    y axis (clamp_frac=False): OpenCV keeps the fraction and only clamps the two row indices
    when it fetches rows (so both taps read the same border row with split weights)."""
    scale = 1.0 / (float(dst) / float(src))       # hal::resize: scale_x = 1./inv_scale_x, inv_scale_x = (double)dst/src
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)      # (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int32)
    fr = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        fr = np.where(lo, np.float32(0), fr); s = np.where(lo, 0, s)
        hi = s >= src - 1
        fr = np.where(hi, np.float32(0), fr); s = np.where(hi, src - 1, s)
    a0 = np.rint((np.float32(1.0) - fr) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(fr * np.float32(2048)).astype(np.int32)
    return s.astype(np.int32), a0, a1


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh)) for uint8 HWC, INTER_LINEAR (default)."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    s3 = src.reshape(sh, sw, -1).astype(np.int64)
    if sw == 2 * dw and sh == 2 * dh:
        # OpenCV: INTER_LINEAR with exact 2x decimation is routed to the fast INTER_AREA path
        out = (s3[0::2, 0::2] + s3[0::2, 1::2] + s3[1::2, 0::2] + s3[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape((dh, dw) + src.shape[2:])
    if sw == dw and sh == dh:
        return src.copy()
    xo, xa0, xa1 = _axis_coeffs(sw, dw)
    yo, ya0, ya1 = _axis_coeffs(sh, dh, clamp_frac=False)
    x1 = np.minimum(xo + 1, sw - 1)
    y1 = np.clip(yo + 1, 0, sh - 1)
    yo = np.clip(yo, 0, sh - 1)
    # horizontal pass (int32): T[y, dx] = S[y, xo]*a0 + S[y, xo+1]*a1
    T = s3[:, xo] * xa0[None, :, None] + s3[:, x1] * xa1[None, :, None]
    T0 = T[yo] >> 4
    T1 = T[y1] >> 4
    out = (((ya0[:, None, None] * T0) >> 16) + ((ya1[:, None, None] * T1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def crop_resize_rgb(img, face, size=128):
    """prediction.py:76-83: square box -> crop (numpy slice semantics) -> resize -> BGR2RGB.
    Returns (rgb u8 [size,size,3], facebox).  Raises ValueError where the reference's cv2.resize
    would raise on an empty crop (box leaving the image on the top/left)."""
    fb = square_box(face)
    H, W = img.shape[:2]
    if fb[0] < 0 or fb[1] < 0 or fb[2] <= fb[0] or fb[3] <= fb[1] or fb[0] >= W or fb[1] >= H:
        raise ValueError("empty crop (reference raises a cv2 assertion here)")
    crop = img[fb[1]:fb[3], fb[0]:fb[2]]
    out = resize_linear_u8(crop, size, size)
    return out[:, :, ::-1].copy(), fb


MEANS = [103.939, 116.779, 123.68]


def get_image_array(img, width, height, imgNorm="sub_mean", ordering="channels_last"):
    """data/generator.py:50-69 on an ndarray input."""
    if imgNorm == "sub_and_divide":
        out = np.float32(resize_linear_u8(img, width, height)) / 127.5 - 1
        out = out.astype(np.float32) if out.dtype != np.float32 else out
    elif imgNorm == "sub_mean":
        out = resize_linear_u8(img, width, height).astype(np.float32)
        out = np.atleast_3d(out)
        for i in range(min(out.shape[2], 3)):
            out[:, :, i] -= MEANS[i]
        out = out[:, :, ::-1]
    elif imgNorm == "divide":
        out = resize_linear_u8(img, width, height).astype(np.float32) / 255.0
    else:
        out = img
    if ordering == "channels_first":
        out = np.rollaxis(out, 2, 0)
    return out
