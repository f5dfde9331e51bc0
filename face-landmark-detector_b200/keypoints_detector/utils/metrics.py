"""Heat-map -> (x, y) landmark decode — drop-in for reference utils/metrics.py:46-109, computed on the
GPU (fld_decode_heatmap_xy).  RMSE / the n_points sweep (metrics.py:112-154) are evaluation-time and out of
scope."""
import numpy as np
import torch

from .. import _native as N


def heatmap_xy_device(hm, n_points=4, thresh=0.0):
    """hm float32 CUDA [B,H,W,L] -> float64 CUDA [B, 2L] = (x0,y0,x1,y1,...); (-1,-1) where sum/n <= thresh."""
    lib = N.load_library()
    assert hm.is_cuda and hm.dtype == torch.float32 and hm.is_contiguous() and hm.dim() == 4
    B, H, W, L = hm.shape
    if n_points > N.MAX_TOPN:
        raise ValueError("n_points must be <= %d" % N.MAX_TOPN)
    xy = torch.empty((B, 2 * L), dtype=torch.float64, device=hm.device)
    if B == 0:
        return xy
    with torch.cuda.device(hm.device):
        h = N.handle(hm.device)
        # the two-pass decode's partials: a per-call buffer from the stream-aware allocator, so decodes enqueued on
        # different streams never share scratch (the library itself owns none)
        need = int(lib.fld_decode_heatmap_scratch_bytes(h, B, H, W, L, int(n_points)))
        scratch = torch.empty(max(need, 1), dtype=torch.uint8, device=hm.device)
        N.check(lib.fld_decode_heatmap_xy(h, N.ptr(hm), B, H, W, L, int(n_points), float(thresh), N.ptr(xy), N.ptr(scratch),
                                          scratch.numel(), N.stream_ptr(hm.device)))
    return xy


def _to_device(a):
    if not torch.cuda.is_available():
        N.handle()  # raises: CUDA-only
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.device("cuda", torch.cuda.current_device()))


def get_average_xy(hmi, height=96, width=96, n_points=4, thresh=0):
    """reference utils/metrics.py:46-80: one heat-map (H,W) -> [x, y].

    `height` / `width` are accepted and ignored exactly like the reference (its body never reads them for the
    top-n branch and uses them only to build index grids that must match hmi's shape).  Top-n ties, which the
    reference leaves to numpy's unstable argsort, resolve towards the higher flat index."""
    hmi = np.asarray(hmi)
    xy = heatmap_xy_device(_to_device(hmi[None, :, :, None]), n_points, thresh)[0].cpu().numpy()
    x, y = float(xy[0]), float(xy[1])
    if x == -1.0 and y == -1.0:
        return [-1, -1]
    return [x, y]


def transfer_xy_coord(hm, n_points=64, thresh=0.2, reproduce_reference_slip=True):
    """reference utils/metrics.py:83-99: (H,W,L) -> list of 2L coordinates.

    The reference passes (n_points, thresh) positionally into get_average_xy's (height, width) slots
    (metrics.py:98), so its effective decode is ALWAYS n_points=4, thresh=0.  That behaviour is
    reproduced by default for parity; pass reproduce_reference_slip=False to honour the arguments."""
    hm = np.asarray(hm)
    assert len(hm.shape) == 3
    n, t = (4, 0) if reproduce_reference_slip else (n_points, thresh)
    xy = heatmap_xy_device(_to_device(hm[None]), n, t)[0].cpu().numpy()
    return _as_list(xy)


def transfer_target(y_pred, thresh=0, n_points=64, reproduce_reference_slip=True):
    """reference utils/metrics.py:102-109: (N,H,W,L) -> np.array (N, 2L)."""
    y_pred = np.asarray(y_pred)
    n, t = (4, 0) if reproduce_reference_slip else (n_points, thresh)
    return heatmap_xy_device(_to_device(y_pred), n, t).cpu().numpy()


def _as_list(xy):
    return [(-1 if v == -1.0 else float(v)) for v in xy]
