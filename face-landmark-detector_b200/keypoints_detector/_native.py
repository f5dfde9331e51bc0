"""ctypes binding of libfld_sm100.so (include/fld.h).  PyTorch tensors are only the I/O containers:
every call passes raw device pointers and the current CUDA stream.

There is no CPU fallback: if the library or an sm_100 GPU is missing, the first use raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfld_sm100.so")

FLD_U8, FLD_F32, FLD_BF16, FLD_BF16X3 = 0, 1, 2, 3
OP_CONV, OP_DECONV, OP_ADD, OP_DENSE, OP_SOFTMAX, OP_DWCONV, OP_MAXPOOL = range(7)
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
MAX_TOPN = 128


class FldError(RuntimeError):
    pass


class LayerDesc(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("in0", ctypes.c_int32), ("in1", ctypes.c_int32),
                ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("stride", ctypes.c_int32),
                ("pad_t", ctypes.c_int32), ("pad_b", ctypes.c_int32), ("pad_l", ctypes.c_int32), ("pad_r", ctypes.c_int32),
                ("cout", ctypes.c_int32), ("act", ctypes.c_int32), ("pool", ctypes.c_int32),
                ("has_bias", ctypes.c_int32), ("has_bn", ctypes.c_int32), ("in_scale", ctypes.c_float)]


_lib = None
_lock = threading.Lock()
_handles = {}

_vp, _i, _sz, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/fld.h declares
SIGNATURES = {
    "fld_abi_version": (_i, []),
    "fld_last_error": (ctypes.c_char_p, []),
    "fld_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "fld_destroy": (None, [_vp]),
    "fld_preprocess_faces": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "fld_preprocess_faces_staged": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fld_image_array": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "fld_net_create": (_i, [_vp, ctypes.POINTER(LayerDesc), _i, _i, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "fld_net_destroy": (None, [_vp]),
    "fld_net_set_weights": (_i, [_vp, _i, _vp, _vp, _vp, _f]),
    "fld_net_finalize": (_i, [_vp]),
    "fld_net_tensor_shape": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int32)]),
    "fld_net_workspace_bytes": (_sz, [_vp, _i]),
    "fld_net_tensor_offset": (ctypes.c_int64, [_vp, _i, _i]),
    "fld_net_forward": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "fld_net_input_staging": (_i, [_vp, _i, _vp, ctypes.POINTER(_vp)]),
    "fld_net_forward_staged": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "fld_net_forward_classmap": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "fld_net_landmarks_workspace_bytes": (_sz, [_vp, _i, _i]),
    "fld_net_retain": (_i, [_vp]),
    "fld_net_release": (_i, [_vp]),
    "fld_net_forward_landmarks": (_i, [_vp, _vp, _i, _vp, _sz, _i, ctypes.c_double, _vp, _vp]),
    "fld_net_set_profiling": (_i, [_vp, _i]),
    "fld_net_layer_times": (_i, [_vp, _vp, _i]),
    "fld_decode_regress": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "fld_decode_classmap": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "fld_decode_heatmap_scratch_bytes": (_sz, [_vp, _i, _i, _i, _i, _i]),
    "fld_decode_heatmap_xy": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _sz, _vp]),
    "fld_align": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fld_warp_affine": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "fld_align_scratch_bytes": (ctypes.c_size_t, [_vp, _i]),
    "fld_align_ordered": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "fld_warp_affine_ordered": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, ctypes.c_size_t, _vp]),
    "fld_launch_count": (ctypes.c_uint64, []),
}


def load_library():
    """Load libfld_sm100.so (no GPU needed for the load itself)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FldError("%s not found: build it with `python __graft_entry__.py` (or face-landmark-detector_b200/build.py); "
                               "there is no CPU fallback" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise FldError("libfld_sm100: %s (status %d)" % (load_library().fld_last_error().decode(errors="replace"), rc))


def handle(device=None):
    """Per-GPU fld_handle (created once)."""
    lib = load_library()
    if not torch.cuda.is_available():
        raise FldError("no CUDA device visible to PyTorch: the face-landmark hot path is CUDA-only (sm_100a), there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev is None:
        dev = torch.cuda.current_device()
    with _lock:
        h = _handles.get(dev)
        if h is None:
            out = _vp()
            check(lib.fld_create(dev, ctypes.byref(out)))
            h = _handles[dev] = out
    return h


def stream_ptr(device=None):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (must be contiguous) or None."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return _vp(t.data_ptr())


def host_ptr(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return _vp(a.ctypes.data)


def launch_count():
    return int(load_library().fld_launch_count())
