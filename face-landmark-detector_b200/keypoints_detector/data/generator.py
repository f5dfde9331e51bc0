"""Inference-side subset of reference data/generator.py: get_image_array (:29-69), the FCN pre-processor
called at prediction.py:207.  The resize + normalisation run on the GPU (fld_image_array, bit-exact
cv2.resize of OpenCV 4.13 followed by the reference's float32 arithmetic).  The training data pipeline
(dataset pairing, augmentation, generators) is out of scope."""
import os
import re

import numpy as np
import six
import torch

from .. import _native as N
from .config import IMAGE_ORDERING  # noqa: F401

ACCEPTABLE_IMAGE_FORMATS = [".jpg", ".jpeg", ".png", ".bmp"]
ACCEPTABLE_KEYPOINTS_FORMATS = [".pts"]
_NORMS = {"sub_mean": 0, "sub_and_divide": 1, "divide": 2}


class DataLoaderError(Exception):
    pass


def image_array_device(images, width, height, imgNorm="sub_mean"):
    """images uint8 CUDA [B,H,W,3] BGR -> float32 CUDA [B,height,width,3] (channels_last)."""
    lib = N.load_library()
    B, H, W, C = images.shape
    assert C == 3 and images.dtype == torch.uint8 and images.is_contiguous()
    out = torch.empty((B, height, width, 3), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        N.check(lib.fld_image_array(N.handle(images.device), N.ptr(images), B, H, W, width, height, _NORMS[imgNorm], N.ptr(out),
                                    N.stream_ptr(images.device)))
    return out


def get_image_array(image, width, height, imgNorm="sub_mean", ordering='channels_first', read_image_type=1, as_tensor=False):
    """Load image array from input (same signature and defaults as the reference, plus as_tensor)."""
    if isinstance(image, np.ndarray):
        img = image
    elif isinstance(image, six.string_types):
        if not os.path.isfile(image):
            raise DataLoaderError("get_image_array: path {0} doesn't exist".format(image))
        import cv2
        img = cv2.imread(image, read_image_type)
    else:
        raise DataLoaderError("get_image_array: Can't process input type {0}".format(str(type(image))))

    if imgNorm not in _NORMS:
        # the reference falls through and returns the raw image for an unknown imgNorm
        out = img
        if ordering == 'channels_first':
            out = np.rollaxis(out, 2, 0)
        return out
    img = np.atleast_3d(np.ascontiguousarray(img, dtype=np.uint8))
    if img.shape[2] != 3:
        raise DataLoaderError("get_image_array: the CUDA path handles 3-channel images (got %d channels)" % img.shape[2])
    if not torch.cuda.is_available():
        N.handle()  # raises the CUDA-only error
    dev = torch.device("cuda", torch.cuda.current_device())
    out = image_array_device(torch.from_numpy(img).to(dev)[None].contiguous(), width, height, imgNorm)[0]
    if ordering == 'channels_first':
        out = out.permute(2, 0, 1).contiguous()
    if as_tensor:
        return out
    return out.cpu().numpy()


def read_keypoints(keypts_path, is_imgaug_kps=False):
    """Read a .pts landmark file (reference data/generator.py:138-160): `version: V`, `n_points: N`, `{`, one
    "x y" line per point, `}`.  Returns (float array [N,2], n_points, version string)."""
    if is_imgaug_kps:
        raise DataLoaderError("read_keypoints: imgaug Keypoint output belongs to the training pipeline (out of scope here)")
    keypoints, n_points, version = [], None, None
    with open(keypts_path, "r") as fp:
        for line in fp.readlines():
            _text = line.strip()
            if re.match(r"{|}", _text):
                continue
            if re.match("version", _text):
                version = re.findall(r"\d+", _text)[0]
            elif re.match("n_points", _text):
                n_points = int(re.findall(r"\d+", _text)[0])
            else:
                keypoints.append([float(cord) for cord in _text.split()])
    return np.array(keypoints), n_points, version


def write_keypoints(keypts_path, landmarks, version=1):
    """Write decoded landmarks ([N,2] array-like, x y per row) in the reference's .pts format
    (scripts/prepare_dataset.py:46-52), so `detect_marks` / heat-map decode results feed the reference's tooling."""
    pts = np.asarray(landmarks).reshape(-1, 2)
    with open(keypts_path, "w") as fp:
        fp.write("version: %d\n" % version)
        fp.write("n_points: %d\n" % len(pts))
        fp.write("{\n")
        for x, y in pts.tolist():
            fp.write(" ".join([str(x), str(y)]) + "\n")
        fp.write("}")
