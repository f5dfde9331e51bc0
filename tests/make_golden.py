"""Generate tests/golden/reference_host.npz by running THE REFERENCE'S OWN host functions
(/root/reference, unmodified, heavy imports stubbed — SURVEY App. G.1) and OpenCV 4.13 on seeded inputs.

    python tests/make_golden.py          # only works where /root/reference exists (the build container)

The outputs pin the oracle (tests/test_oracle_pinning.py) and, through it, the CUDA path.  The reference
cannot travel to the GPU box, hence the committed fixture.
"""
import hashlib
import os
import sys
from unittest.mock import MagicMock

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import golden_inputs as gi  # noqa: E402

REF = "/root/reference"


def load_reference():
    for name in ["tensorflow", "tensorflow.keras", "tensorflow.keras.callbacks", "tensorflow.keras.losses",
                 "tensorflow.keras.models", "tensorflow.keras.layers", "tensorflow.keras.backend", "tensorflow.keras.utils",
                 "keras", "keras.layers", "keras.models", "matplotlib", "matplotlib.pyplot", "glob2", "imgaug",
                 "imgaug.augmenters", "imgaug.augmentables", "imgaug.augmentables.kps", "imgaug.augmentables.heatmaps",
                 "skimage", "skimage.transform", "skimage.exposure"]:
        sys.modules[name] = MagicMock(name=name)
    tf = sys.modules["tensorflow"]
    tf.uint8 = np.uint8
    tf.constant = lambda a, dtype=None: np.asarray(a, dtype=dtype)
    tf.Tensor = type("T", (), {})
    sys.modules["tensorflow.keras.callbacks"].Callback = type("Callback", (), {})
    sys.path.insert(0, REF)
    from keypoints_detector import prediction
    from keypoints_detector.utils import metrics
    from keypoints_detector.data.generator import get_image_array
    assert prediction.__file__.startswith(REF)
    return prediction, metrics, get_image_array


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    prediction, metrics, get_image_array = load_reference()
    out = {"opencv_version": np.array(cv2.__version__)}

    # A. detect_marks (prediction.py:16-96): marks + the exact uint8 tensor handed to the model
    for i, (seed, h, w, face) in enumerate(gi.DETECT_CASES):
        img = gi.image(seed, h, w)
        seen = {}

        class FakeModel:
            signatures = {"predict": None}

        def predict(x, seed=seed, seen=seen):
            seen["x"] = np.array(x)
            return {"output": gi.fake_outputs(seed)}

        FakeModel.signatures = {"predict": predict}
        marks = prediction.detect_marks(img, FakeModel, list(face))
        assert marks.shape == (68, 2) and marks.dtype == np.uint64
        out["detect_marks_%d" % i] = marks
        out["detect_input_sha_%d" % i] = sha(seen["x"])
        if i < 2:
            out["detect_input_%d" % i] = seen["x"]

    # B. get_image_array (data/generator.py:29-69)
    img = gi.image(21, 45, 60)
    for norm in ("sub_mean", "sub_and_divide", "divide"):
        a = get_image_array(img, 48, 32, imgNorm=norm, ordering="channels_last")
        out["image_array_" + norm] = np.ascontiguousarray(a)
    out["image_array_cf"] = np.ascontiguousarray(get_image_array(img, 48, 32, ordering="channels_first"))

    # C. _prediction (prediction.py:199-222) with a fake .predict
    for i, (oh, ow, n) in enumerate([(12, 12, 5), (9, 14, 68)]):
        p = gi.probs(31 + i, oh * ow, n)

        class M:
            pass

        m = M()
        m.predict = lambda x, p=p: p
        inp = gi.image(32 + i, 40, 52)
        pr = prediction._prediction(m, inp, 32, 32, oh, ow, n, [(1, 2, 3)] * 80, False, None, (ow, oh), False, None)
        assert pr.dtype == np.int64 and pr.shape == (oh, ow)
        out["class_map_%d" % i] = pr

    # D. utils/metrics.py:46-109
    hm = gi.heatmaps(41, 2, 24, 20, 3)
    cases = [(4, 0), (1, 0), (9, 0), (4, 0.3), (0, 0), (0, 0.01)]
    res = np.zeros((len(cases), 2, 3, 2))
    for ci, (npnt, th) in enumerate(cases):
        for b in range(2):
            for l in range(3):
                res[ci, b, l] = metrics.get_average_xy(hm[b, :, :, l], 24, 20, npnt, th)
    out["average_xy_cases"] = np.array(cases, dtype=np.float64)
    out["average_xy"] = res
    out["transfer_xy_coord"] = np.array(metrics.transfer_xy_coord(hm[0], n_points=9, thresh=0.5), dtype=np.float64)
    out["transfer_target"] = np.array(metrics.transfer_target(hm, thresh=0.4, n_points=16), dtype=np.float64)

    # E. OpenCV 4.13 integer schemes
    for i, (sh, sw, dh, dw) in enumerate(gi.RESIZE_SHAPES):
        out["resize_sha_%d" % i] = sha(cv2.resize(gi.image(50 + i, sh, sw), (dw, dh)))
    frame = gi.image(60, 1080, 1920)
    for i in range(6):
        M = gi.similarity(i)
        crop = cv2.warpAffine(frame, M, (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        out["warp_sha_%d" % i] = sha(crop)
        if i == 0:
            out["warp_crop_0"] = crop

    # F. .pts reader (data/generator.py:138-160) on a file in the reference writer's format (scripts/prepare_dataset.py:46-52)
    import tempfile
    from keypoints_detector.data.generator import read_keypoints
    with tempfile.NamedTemporaryFile("w", suffix=".pts", delete=False) as fp:
        fp.write(gi.PTS_TEXT)
    kps, n_points, version = read_keypoints(fp.name)
    os.unlink(fp.name)
    out["pts_keypoints"] = np.asarray(kps, dtype=np.float64)
    out["pts_n_points"] = np.array(n_points)
    out["pts_version"] = np.array(str(version))

    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    path = os.path.join(HERE, "golden", "reference_host.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
