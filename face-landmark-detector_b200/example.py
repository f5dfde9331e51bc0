"""example.py — BASELINE.json configs[0]: one synthetic face crop, batch 1, random-init landmark CNN,
68-point regression + similarity-aligned crop, through the drop-in API.

(The reference's example.py is a training notebook with hard-coded dataset paths, reference example.py:12-62;
it never reaches the prediction path, so this build ships an example of the hot path instead.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from keypoints_detector import prediction  # noqa: E402
from keypoints_detector.data import synthetic  # noqa: E402
from keypoints_detector.networks.regression import landmark_regressor  # noqa: E402


def main():
    frame = synthetic.make_frames(1, 480, 640, seed=1234)[0]
    face = [200, 120, 400, 360]
    model = landmark_regressor().init_weights(seed=1234)
    marks = prediction.detect_marks(frame, model, face)                      # (68, 2) np.uint, like the reference
    print("landmarks", marks.shape, marks.dtype, marks[:3].tolist())
    marks_f, _ = prediction.detect_marks_batch(frame[None], [face], [0], model)
    crop, M = prediction.align_faces(frame, marks_f, return_matrix=True)    # (1, 112, 112, 3) uint8
    print("aligned crop", crop.shape, crop.dtype, "M =", np.round(M[0], 4).tolist())
    # compute modes: "bf16x3" (default: fp32-accurate tensor cores), "float32" (CUDA cores, the reference's arithmetic),
    # "bfloat16" (plain bf16 tensor cores, fastest)
    for mode in ("float32", "bfloat16"):
        m = prediction.detect_marks_batch(frame[None], [face], [0], model, dtype=mode)[0]
        print("max |%s - bf16x3| landmark delta: %.5f px" % (mode, np.abs(m - marks_f).max()))


if __name__ == "__main__":
    main()
