"""Small end-to-end target for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once, tiny sizes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from keypoints_detector import prediction  # noqa: E402
from keypoints_detector.data import synthetic  # noqa: E402
from keypoints_detector.networks import fcn  # noqa: E402
from keypoints_detector.networks.regression import landmark_regressor  # noqa: E402
from keypoints_detector.utils import metrics  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
which = sys.argv[1:] or ["pipeline", "fcn", "encoders", "align"]
if "pipeline" in which:
    frames = torch.from_numpy(synthetic.make_frames(2, 240, 320, seed=1)).to(dev)
    boxes = torch.from_numpy(synthetic.make_boxes(5, 240, 320, seed=2, max_side=150)).to(dev)
    f2f = torch.from_numpy((np.arange(5) % 2).astype(np.int32)).to(dev)
    m = landmark_regressor().init_weights(seed=0)
    for dt in ("bfloat16", "float32"):
        r = prediction.LandmarkPipeline(m, dtype=dt, device=dev).run_device(frames, boxes, f2f, want_uint=True)
        torch.cuda.synchronize()
        print("pipeline", dt, float(r["marks"].abs().max()))
if "fcn" in which:
    m = fcn.fcn_8(68, input_height=64, input_width=96).init_weights(1)
    x = torch.randn((2, 64, 96, 3), device=dev) * 40
    for dt in ("bfloat16", "float32"):
        p = m.forward_device(x, dt)
        c = m.forward_classmap_device(x, dt)
        xy = m.forward_landmarks_device(x, dt)
        t4 = metrics.heatmap_xy_device(p.view(2, m.output_height, m.output_width, 68).contiguous(), 4, 0.0)
        torch.cuda.synchronize()
        print("fcn", dt, float(p.sum()), int(c.max()), float(xy.max()), float(t4.max()))
    m32 = fcn.fcn_32(12, input_height=64, input_width=64).init_weights(2)
    print("fcn32", float(m32.forward_device(torch.randn((1, 64, 64, 3), device=dev), "bfloat16").sum()))
if "encoders" in which:
    for build in (fcn.fcn_8_mobilenet, fcn.fcn_8_resnet50, fcn.fcn_8_vgg):
        m = build(68, 64, 64).init_weights(3)
        p = m.forward_device(torch.randn((1, 64, 64, 3), device=dev) * 40, "bfloat16")
        torch.cuda.synchronize()
        print(m.model_name, float(p.sum()))
if "align" in which:
    # ordered mode (fit kernel + counting sort + tile kernel with warp pairs) on a few faces of every box class
    os.environ["FLD_ALIGN_ORDER_MIN"] = "8"
    prediction.ALIGN_ORDER_MIN_FACES = 8
    frames = torch.from_numpy(synthetic.make_frames(2, 1080, 1920, seed=1)).to(dev)
    pts, _ = synthetic.make_similarity_landmarks(24, 1080, 1920, prediction.TEMPLATE_112, seed=5, scale=(0.2, 1.4))
    f2f = torch.from_numpy((np.arange(24) % 2).astype(np.int32)).to(dev)
    a, M = prediction.align_device(frames, f2f, torch.from_numpy(pts).to(dev), None, (112, 112), five_point=False)
    w = prediction.warp_affine_device(frames, f2f, torch.nan_to_num(M), (112, 112))
    torch.cuda.synchronize()
    print("align ordered", int(a.sum()), bool(torch.equal(a, w)))
print("sanitize target done")
