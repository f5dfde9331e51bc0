"""Config C5 on one GPU: the regression pipeline (crop/resize -> trunk + FC -> decode -> fit + warp) over batch sizes
1 ... 65536, eager and (small batches) as a replayed CUDA graph.  One JSON line per batch size."""
import json
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
from keypoints_detector.networks.regression import landmark_regressor  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dtype = sys.argv[1] if len(sys.argv) > 1 else "bfloat16"
sizes = [int(a) for a in sys.argv[2:]] or [1, 4, 16, 64, 256, 1024, 4096, 16384, 65536]
model = landmark_regressor().init_weights(seed=0)
pipe = prediction.LandmarkPipeline(model, dtype=dtype, device=dev)
F = 64
frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in sizes:
    boxes = torch.from_numpy(synthetic.make_boxes(B, 1080, 1920, seed=B)).to(dev)
    f2f = torch.from_numpy((np.arange(B) // max(1, -(-B // F)) % F).astype(np.int32)).to(dev)
    reps = max(3, min(200, 20000 // max(B, 1)))
    ms = timed(lambda: pipe.run_device(frames, boxes, f2f), reps)
    line = {"config": "C5 sweep", "dtype": dtype, "batch": B, "ms_eager": ms, "faces_per_s_eager": B / ms * 1e3}
    if B <= 1024:
        g, res = pipe.capture(frames, boxes, f2f)
        eager = {k: v.clone() for k, v in pipe.run_device(frames, boxes, f2f).items() if v is not None}
        msg = timed(g.replay, reps)
        same = all(torch.equal(eager[k], res[k]) for k in eager)
        line.update({"ms_graph": msg, "faces_per_s_graph": B / msg * 1e3, "graph_equals_eager": bool(same)})
    print(json.dumps(line), flush=True)
    del boxes, f2f
    pipe._bufs.clear()
    torch.cuda.empty_cache()
