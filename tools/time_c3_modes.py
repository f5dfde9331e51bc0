"""fcn_8/vanilla@224 at batch B: whole-forward time in the three decode modes (probabilities / class map / soft centroid), CUDA events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import torch
import __graft_entry__ as entry
entry.build()
from keypoints_detector.networks.fcn import fcn_8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
x = torch.randn((B, 224, 224, 3), device="cuda") * 50
res = {"batch": B}
for name, fn in (("classmap", lambda: m.forward_classmap_device(x, "bfloat16")),
                 ("centroid", lambda: m.forward_landmarks_device(x, "bfloat16", n_points=0)),
                 ("probs", lambda: m.forward_device(x, "bfloat16"))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res[name] = {"ms_median": round(ts[len(ts) // 2], 3), "ms_min": round(ts[0], 3), "img_per_s": round(B / ts[len(ts) // 2] * 1e3)}
print(json.dumps(res))
