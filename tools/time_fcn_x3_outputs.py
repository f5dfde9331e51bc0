"""fcn_8/vanilla@224 in the fp32-accurate bf16x3 mode: class map / probabilities / soft-centroid landmarks, batch B (CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import torch
import __graft_entry__ as entry
entry.build()
from keypoints_detector.networks.fcn import fcn_8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
x = torch.randn((B, 224, 224, 3), device="cuda") * 50
res = {"batch": B, "two_pass": not os.environ.get("FLD_X3_DECONV_2PASS_OFF")}
for name, fn in (("classmap", lambda: m.forward_classmap_device(x, "bf16x3")), ("probs", lambda: m.forward_device(x, "bf16x3")),
                 ("centroid", lambda: m.forward_landmarks_device(x, "bf16x3", n_points=0)), ("top4", lambda: m.forward_landmarks_device(x, "bf16x3", n_points=4))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res[name] = {"ms": round(ts[2], 3), "img_per_s": round(B / ts[2] * 1e3)}
print(json.dumps(res))
