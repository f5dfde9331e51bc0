"""fcn_8/vanilla@224 forward (probabilities) in the three compute modes, batch B: CUDA events."""
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
res = {"batch": B}
for mode in ("bfloat16", "bf16x3", "float32"):
    fn = lambda: m.forward_classmap_device(x, mode)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    m.set_profiling(True, torch.device("cuda", 0), mode); fn(); lt = dict(m.layer_times(torch.device("cuda", 0), mode)); m.set_profiling(False, torch.device("cuda", 0), mode)
    top = sorted(lt.items(), key=lambda kv: -kv[1])[:6]
    res[mode] = {"ms": round(ts[len(ts) // 2], 3), "img_per_s": round(B / ts[len(ts) // 2] * 1e3), "top": {k: round(v, 3) for k, v in top}}
print(json.dumps(res))
