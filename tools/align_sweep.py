"""Alignment kernel sweeps: run C4 and single-scale bands under the FLD_ALIGN_* environment switches given on the command line
(e.g. `python tools/align_sweep.py none FLD_ALIGN_YSPLIT=2 FLD_ALIGN_RING_KB=24,FLD_ALIGN_PAIR_MAX=0`)."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from keypoints_detector import prediction
from keypoints_detector.data import synthetic
from bench_kernels import timeit
dev = torch.device("cuda", 0)
F, B = 64, 4096
frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, device=dev)
f2f = torch.from_numpy((np.arange(B) // 64).astype(np.int32)).to(dev)
res = {}
for tag, kw in (("C4", {}), ("s1.0-1.4", dict(scale=(1.0, 1.4))), ("s0.5-0.7", dict(scale=(0.5, 0.7))), ("s0.4-0.5", dict(scale=(0.4, 0.5))), ("s0.33-0.4", dict(scale=(0.33, 0.4))), ("s0.28-0.33", dict(scale=(0.28, 0.33)))):
    pts, Ms = synthetic.make_similarity_landmarks(B, 1080, 1920, prediction.TEMPLATE_112, seed=4, **kw)
    marks = torch.from_numpy(pts).to(dev)
    out = torch.empty((B, 112, 112, 3), dtype=torch.uint8, device=dev); M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
    med, mn = timeit(lambda: prediction.align_device(frames, f2f, marks, None, (112, 112), five_point=False, out=out, out_matrix=M))
    res[tag] = round(med, 4)
print(json.dumps(res))
''' % ROOT
for env in sys.argv[1:]:
    e = dict(os.environ); 
    for kv in env.split(","):
        if "=" in kv: k, v = kv.split("="); e[k] = v
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
