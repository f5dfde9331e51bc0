"""Fine-grained clock64 trace of the producer / MMA loops (FLD_TC_TRACE): median gaps between consecutive tags."""
import glob, os, sys, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
out = os.path.join(ROOT, "gpurun_out", "trace"); os.makedirs(out, exist_ok=True)
for f in glob.glob(out + "/trace_*.txt"): os.remove(f)
os.environ["FLD_TC_TRACE"] = out
import __graft_entry__ as e; e.build()
from keypoints_detector.networks.regression import landmark_regressor
m = landmark_regressor().init_weights(0)
x = torch.randint(0, 256, (256, 128, 128, 3), dtype=torch.uint8, device="cuda")
for _ in range(2): m.forward_device(x, "bfloat16")
torch.cuda.synchronize()
for f in sorted(glob.glob(out + "/trace_*.txt"))[-4:-2]:
    ev = np.loadtxt(f, dtype=np.int64).reshape(-1, 3)
    print("==", os.path.basename(f))
    for role, name in ((0, "producer"), (1, "mma")):
        r = ev[ev[:, 0] == role]
        gaps = collections.defaultdict(list)
        for (a, b) in zip(r[40:-1], r[41:]):      # skip the pipeline fill
            gaps[(int(a[1]), int(b[1]))].append(int(b[2] - a[2]))
        print(" ", name, {k: int(np.median(v)) for k, v in sorted(gaps.items()) if len(v) > 5})
