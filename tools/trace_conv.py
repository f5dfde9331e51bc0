"""Run one bf16 forward of the regression net with FLD_TC_TRACE and summarise CTA 0's event log per conv layer."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
out = os.path.join(ROOT, "gpurun_out", "trace"); os.makedirs(out, exist_ok=True)
os.environ["FLD_TC_TRACE"] = out
import __graft_entry__ as e; e.build()
from keypoints_detector.networks.regression import landmark_regressor
m = landmark_regressor().init_weights(0)
x = torch.randint(0, 256, (256, 128, 128, 3), dtype=torch.uint8, device="cuda")
for _ in range(2): m.forward_device(x, "bfloat16")
torch.cuda.synchronize()
for f in sorted(glob.glob(out + "/trace_*.txt"))[-4:]:
    ev = np.loadtxt(f, dtype=np.int64).reshape(-1, 3)
    t0 = ev[:, 2].min()
    print("==", os.path.basename(f))
    for role, name in ((0, "producer"), (1, "mma"), (2, "epilogue")):
        r = ev[ev[:, 0] == role]
        tags = r[:, 1]; t = r[:, 2] - t0
        print(" %s: %d events, span %d cycles" % (name, len(r), t.max() - t.min() if len(t) else 0))
        if role == 0:
            w = t[tags == 1]; i = t[tags == 2]
            n = min(len(w), len(i))
            print("   wait-empty->issued (issue cost) median %d; issued->next wait-return median %d; k-block period median %d"
                  % (np.median(i[:n] - w[:n]), np.median(w[1:n] - i[:n - 1]), np.median(np.diff(w))))
            print("   first 14 wait-return times:", w[:14].tolist())
        if role == 1:
            w = t[tags == 1]; c = t[tags == 2]; n = min(len(w), len(c))
            print("   wait-full->commit median %d; commit->next wait-return median %d; k-block period median %d"
                  % (np.median(c[:n] - w[:n]), np.median(w[1:n] - c[:n - 1]), np.median(np.diff(w))))
            ts = t[tags == 0]; te = t[tags == 3]; tf = t[tags == 4]
            print("   tile start->tempty wait cost median %d; tile period median %d" % ((np.median(te[:len(ts)] - ts[:len(te)]) if len(ts) and len(te) else -1), (np.median(np.diff(ts)) if len(ts) > 1 else -1)))
            print("   first 14 wait-full return:", w[:14].tolist())
        if role == 2:
            w = t[tags == 1]; a = t[tags == 2]; s = t[tags == 0]; n = min(len(w), len(a))
            print("   tfull-wait return->arrive (epilogue busy) median %d; tile period median %d; idle (arrive->next tfull) median %d"
                  % (np.median(a[:n] - w[:n]), (np.median(np.diff(w)) if len(w) > 1 else -1), (np.median(w[1:n] - a[:n - 1]) if n > 1 else -1)))
