import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import numpy as np, torch
from keypoints_detector import prediction
from keypoints_detector.data import synthetic
dev = torch.device("cuda", 0)
B = int(sys.argv[1]); F = max(1, B // 64)
frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, device=dev)
f2f = torch.from_numpy((np.arange(B) // 64).astype(np.int32) % F).to(dev)
pts, Ms = synthetic.make_similarity_landmarks(B, 1080, 1920, prediction.TEMPLATE_112, seed=4)
marks = torch.from_numpy(pts).to(dev)
out = torch.empty((B, 112, 112, 3), dtype=torch.uint8, device=dev); M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
fn = lambda: prediction.align_device(frames, f2f, marks, None, (112, 112), five_point=False, out=out, out_matrix=M)
for _ in range(3): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fn()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 20 * 1e3)
print(B, os.environ.get("FLD_ALIGN_YSPLIT", "default"), "us per call: %.1f" % sorted(ts)[len(ts) // 2])
