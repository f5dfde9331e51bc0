"""ncu target: one fcn_8/vanilla@224 forward with the fused soft-centroid decode (config C3's path) at a small batch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import __graft_entry__ as entry
entry.build()
from keypoints_detector.networks.fcn import fcn_8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
x = torch.randn((B, 224, 224, 3), device="cuda") * 50
for _ in range(2):
    xy = m.forward_landmarks_device(x, "bfloat16", n_points=0)
torch.cuda.synchronize()
print("ok", xy.shape)
