"""ncu target: one fcn_8/vanilla@224 bf16 forward in probability mode, then one in class-map mode (B from argv, default 64)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from keypoints_detector.networks.fcn import fcn_8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
x = torch.randn((B, 224, 224, 3), dtype=torch.float32, device=dev) * 50
for _ in range(2):
    p = m.forward_device(x, "bfloat16")
    c = m.forward_classmap_device(x, "bfloat16")
torch.cuda.synchronize()
print("ok", tuple(p.shape), tuple(c.shape))
