"""Clock trace of the up8 transposed-conv kernel (FLD_TC_TRACE): class-map and soft-centroid modes, CTA 0."""
import os, sys, glob, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
out = os.path.join(ROOT, "gpurun_out", "trace"); os.makedirs(out, exist_ok=True)
os.environ["FLD_TC_TRACE"] = out
import torch
import __graft_entry__ as entry
entry.build()
from keypoints_detector.networks.fcn import fcn_8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
x = torch.randn((B, 224, 224, 3), device="cuda") * 50
m.forward_classmap_device(x, "bfloat16"); m.forward_classmap_device(x, "bfloat16")
m.forward_landmarks_device(x, "bfloat16", n_points=0); m.forward_landmarks_device(x, "bfloat16", n_points=0)
torch.cuda.synchronize()
for f in sorted(glob.glob(out + "/trace_deconv_*_s8_*.txt")):
    ev = collections.defaultdict(list)
    for line in open(f):
        r, tag, clk = line.split(); ev[int(r)].append((int(tag), int(clk)))
    print("==", os.path.basename(f))
    # MMA warp: time waiting for weights (prev event -> tag1), for tempty (-> tag3), issuing (tag1 -> tag2)
    mm = ev[1]; w_full = w_empty = issue = 0; n_t = 0
    for (t0, c0), (t1, c1) in zip(mm, mm[1:]):
        if t1 == 1: w_full += c1 - c0
        elif t1 == 3: w_empty += c1 - c0; n_t += 1
        elif t1 == 2: issue += c1 - c0
    span = mm[-1][1] - mm[0][1]
    print(f" mma warp: span {span} clk over {n_t} tiles = {span / max(n_t,1):.0f}/tile; wait weights {w_full / max(n_t,1):.0f}, wait acc-empty {w_empty / max(n_t,1):.0f}, issue {issue / max(n_t,1):.0f} per tile")
    ep = ev[2]; w_tfull = ld = rest = 0; n_e = 0
    for (t0, c0), (t1, c1) in zip(ep, ep[1:]):
        if t1 == 1: w_tfull += c1 - c0; n_e += 1
        elif t1 == 2: ld += c1 - c0
        elif t1 == 0: rest += c1 - c0
    print(f" epilogue warp 2: {n_e} tile-phases; wait tfull {w_tfull / max(n_e,1):.0f}, tmem load {ld / max(n_e,1):.0f}, math+store {rest / max(n_e,1):.0f} clk per tile-phase")
    pr = ev[0]; w_slot = 0; n_p = 0
    for (t0, c0), (t1, c1) in zip(pr, pr[1:]):
        if t1 == 1: w_slot += c1 - c0; n_p += 1
    print(f" producer: wait slot {w_slot / max(n_p,1):.0f} clk per k-block over {n_p}")
