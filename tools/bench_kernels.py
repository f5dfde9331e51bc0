#!/usr/bin/env python
"""Per-kernel microbenchmarks at the BASELINE.json config sizes (C3 decode, C4 align, pre-processing, FCN
forward), CUDA-event timed, inputs larger than L2 or rotated.  Writes one JSON object per line to stdout.
These are diagnostic numbers for profiles/ — bench.py stays the contract benchmark."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "face-landmark-detector_b200"))
import __graft_entry__ as entry  # noqa: E402

HBM = 6551.7
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps=20, warm=3, inner=4):
    """Median / min time of one call.  `inner` calls are enqueued back to back between the two events so that the Python /
    ctypes launch path (tens of microseconds) is hidden behind the previous call's kernels instead of being timed."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts)), float(np.min(ts))


def main():
    entry.build()
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    from keypoints_detector.utils import metrics
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    which = sys.argv[1:] or ["align", "decode", "preprocess", "fcn", "encoders"]

    if "align" in which:
        F, B = 64, 4096
        frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, device=dev)
        pts, Ms = synthetic.make_similarity_landmarks(B, 1080, 1920, prediction.TEMPLATE_112, seed=4)
        f2f = torch.from_numpy((np.arange(B) // 64).astype(np.int32)).to(dev)
        marks = torch.from_numpy(pts).to(dev)
        med, mn = timeit(lambda: prediction.align_device(frames, f2f, marks, None, (112, 112), five_point=False))
        s2 = np.array([np.linalg.det(M[:, :2]) for M in Ms])            # scale^2 per face
        foot = float((112 * 112 * 3 / s2).sum())                           # unique source footprint bytes (unclipped)
        alg = B * 37632 + foot + B * (5 * 2 * 4 + 48)
        print(json.dumps({"kernel": "align_warp_kernel (C4: 4096 faces, 64 x 1080p frames -> 112x112x3)", "ms_median": med, "ms_min": mn,
                          "faces_per_s": B / med * 1e3, "algorithmic_bytes": alg, "bytes_per_face": alg / B,
                          "achieved_GBs": alg / med / 1e6, "peak_GBs": HBM, "frac_of_measured_hbm": alg / med / 1e6 / HBM}))

    if "preprocess" in which:
        F, B = 64, 4096
        frames = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8, device=dev)
        boxes = torch.from_numpy(synthetic.make_boxes(B, 1080, 1920, seed=3)).to(dev)
        f2f = torch.from_numpy((np.arange(B) // 64).astype(np.int32)).to(dev)
        med, mn = timeit(lambda: prediction.preprocess_faces_device(frames, boxes, f2f))
        side = (boxes[:, 2] - boxes[:, 0]).float().cpu().numpy()
        alg = float((np.maximum(side, boxes[:, 3].cpu().numpy() - boxes[:, 1].cpu().numpy()) ** 2 * 3).sum()) + B * 49152
        print(json.dumps({"kernel": "resize_kernel<faces> (4096 faces -> 128x128x3)", "ms_median": med, "faces_per_s": B / med * 1e3,
                          "algorithmic_bytes": alg, "achieved_GBs": alg / med / 1e6, "frac_of_measured_hbm": alg / med / 1e6 / HBM}))

    if "decode" in which:
        B, oh, ow, L = 64, 232, 232, 68
        logits = torch.randn((B, oh * ow, L), dtype=torch.float32, device=dev)
        nbytes = logits.numel() * 4
        med, _ = timeit(lambda: prediction.class_map_device(logits, oh, ow))
        print(json.dumps({"kernel": "classmap_kernel (C3 shape, B=64)", "ms_median": med, "images_per_s": B / med * 1e3,
                          "algorithmic_bytes": nbytes + B * oh * ow * 8, "achieved_GBs": (nbytes + B * oh * ow * 8) / med / 1e6,
                          "frac_of_measured_hbm": (nbytes + B * oh * ow * 8) / med / 1e6 / HBM}))
        hm = logits.view(B, oh, ow, L)
        for n in (4, 0):
            med, _ = timeit(lambda: metrics.heatmap_xy_device(hm, n, 0.0))
            print(json.dumps({"kernel": "heatmap_xy (n_points=%d, C3 shape, B=64)" % n, "ms_median": med, "images_per_s": B / med * 1e3,
                              "algorithmic_bytes": nbytes, "achieved_GBs": nbytes / med / 1e6, "frac_of_measured_hbm": nbytes / med / 1e6 / HBM}))

    def fcn_bench(label, m, dtype, B, H=224, W=224):
        x = torch.randn((B, H, W, 3), dtype=torch.float32, device=dev) * 50
        m.forward_device(x, dtype)
        m.set_profiling(True, dev, dtype)
        acc = np.zeros(len(m.graph.layers))
        for _ in range(3):
            m.forward_device(x, dtype)
            acc += np.array([t for _, t in m.layer_times(dev, dtype)])
        accc = np.zeros(len(m.graph.layers))
        for _ in range(3):
            m.forward_classmap_device(x, dtype)
            accc += np.array([t for _, t in m.layer_times(dev, dtype)])
        m.set_profiling(False, dev, dtype)
        acc /= 3
        accc /= 3
        med_p, _ = timeit(lambda: m.forward_device(x, dtype), reps=5, warm=1)
        med_c, _ = timeit(lambda: m.forward_classmap_device(x, dtype), reps=5, warm=1)
        med_l, _ = timeit(lambda: m.forward_landmarks_device(x, dtype), reps=5, warm=1)
        med_t, _ = timeit(lambda: m.forward_landmarks_device(x, dtype, n_points=4), reps=5, warm=1)
        names = [L["name"] for L in m.graph.layers]
        top = sorted(zip(names, acc), key=lambda t: -t[1])[:12]
        print(json.dumps({"kernel": label + " forward", "dtype": dtype, "batch": B, "ms_layers_sum": float(acc.sum()),
                          "ms_forward_probs": med_p, "ms_forward_classmap": med_c, "images_per_s_probs": B / med_p * 1e3,
                          "images_per_s_classmap": B / med_c * 1e3, "ms_forward_soft_centroid": med_l, "images_per_s_soft_centroid": B / med_l * 1e3,
                          "ms_forward_top4_centroid": med_t, "images_per_s_top4_centroid": B / med_t * 1e3,
                          "last_deconv_ms": {"probs": round(float(acc[-2]), 4), "classmap": round(float(accc[-2]), 4)},
                          "layer_ms": {n: round(float(t), 4) for n, t in (zip(names, acc) if len(names) <= 20 else top)}}), flush=True)
        del x
        m._release()
        torch.cuda.empty_cache()

    if "fcn" in which:
        from keypoints_detector.networks.fcn import fcn_8
        m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
        for dtype, B in (("bfloat16", 32), ("bfloat16", 256), ("bfloat16", 1024), ("float32", 4)):   # 1024 = config C3
            fcn_bench("fcn_8/vanilla@224", m, dtype, B)

    if "encoders" in which:
        from keypoints_detector.networks import fcn
        for name, build in (("fcn_8_mobilenet", fcn.fcn_8_mobilenet), ("fcn_8_resnet50", fcn.fcn_8_resnet50), ("fcn_8_vgg", fcn.fcn_8_vgg)):
            m = build(68, 224, 224).init_weights(0)
            for B in (32, 128):
                fcn_bench(name + "@224", m, "bfloat16", B)


if __name__ == "__main__":
    main()
