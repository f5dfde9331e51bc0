#!/usr/bin/env python
"""bench.py — aligned faces/sec of the hot path (crop/resize -> landmark CNN -> decode -> Umeyama + warp).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--dtype bf16|fp32] [--impl reference]

A step = one pass of the hot path over one batch of B synthetic faces (BASELINE.json configs[1]: batch-256,
vanilla trunk @128 + FC-136 head) taken from 1080p frames, 64 faces per frame; aligned 112x112 crops out.
Under torchrun every rank owns one GPU and its own B faces (weak scaling, no data-path collective).
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "face-landmark-detector_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "aligned faces/sec (CNN fwd+decode+warp)"
FRAME_H, FRAME_W, FACES_PER_FRAME = 1080, 1920, 64


# ----------------------------------------------------------------------------------------- multi-rank plumbing
def rank_shard(n, rank, world):
    per = -(-n // world)
    return min(rank * per, n), min((rank + 1) * per, n)


def _reduce(val, device, op):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(val)
    t = torch.tensor([float(val)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(val, device):
    import torch.distributed as dist
    return _reduce(val, device, dist.ReduceOp.MAX)


def sum_over_ranks(val, device):
    import torch.distributed as dist
    return _reduce(val, device, dist.ReduceOp.SUM)


# ----------------------------------------------------------------------------------------- synthetic workload
def make_set(batch, seed, device=None):
    """One batch worth of inputs: frames uint8 [F,1080,1920,3], boxes int32 [B,4], face2frame int32 [B]."""
    from keypoints_detector.data import synthetic
    n_frames = -(-batch // FACES_PER_FRAME)
    g = torch.Generator(device="cpu").manual_seed(seed)
    noise = torch.randint(0, 256, (n_frames, FRAME_H, FRAME_W, 3), dtype=torch.uint8, generator=g)
    yy = torch.arange(FRAME_H, dtype=torch.float32)[:, None]
    xx = torch.arange(FRAME_W, dtype=torch.float32)[None, :]
    smooth = (127.5 + 127.5 * torch.sin(xx * 0.013 + yy * 0.021 + seed)).to(torch.uint8)
    frames = ((noise.to(torch.int16) + smooth[None, :, :, None].to(torch.int16)) // 2).to(torch.uint8)
    boxes = torch.from_numpy(synthetic.make_boxes(batch, FRAME_H, FRAME_W, seed=seed, min_side=96, max_side=400))
    f2f = (torch.arange(batch, dtype=torch.int32) // FACES_PER_FRAME).to(torch.int32)
    if device is not None:
        return frames.to(device), boxes.to(device), f2f.to(device)
    return frames, boxes, f2f


def valid_tap_macs(h, w, cin, cout, k=3):
    """MACs of a kxk pad-(k//2) stride-1 conv on an h x w map counting only in-image taps."""
    def taps(n):
        return sum(min(n, i + k // 2 + 1) - max(0, i - k // 2) for i in range(n))
    return taps(h) * taps(w) * cin * cout


TRUNK = [(128, 3, 64), (64, 64, 128), (32, 128, 256), (16, 256, 256), (8, 256, 256)]  # (map size, Cin, Cout)


def ncu_traffic_bytes(kernels=("conv_halo_kernel", "conv_tma_kernel")):
    """dram__bytes_read + dram__bytes_write of the dominant kernel family for ONE step (= one launch set), from the
    newest committed `ncu --set full` summary under profiles/ (tools/summarize_profiles.py); None if there is none."""
    import glob, re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_conv_full.txt")))
    if not files:
        return None, None
    total, cur = 0.0, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(files[-1]):
        if line.startswith("=="):
            cur = line
        m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m and cur and any(k in cur for k in kernels):
            total += float(m.group(2)) * unit.get(m.group(3), 1.0)
    return (total or None), os.path.basename(files[-1])


def flops_per_face():
    macs = [valid_tap_macs(s, s, ci, co) for s, ci, co in TRUNK]
    return [2.0 * m for m in macs], 2.0 * 4096 * 136


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The region is a few milliseconds long, so the sampler
    polls NVML from a thread (sub-millisecond per sample); `nvidia-smi -lms` (one sample per ~100 ms) is only the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index, period=0.0005):
        self.period, self.power = period, []
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.h, self.samples, self.run, self.thread = None, None, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                pr = torch.cuda.get_device_properties(index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.samples.append((mhz, mask))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _poll(self):
        while self.run:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            try:
                self._sample()                                # the GPU has just finished the last timed kernel
            except Exception:
                pass
            self.run = False
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            if not self.samples:
                return None
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": float(np.median([c for c, _ in self.samples])), "sm_max_mhz": self.max_mhz,
                    "sm_mhz_min": float(min(c for c, _ in self.samples)), "power_w_max": max(self.power) if self.power else None,
                    "power_w_median": float(np.median(self.power)) if self.power else None,
                    "reasons": sorted(n for bit, n in self.REASONS if mask & bit), "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------- CPU reference arm
def cpu_pipeline(frames, boxes, f2f, weights):
    """The reference's CPU path for the same step (oracle port; third-party cv2 / torch-CPU do the arithmetic the
    reference delegates to OpenCV / TensorFlow): prediction.py:76-94 per face + build-defined alignment."""
    import cv2
    from oracle import align as o_align, cnn as o_cnn, decode as o_decode, preprocess as o_pre
    crops, fbs = [], []
    for i in range(len(boxes)):
        fb = o_pre.square_box(boxes[i])
        fr = frames[f2f[i]]
        face = fr[max(fb[1], 0):fb[3], max(fb[0], 0):fb[2]]
        face = cv2.cvtColor(cv2.resize(face, (128, 128)), cv2.COLOR_BGR2RGB)
        crops.append(face); fbs.append(fb)
    crops = np.stack(crops)
    outs = [o_cnn.regression_forward(crops[s:s + 32], weights, torch.float32) for s in range(0, len(crops), 32)]  # Keras predict batch 32
    out = np.concatenate(outs)
    aligned = []
    for i in range(len(boxes)):
        marks, _ = o_decode.regression_decode(out[i], fbs[i])
        M = o_align.umeyama(o_align.five_points(marks), o_align.TEMPLATE_112)
        aligned.append(cv2.warpAffine(frames[f2f[i]], M, (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT))
    return np.stack(aligned)


def time_cpu(sample_faces, steps, warmup, seed=0):
    import cv2
    from keypoints_detector.networks.regression import landmark_regressor
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    model = landmark_regressor().init_weights(seed)
    frames, boxes, f2f = make_set(sample_faces, seed)
    frames, boxes, f2f = frames.numpy(), boxes.numpy(), f2f.numpy()
    for _ in range(warmup):
        cpu_pipeline(frames, boxes, f2f, model.weights)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pipeline(frames, boxes, f2f, model.weights)
    dt = time.perf_counter() - t0
    return sample_faces * steps / dt, dt / steps * 1e3, cores


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    sample = 64
    fps, ms, cores = time_cpu(sample, max(1, min(args.steps, 5)), min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "faces/s", "n_gpus": args.gpus, "steps": min(args.steps, 5),
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: batch-256 crops, vanilla trunk@128 + FC-136 head, + decode + 5-point align warp "
                                   "to 112x112 from 1080p frames", "sample_faces_per_step": sample},
            "cpu_baseline": {"value": fps, "unit": "faces/s", "cores": cores, "kind": "port",
                             "sample": "%d faces/step: cv2.resize+cvtColor, torch-CPU fp32 restatement of the Keras graph "
                                       "(TensorFlow not installable), numpy decode, fp64 Umeyama, cv2.warpAffine" % sample},
            "e2e": {"value": fps, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host memory is allocated, so the
    pinned frame / result buffers are first-touched on the GPU's NUMA node (the e2e leg is host-link-bound).
    Returns the number of CPUs bound to, or None when NVML gives no usable mask."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------- GPU arm
def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _cuda_time(fn, reps, dev, inner=1):
    """median / min ms of fn() (enqueue only) over `reps` CUDA-event brackets on the current stream"""
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts)), float(np.min(ts))


def sub_c4(dev, peaks):
    """config C4: 4096 faces from 64 synthetic 1080p frames, 5-point fit + warp to 112x112 (align-only), HBM roofline."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    F, B = 64, 4096
    g = torch.Generator(device="cpu").manual_seed(7)
    frames = torch.randint(0, 256, (F, FRAME_H, FRAME_W, 3), dtype=torch.uint8, generator=g).to(dev)   # 398 MB > L2
    pts, Ms = synthetic.make_similarity_landmarks(B, FRAME_H, FRAME_W, prediction.TEMPLATE_112, seed=4)
    f2f = torch.from_numpy((np.arange(B) // 64).astype(np.int32)).to(dev)
    marks = torch.from_numpy(pts).to(dev)
    out = torch.empty((B, 112, 112, 3), dtype=torch.uint8, device=dev)
    M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
    fn = lambda: prediction.align_device(frames, f2f, marks, None, (112, 112), five_point=False, out=out, out_matrix=M)
    for _ in range(3):
        fn()
    med, mn = _cuda_time(fn, 20, dev, inner=4)
    s2 = np.array([np.linalg.det(m[:, :2]) for m in Ms])                     # scale^2 per face
    alg = B * 37632 + float((112 * 112 * 3 / s2).sum()) + B * (5 * 2 * 4 + 48)  # crop write + unique source footprint + marks + M
    hbm = peaks.get("hbm_gbs", 6551.7)
    return {"workload": "configs[3]: 4096 faces, 64 x 1080p frames -> 112x112x3, fit + warp", "kernel": "align_tile_kernel",
            "ms_median": med, "ms_min": mn, "faces_per_s": B / med * 1e3, "algorithmic_bytes": alg, "bytes_per_face": alg / B,
            "achieved_GBs": alg / med / 1e6, "peak_GBs": hbm, "frac_of_hbm": alg / med / 1e6 / hbm, "bound": "hbm",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}


def sub_c3(dev, peaks, batch=1024):
    """config C3: fcn_8 over the vanilla encoder @224x224, 68 classes, batch 1024, soft-argmax fused into the last transposed
    conv's epilogue (bf16 tensor cores): images/s, TFLOP/s on the valid-tap FLOP count of SURVEY §8d, share of up8."""
    from keypoints_detector.networks.fcn import fcn_8
    m = fcn_8(68, input_height=224, input_width=224).init_weights(0)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn((64, 224, 224, 3), generator=g) * 50).to(dev).repeat(batch // 64, 1, 1, 1).contiguous()
    x += torch.arange(batch, device=dev, dtype=torch.float32).view(-1, 1, 1, 1) * 0.01        # no two images alike
    fn = lambda: m.forward_landmarks_device(x, "bfloat16", n_points=0)
    for _ in range(2):
        fn()
    med, mn = _cuda_time(fn, 5, dev)
    m.set_profiling(True, dev, "bfloat16")
    fn()
    lt = dict(m.layer_times(dev, "bfloat16"))
    m.set_profiling(False, dev, "bfloat16")
    gflop = 11.37                                                              # per image, valid taps
    tf = gflop * 1e9 * batch / (med * 1e-3) / 1e12
    burst, sust = peaks.get("bf16_tflops", 1642.1), peaks.get("bf16_tflops_sustained", 1351.1)
    top = sorted(lt.items(), key=lambda kv: -kv[1])[:6]
    rec = {"workload": "configs[2]: fcn_8/vanilla @224x224x3, 68 classes, batch %d, fused soft-argmax decode" % batch, "dtype": "bf16",
           "ms_per_batch": med, "ms_min": mn, "images_per_s": batch / med * 1e3, "gflop_per_image_valid_taps": gflop,
           "achieved_TFLOPs": tf, "frac_burst": tf / burst, "frac_sustained": tf / sust,
           "up8_ms": lt.get("up8"), "up8_share": (lt.get("up8", 0.0) / sum(lt.values())) if lt else None,
           "layer_ms_top": {k: round(v, 4) for k, v in top}}
    # the reference's own FCN output is the argmax class map (prediction.py:209): the same net as class maps, in plain bf16 and in
    # the drop-in's default fp32-accurate mode (bf16x3: split-operand tensor-core convs, first layer and last transposed conv)
    xs = x[:256]
    cm = {}
    for mode in ("bfloat16", "bf16x3"):
        f2 = lambda: m.forward_classmap_device(xs, mode)
        for _ in range(2):
            f2()
        med2, _ = _cuda_time(f2, 5, dev)
        cm[mode] = {"ms_per_256": med2, "images_per_s": 256 / med2 * 1e3}
    rec["classmap"] = cm
    f3 = lambda: m.forward_landmarks_device(xs, "bf16x3", n_points=0)
    for _ in range(2):
        f3()
    med3, _ = _cuda_time(f3, 5, dev)
    rec["soft_argmax_bf16x3"] = {"ms_per_256": med3, "images_per_s": 256 / med3 * 1e3}
    del x, xs
    m._release()
    torch.cuda.empty_cache()
    return rec


def sub_mode(dev, model, dev_set, dtype, steps, batch):
    """config C2 in another compute mode (fp32 CUDA cores / fp32-accurate bf16x3 tensor cores), device-resident, one lane."""
    from keypoints_detector import prediction
    pipe = prediction.LandmarkPipeline(model, dtype=dtype, device=dev)
    cap = pipe.capture(*dev_set, lane=7)
    for _ in range(2):
        cap.replay()
    med, mn = _cuda_time(cap.replay, steps, dev)
    cap.close()
    return {"workload": "configs[1], %s mode, device-resident, 1 lane" % dtype, "dtype": dtype, "ms_per_step": med,
            "faces_per_s": batch / med * 1e3, "TFLOPs_valid_taps": sum(flops_per_face()[0]) * batch / (med * 1e-3) / 1e12}


def sub_strong(model, dtype, n_gpus, total_faces=65536, frames_n=128):
    """config C5, strong scaling through the product's multi-GPU API: ONE batch of 65 536 faces (128 distinct 1080p frames) in
    pinned host memory, sharded by frame over the GPUs by MultiGpuPipeline (thread per GPU, no collective), results gathered
    into one pinned host buffer.  Wall-clock per call (host threads included)."""
    from keypoints_detector import prediction
    from keypoints_detector.data import synthetic
    g = torch.Generator(device="cpu").manual_seed(77)
    frames = torch.randint(0, 256, (frames_n, FRAME_H, FRAME_W, 3), dtype=torch.uint8, generator=g).pin_memory()
    boxes = torch.from_numpy(synthetic.make_boxes(total_faces, FRAME_H, FRAME_W, seed=78, min_side=96, max_side=400)).pin_memory()
    f2f = (torch.arange(total_faces, dtype=torch.int32) // (total_faces // frames_n)).to(torch.int32).pin_memory()
    mg = prediction.MultiGpuPipeline(model, devices=list(range(n_gpus)), dtype=dtype)
    mg.run(frames, boxes, f2f)                                                # warm-up: plans, buffers, pinned gather buffers
    walls, devms = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        r = mg.run(frames, boxes, f2f)
        walls.append(time.perf_counter() - t0)
        devms.append(max(mg.last_device_ms))
    import hashlib
    digest = hashlib.sha256(r["marks"].tobytes() + r["aligned"].tobytes()).hexdigest()
    mg.close()
    w = float(np.median(walls))
    return {"workload": "configs[4]: one batch of %d faces from %d 1080p frames in pinned host memory, sharded by frame over %d GPU(s), "
                        "host-side gather" % (total_faces, frames_n, n_gpus), "scaling": "strong", "n_gpus": n_gpus,
            "faces_per_s": total_faces / w, "wall_ms": w * 1e3, "max_device_ms": float(np.median(devms)),
            "h2d_bytes": int(frames.numel() + boxes.numel() * 4 + f2f.numel() * 4), "d2h_bytes": int(total_faces * (112 * 112 * 3 + 68 * 2 * 4 + 48 + 16)),
            "result_sha256": digest[:16]}


def run_gpu(args, rank, world, local_rank):
    from keypoints_detector import _native, prediction
    from keypoints_detector.networks.regression import landmark_regressor
    import __graft_entry__ as entry
    import torch.distributed as dist
    if local_rank == 0:
        entry.build()
    cpu_group = None
    if world > 1:
        dist.barrier()
        cpu_group = dist.new_group(backend="gloo")     # CPU-side waits: an NCCL barrier would spin a kernel on every waiting GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    bound_cpus = bind_to_gpu_numa_node(local_rank)
    B = args.batch
    dtype = {"bf16": "bfloat16", "fp32": "float32", "bf16x3": "bf16x3"}[args.dtype]
    peaks = _peaks()
    model = landmark_regressor().init_weights(seed=0)
    pipe = prediction.LandmarkPipeline(model, dtype=dtype, device=dev)

    n_sets = args.sets
    host_sets = [make_set(B, 100 + 10 * rank + s) for s in range(n_sets)]
    dev_sets = [tuple(t.to(dev) for t in hs) for hs in host_sets]
    set_bytes = sum(t.numel() * t.element_size() for t in host_sets[0])

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, streams):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for k in range(steps):
            fn(k)
        for st in streams:
            cur.wait_stream(st)
        e1.record(cur)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- device-resident throughput.  Consecutive steps are independent batches: they alternate between `lanes` streams
    # (own activation workspace each), so the crop/resize, FC, decode and warp kernels of one batch overlap the tensor-core
    # convs of the other.  Every step still runs completely inside the timed region.
    n_lanes = max(1, args.lanes)
    lane_streams = [torch.cuda.Stream(dev) for _ in range(n_lanes)]
    model_profiling = [False]                                 # the per-layer event pass runs eagerly

    # Each (input set, lane) pair is recorded once into a CUDA graph (LandmarkPipeline.capture): a step is then ONE launch
    # from the host instead of ~12 ctypes calls, which otherwise cost as much host time as the step takes on the GPU.
    use_graph = not args.no_graph
    if n_sets % n_lanes != 0:
        n_sets -= n_sets % n_lanes                            # set s always runs on lane s % n_lanes
    launches_per_step = None
    graphs = []
    if use_graph:
        try:
            for sidx in range(n_sets):
                l_before = _native.launch_count()
                graphs.append(pipe.capture(*dev_sets[sidx], lane=sidx % n_lanes))
                launches_per_step = (_native.launch_count() - l_before) // 2  # capture() = one warm-up run + the recorded run
        except Exception as e:                                # keep measuring: every kernel is then enqueued from Python
            print("bench: CUDA graph capture failed (%s); running eagerly" % e, file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize(dev)

    def step_resident(k):
        ln = k % n_lanes
        with torch.cuda.stream(lane_streams[ln]):
            if use_graph and not model_profiling[0]:
                graphs[k % n_sets].replay()
            else:
                pipe.run_device(*dev_sets[k % n_sets], lane=ln)

    for k in range(args.warmup):
        step_resident(k)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _native.launch_count()
    ms = timed(step_resident, args.steps, lane_streams)
    launches = (_native.launch_count() - l0) if not use_graph else launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    total_faces = sum_over_ranks(B, dev)
    value = total_faces * args.steps / (ms * 1e-3)

    # ---- second timed pass with per-layer CUDA events: live duration of the dominant kernel (roofline)
    model.set_profiling(True, dev, dtype)
    model_profiling[0] = True
    per_layer = np.zeros(len(model.graph.layers))
    barrier()
    n_prof = min(args.steps, 50)
    for k in range(n_prof):
        step_resident(k)
        per_layer += np.array([t for _, t in model.layer_times(dev, dtype)])
    model.set_profiling(False, dev, dtype)
    model_profiling[0] = False
    per_layer /= n_prof
    conv_flops, fc_flops = flops_per_face()
    burst, sust = peaks.get("bf16_tflops", 1642.1), peaks.get("bf16_tflops_sustained", 1351.1)
    # which regime the timed region ran in: seconds at full power end up at the sustained clocks, a short burst does not
    regime = "sustained" if ms >= 2000.0 else "burst"
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"

    def tensor_roof(kernel, fl, t_s, extra=None):
        ach = fl / t_s / 1e12
        r = {"kernel": kernel, "bound": "tensor", "achieved": ach, "unit": "TFLOP/s",
             "peak": burst if regime == "burst" else sust, "peak_regime": regime,
             "peak_source": "%s bf16_tflops%s (timed region %.0f ms: %s regime)" % (src, "" if regime == "burst" else "_sustained", ms, regime),
             "frac": ach / (burst if regime == "burst" else sust), "frac_burst": ach / burst, "frac_sustained": ach / sust,
             "frac_nominal_2250": ach / 2250.0, "flops_per_launch_set": fl, "ms_per_launch_set": t_s * 1e3}
        if extra:
            r.update(extra)
        return r

    if dtype in ("bfloat16", "bf16x3"):
        traffic, traffic_src = ncu_traffic_bytes()
        roof = tensor_roof("tcgen05 conv trunk: conv_halo_kernel x3 (conv2..conv4) + conv_tma_kernel (conv5), 4 launches/step",
                           sum(conv_flops[1:5]) * B, float(per_layer[1:5].sum()) * 1e-3,
                           {"traffic": traffic, "traffic_source": traffic_src, "traffic_note": "DRAM bytes per launch set at batch 256"})
        roof_cnn = tensor_roof("whole CNN: conv1 (conv_s2d_kernel) + conv2..conv5 + FC (split-K conv_tma_kernel + dense_reduce_kernel), 7 launches/step",
                               (sum(conv_flops) + fc_flops) * B, float(per_layer.sum()) * 1e-3)
    else:
        peak = 75.0  # fp32 FMA nominal: 148 SMs x 128 lanes x 2 x ~1.97 GHz
        t_dom = float(per_layer[0:5].sum()) * 1e-3
        fl = sum(conv_flops) * B
        roof = {"kernel": "conv_simt_kernel (conv1..conv5, fp32 CUDA cores)", "bound": "fp32-fma", "achieved": fl / t_dom / 1e12, "peak": peak,
                "peak_source": "nominal fp32 FMA (no measured fp32 peak in MEASURED_PEAKS.json)", "unit": "TFLOP/s", "traffic": None,
                "flops_per_launch_set": fl, "ms_per_launch_set": t_dom * 1e3, "frac": fl / t_dom / 1e12 / peak}
        roof_cnn = None
    roof["layer_ms"] = {L["name"]: round(float(t), 5) for L, t in zip(model.graph.layers, per_layer)}

    # ---- end to end through the package's host-buffer API (prediction.HostStream): H2D + compute + D2H every step
    pin_sets = [tuple(t.pin_memory() for t in hs) for hs in host_sets]
    n_slots = max(2, n_lanes)   # in-flight steps: H2D of step k+2 never waits for the compute of step k when there are 3
    hs = prediction.HostStream(pipe, B, host_sets[0][0].shape[0], (FRAME_H, FRAME_W), n_slots=n_slots, use_graph=use_graph)

    def step_e2e(k):
        hs.submit(*pin_sets[k % n_sets])

    timed(step_e2e, max(args.warmup, 2), hs.streams())
    host_t = [0.0]

    def step_e2e_timed(k):
        t0 = time.perf_counter()
        step_e2e(k)
        host_t[0] += time.perf_counter() - t0

    ms_e2e = timed(step_e2e_timed, args.steps, hs.streams())
    e2e_val = total_faces * args.steps / (ms_e2e * 1e-3)
    last = hs.result(hs.n_submitted - 1)
    e2e_check = bool(np.isfinite(last["marks"]).all() and last["aligned"].any())
    d2h = hs.d2h_bytes

    # ---- what the host link of this box allows: the same pinned copies alone, each direction and both at once
    cur = torch.cuda.current_stream(dev)
    s_a, s_b = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    n_link = max(args.steps, 10)

    def h2d(k):
        with torch.cuda.stream(s_a):
            for d, h in zip(hs.d_in[k % n_slots], pin_sets[k % n_sets]):
                d.copy_(h, non_blocking=True)

    res0 = hs.caps[0].results if hs.caps else pipe.run_device(*hs.d_in[0], lane=0)

    def d2h_fn(k):
        with torch.cuda.stream(s_b):
            hs.h_out[k % n_slots]["marks"].copy_(res0["marks"], non_blocking=True)
            hs.h_out[k % n_slots]["aligned"].copy_(res0["aligned"], non_blocking=True)

    ms_h2d = timed(h2d, n_link, [s_a]) / n_link
    ms_d2h = timed(d2h_fn, n_link, [s_b]) / n_link
    ms_dup = timed(lambda k: (h2d(k), d2h_fn(k)), n_link, [s_a, s_b]) / n_link
    link = {"h2d_GBps": set_bytes / ms_h2d / 1e6, "d2h_GBps": d2h / ms_d2h / 1e6,
            "duplex_ms_per_step": ms_dup, "duplex_h2d_GBps": set_bytes / ms_dup / 1e6, "duplex_d2h_GBps": d2h / ms_dup / 1e6,
            "h2d_GBps_all_ranks": sum_over_ranks(set_bytes / ms_h2d / 1e6, dev), "d2h_GBps_all_ranks": sum_over_ranks(d2h / ms_d2h / 1e6, dev),
            "bytes_per_face_h2d": set_bytes / B, "bytes_per_face_d2h": d2h / B,
            "link_bound_faces_per_s": total_faces / (ms_dup * 1e-3),
            "note": "copy-only passes over the same pinned buffers (max over ranks); link_bound_faces_per_s = faces per step / duplex copy time: "
                    "the e2e leg cannot exceed it on this box whatever the kernels do"}
    try:    # write-combined pinned source for the H2D direction (cudaHostAllocWriteCombined = 4)
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        ptr = ctypes.c_void_p()
        nbytes = host_sets[0][0].numel()
        if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(4)) == 0:
            wc = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr.value)))
            wc.copy_(host_sets[0][0].reshape(-1))
            dst = hs.d_in[0][0].view(-1)

            def h2d_wc(k):
                with torch.cuda.stream(s_a):
                    dst.copy_(wc, non_blocking=True)
            link["h2d_write_combined_GBps"] = nbytes / (timed(h2d_wc, n_link, [s_a]) / n_link) / 1e6
            torch.cuda.synchronize(dev)
            rt.cudaFreeHost(ptr)
    except Exception as e:
        link["h2d_write_combined_GBps"] = None

    # ---- bit identity across GPUs: every rank runs the SAME seeded batch; the digests of landmarks + aligned crops must agree
    import hashlib
    fx = tuple(t.to(dev) for t in make_set(B, 4242))
    rfix = pipe.run_device(*fx, lane=0)
    torch.cuda.synchronize(dev)
    digest = hashlib.sha256(rfix["marks"].cpu().numpy().tobytes() + rfix["aligned"].cpu().numpy().tobytes() + rfix["M"].cpu().numpy().tobytes()).hexdigest()
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, digest, group=cpu_group)
    else:    # one GPU: a second lane on another stream must reproduce it
        with torch.cuda.stream(lane_streams[-1]):
            r2 = pipe.run_device(*fx, lane=n_lanes - 1 if n_lanes > 1 else 5)
        torch.cuda.synchronize(dev)
        digests = [digest, hashlib.sha256(r2["marks"].cpu().numpy().tobytes() + r2["aligned"].cpu().numpy().tobytes() + r2["M"].cpu().numpy().tobytes()).hexdigest()]
    bit_identical = len(set(digests)) == 1

    # ---- sub-records: other configs / modes the driver should see in the same line (rank 0; the other ranks wait on the CPU)
    subs = {}
    if rank == 0 and not args.no_sub:
        wanted = [w for w in args.sub.split(",") if w]

        def guarded(name, fn):
            if name not in wanted:
                return
            t0 = time.perf_counter()
            try:
                subs[name] = fn()
            except Exception as e:   # a failing side measurement must not lose the contract line
                subs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.synchronize(dev)
            subs[name]["bench_seconds"] = round(time.perf_counter() - t0, 2)

        def sustained():
            n = int(2300.0 / (ms / args.steps)) + 1
            smp = ClockSampler(local_rank, period=0.01)
            smp.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            cur.wait_stream(lane_streams[0])
            e0.record(cur)
            for st in lane_streams:
                st.wait_event(e0)
            for k in range(n):
                step_resident(k)
            for st in lane_streams:
                cur.wait_stream(st)
            e1.record(cur)
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1)
            ck = smp.stop()
            v = B * n / (t * 1e-3)
            cnn_fl = (sum(conv_flops) + fc_flops) * B * n / (t * 1e-3) / 1e12
            return {"workload": "configs[1] loop for %.1f s (%d steps), this GPU" % (t / 1e3, n), "faces_per_s": v, "ms_per_step": t / n,
                    "vs_short_run": v / (value / world), "whole_step_TFLOPs": cnn_fl, "frac_sustained_peak": cnn_fl / sust, "clocks": ck}

        guarded("sustained", sustained)
        guarded("c4", lambda: sub_c4(dev, peaks))
        guarded("fp32", lambda: sub_mode(dev, model, dev_sets[0], "float32", 5, B))
        if dtype != "bf16x3":
            guarded("bf16x3", lambda: sub_mode(dev, model, dev_sets[0], "bf16x3", 20, B))
        guarded("c3", lambda: sub_c3(dev, peaks, 1024))
        guarded("strong", lambda: sub_strong(model, dtype, world))
    if world > 1:
        dist.barrier(group=cpu_group)

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            os.sched_setaffinity(0, all_cpus)             # the CPU leg gets every host core again
            fps, cms, cores = time_cpu(64, 3, 1)
            cpu = {"value": fps, "unit": "faces/s", "cores": cores, "kind": "port",
                   "sample": "3 steps x 64 faces of the same workload: cv2.resize+cvtColor, torch-CPU fp32 restatement of the Keras "
                             "graph (TensorFlow not installable), numpy decode, fp64 Umeyama, cv2.warpAffine"}
        line = {"metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bfloat16": "bf16", "float32": "f32", "bf16x3": "bf16x3 (fp32-accurate split)"}[dtype], "data": "synthetic",
                "config": {"workload": "configs[1]: batch-%d crops/GPU from %d 1080p frames, vanilla trunk@128 + FC-136 head, "
                                       "+ decode + 5-point align warp to 112x112" % (B, -(-B // FACES_PER_FRAME)),
                           "batch_per_gpu": B, "global_batch": int(total_faces), "parallelism": "faces sharded, no collective",
                           "lanes": n_lanes, "cuda_graph": bool(use_graph),
                           "l2": "inputs rotate over %d distinct sets (%.0f MB) > 126 MB L2; activations workspace rewritten every step"
                                 % (n_sets, n_sets * set_bytes / 1e6)},
                "e2e": {"value": e2e_val, "unit": "faces/s", "h2d_bytes_per_step": int(set_bytes), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": ms_e2e / args.steps, "host_enqueue_ms_per_step": host_t[0] * 1e3 / args.steps, "api": "keypoints_detector.prediction.HostStream.submit / result (pinned host "
                        "buffers in and out, %d slots)" % n_slots, "results_checked": e2e_check, "cpus_bound": bound_cpus, "link": link,
                        "frac_of_link_bound": e2e_val / link["link_bound_faces_per_s"]},
                "gpu_launches": int(launches), "roofline": roof, "roofline_cnn": roof_cnn, "cross_gpu_bit_identical": bit_identical,
                "cross_gpu_digests": sorted(set(d[:12] for d in digests)), "sub": subs, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # a step is ~0.45 ms: 200 keeps pipeline fill / drain under 1 % of the region
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sets", type=int, default=6)
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--lanes", type=int, default=3, help="independent batches in flight (streams with their own workspace)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (other configs / modes)")
    ap.add_argument("--sub", default="sustained,c4,fp32,bf16x3,c3,strong", help="comma-separated sub-records to measure on rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        return run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
