"""keypoints_detector.prediction — drop-in for the reference module of the same name, with the arithmetic
on a B200 (sm_100a CUDA kernels behind libfld_sm100.so).

Unchanged entry points (reference prediction.py): detect_marks :16-96, video_predict :99-113,
model_from_checkpoint_path :116-133, keypts_predict :158-196, _prediction :199-222.
Additions: detect_marks_batch, align_faces, LandmarkPipeline (batched path: run_device for device-resident tensors, run_host
for host arrays), HostStream (overlapped copy / compute / copy pipeline over pinned host buffers — what bench.py's e2e leg
calls), MultiGpuPipeline (one batch sharded by frame over the GPUs of a box with a host-side gather), shard_faces /
shard_by_frames.

Where the reference crashes as written behaviour is defined and documented inline.
"""
import json
import os
import typing

import numpy as np
import six
import torch

from . import _native as N
from .data.config import IMAGE_ORDERING
from .data.generator import get_image_array
from .training import find_latest_checkpoint
from .utils.plots import class_colors, draw_marks, visualize_keypoints

# widely used 112x112 five-point alignment template
TEMPLATE_112 = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                         [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float64)


def _device(device=None):
    if not torch.cuda.is_available():
        N.handle()  # raises: CUDA-only, no CPU fallback
    return torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index)


# ----------------------------------------------------------------------------------------------- device ops
def preprocess_faces_device(frames, boxes, face2frame, size=128, swap_rb=True, out=None, out_boxes=None, staging=None):
    """frames uint8 CUDA [F,H,W,3] BGR; boxes int32 CUDA [B,4]; face2frame int32 CUDA [B].
    -> (crops uint8 [B,size,size,3] RGB, faceboxes int32 [B,4]); reference prediction.py:76-83.
    `staging` (Model.input_staging): also write the crops into the network's first-layer operand staging buffer."""
    lib = N.load_library()
    F, H, W, C = frames.shape
    assert C == 3 and frames.dtype == torch.uint8
    B = boxes.shape[0]
    out = torch.empty((B, size, size, 3), dtype=torch.uint8, device=frames.device) if out is None else out
    fb = torch.empty((B, 4), dtype=torch.int32, device=frames.device) if out_boxes is None else out_boxes
    with torch.cuda.device(frames.device):
        if staging is not None:
            N.check(lib.fld_preprocess_faces_staged(N.handle(frames.device), N.ptr(frames), F, H, W, N.ptr(boxes), N.ptr(face2frame), B,
                                                    size, int(swap_rb), N.ptr(out), N.ptr(fb), staging, N.stream_ptr(frames.device)))
        else:
            N.check(lib.fld_preprocess_faces(N.handle(frames.device), N.ptr(frames), F, H, W, N.ptr(boxes), N.ptr(face2frame), B, size,
                                             int(swap_rb), N.ptr(out), N.ptr(fb), N.stream_ptr(frames.device)))
    return out, fb


def decode_regress_device(out136, faceboxes, want_uint=True, out=None, out_uint=None):
    """reference prediction.py:88-94 on device: -> (marks float32 [B,68,2], marks uint64-as-int64 [B,68,2] or None)."""
    lib = N.load_library()
    B = out136.shape[0]
    marks = torch.empty((B, 68, 2), dtype=torch.float32, device=out136.device) if out is None else out
    marks_u = None
    if want_uint:
        marks_u = torch.empty((B, 68, 2), dtype=torch.int64, device=out136.device) if out_uint is None else out_uint
    with torch.cuda.device(out136.device):
        N.check(lib.fld_decode_regress(N.handle(out136.device), N.ptr(out136), out136.shape[1], N.ptr(faceboxes), B, N.ptr(marks),
                                       N.ptr(marks_u), N.stream_ptr(out136.device)))
    return marks, marks_u


_TEMPLATE_CACHE = {}


def _template_device(template, device):
    """fp64 template resident on `device`, uploaded once: a per-call pageable H2D copy would synchronise the stream and
    serialise the whole pipeline behind it."""
    tmpl = TEMPLATE_112 if template is None else np.ascontiguousarray(template, dtype=np.float64)
    key = (device.index, tmpl.shape, tmpl.tobytes())
    t = _TEMPLATE_CACHE.get(key)
    if t is None:
        if len(_TEMPLATE_CACHE) > 64:
            _TEMPLATE_CACHE.clear()
        t = torch.from_numpy(np.array(tmpl, dtype=np.float64)).to(device)
        _TEMPLATE_CACHE[key] = t
    return t


ALIGN_ORDER_MIN_FACES = 2048     # from here on the library warps the faces in order of decreasing box size (needs a scratch buffer)


def _align_scratch(lib, device, B, scratch):
    """uint8 scratch tensor for the ordered alignment calls (fld_align_scratch_bytes), or None for small batches."""
    if B < ALIGN_ORDER_MIN_FACES:
        return None
    need = int(lib.fld_align_scratch_bytes(N.handle(device), B))
    if scratch is None:
        return torch.empty((need,), dtype=torch.uint8, device=device)
    if scratch.numel() * scratch.element_size() < need or scratch.device != device:
        raise ValueError("align scratch needs %d bytes on %s" % (need, device))
    return scratch


def align_device(frames, face2frame, marks, template=None, out_size=(112, 112), five_point=True, return_matrix=True, out=None,
                 out_matrix=None, scratch=None):
    """Umeyama fit + cv2.warpAffine-exact warp on device.  marks float32 CUDA [B,N,2] in frame pixels.  Batches of
    ALIGN_ORDER_MIN_FACES faces or more go through fld_align_ordered (same results; `scratch`: optional caller-owned uint8
    tensor, else one is taken from torch's stream-ordered allocator)."""
    lib = N.load_library()
    F, H, W, C = frames.shape
    B, Np = marks.shape[0], marks.shape[1]
    t = _template_device(template, frames.device)
    oh, ow = out_size
    crops = torch.empty((B, oh, ow, C), dtype=torch.uint8, device=frames.device) if out is None else out
    M = None
    if return_matrix:
        M = torch.empty((B, 2, 3), dtype=torch.float64, device=frames.device) if out_matrix is None else out_matrix
    with torch.cuda.device(frames.device):
        sc = _align_scratch(lib, frames.device, B, scratch)
        if sc is None:
            N.check(lib.fld_align(N.handle(frames.device), N.ptr(frames), F, H, W, C, N.ptr(face2frame), N.ptr(marks), Np, N.ptr(t),
                                  t.shape[0], int(five_point), B, oh, ow, N.ptr(M), N.ptr(crops), N.stream_ptr(frames.device)))
        else:
            N.check(lib.fld_align_ordered(N.handle(frames.device), N.ptr(frames), F, H, W, C, N.ptr(face2frame), N.ptr(marks), Np,
                                          N.ptr(t), t.shape[0], int(five_point), B, oh, ow, N.ptr(M), N.ptr(crops), N.ptr(sc),
                                          sc.numel(), N.stream_ptr(frames.device)))
    return crops, M


def warp_affine_device(frames, face2frame, M, out_size=(112, 112), scratch=None):
    lib = N.load_library()
    F, H, W, C = frames.shape
    B = M.shape[0]
    oh, ow = out_size
    crops = torch.empty((B, oh, ow, C), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        sc = _align_scratch(lib, frames.device, B, scratch)
        if sc is None:
            N.check(lib.fld_warp_affine(N.handle(frames.device), N.ptr(frames), F, H, W, C, N.ptr(face2frame), N.ptr(M), B, oh, ow,
                                        N.ptr(crops), N.stream_ptr(frames.device)))
        else:
            N.check(lib.fld_warp_affine_ordered(N.handle(frames.device), N.ptr(frames), F, H, W, C, N.ptr(face2frame), N.ptr(M), B,
                                                oh, ow, N.ptr(crops), N.ptr(sc), sc.numel(), N.stream_ptr(frames.device)))
    return crops


def class_map_device(scores, oh, ow):
    """reference prediction.py:209 on device: scores float32 CUDA [B, oh*ow, L] -> int64 [B,oh,ow]."""
    lib = N.load_library()
    B, hw, L = scores.shape
    out = torch.empty((B, oh, ow), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        N.check(lib.fld_decode_classmap(N.handle(scores.device), N.ptr(scores), B, hw, L, N.ptr(out), N.stream_ptr(scores.device)))
    return out


# ----------------------------------------------------------------------------------------------- batched pipeline
class CapturedRun:
    """One LandmarkPipeline run recorded as a CUDA graph.  The graph's kernels have raw device pointers baked in (inputs, the
    lane's activation workspace and result buffers, the kernel plans' tile schedules), so this object keeps all of them alive
    and in place: it holds the tensors, pins the lane's workspace against re-allocation and retains the net's plans
    (fld_net_retain).  Replacing the model's weights invalidates it — replay() then raises instead of touching freed memory.
    Unpacks like the (graph, results) pair capture() used to return."""

    def __init__(self, graph, results, keep):
        self.graph, self.results, self._keep = graph, results, keep
        self._release, self._why_invalid = None, None

    def replay(self):
        if self._why_invalid is not None:
            raise N.FldError("captured run is no longer valid: " + self._why_invalid)
        self.graph.replay()
        return self.results

    def invalidate(self, why):
        self._why_invalid = why
        self._release = None

    def close(self):
        if self._release is not None:
            self._release()
            self._release = None
        self._why_invalid = self._why_invalid or "closed"

    def __del__(self):
        try:
            if self._release is not None:
                self._release()
        except Exception:
            pass

    def __iter__(self):
        return iter((self, self.results))


class LandmarkPipeline:
    """frames + detector boxes -> 68 landmarks -> aligned crops, entirely on one GPU:
    crop/resize (a1) -> regression CNN (a2) -> decode (a3) -> Umeyama + warp (a10)."""

    def __init__(self, model, dtype=None, out_size=(112, 112), template=None, device=None):
        self.model = model
        self.dtype = dtype or model.compute_dtype        # default: the model's mode ("bf16x3": fp32-accurate tensor cores)
        self.out_size = tuple(out_size)
        self.template = TEMPLATE_112 if template is None else np.asarray(template, dtype=np.float64)
        self.device = _device(device)
        self.input_size = model.input_height
        self._bufs = {}
        self._host_state = {}                 # run_host's cached staging / result buffers
        self.max_batch = 4096
        with torch.cuda.device(self.device):
            model.compiled(self.device, dtype)

    def _buffers(self, lane, B, C, dev):
        """Per-lane result buffers, allocated once per batch size: no allocator traffic (and no allocator-induced stream
        synchronisation) in the steady state.  They are overwritten by the lane's next run."""
        key = (lane, B, C)
        buf = self._bufs.get(key)
        if buf is None:
            S, (oh, ow) = self.input_size, self.out_size
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            buf = {"crops": e((B, S, S, 3), torch.uint8), "faceboxes": e((B, 4), torch.int32),
                   "net_out": e((B, self.model.n_classes), torch.float32), "marks": e((B, 68, 2), torch.float32),
                   "marks_uint": e((B, 68, 2), torch.int64), "aligned": e((B, oh, ow, C), torch.uint8),
                   "M": e((B, 2, 3), torch.float64)}
            self._bufs[key] = buf
        return buf

    def run_device(self, frames, boxes, face2frame, want_uint=False, lane=0):
        """Everything is enqueued on the current CUDA stream.  Independent batches may be enqueued on different streams
        concurrently when each uses its own `lane` (activation workspace + result buffers): the HBM/issue-bound kernels of
        one batch (crop/resize, FC, decode, warp) then fill the issue slots the tensor-core convs of the other leave idle.
        The returned tensors belong to the lane and are overwritten by its next run.  Batches larger than `max_batch`
        are processed in chunks (bounded activation workspace), results land in one buffer."""
        B = boxes.shape[0]
        b = self._buffers(lane, B, frames.shape[3], frames.device)
        for s0 in range(0, max(B, 1), self.max_batch):
            s1 = min(B, s0 + self.max_batch)
            # the crop / resize kernel also fills the first conv layer's operand staging when the net has one (same values as the
            # network's own widening pass over the crops, which is then skipped)
            stg = self.model.input_staging(s1 - s0, frames.device, self.dtype, lane) if s1 > s0 else None
            crops128, fb = preprocess_faces_device(frames, boxes[s0:s1], face2frame[s0:s1], self.input_size, True,
                                                   out=b["crops"][s0:s1], out_boxes=b["faceboxes"][s0:s1], staging=stg)
            out = self.model.forward_device(crops128, self.dtype, out=b["net_out"][s0:s1], lane=lane, staged=stg is not None)
            marks, marks_u = decode_regress_device(out, fb, want_uint, out=b["marks"][s0:s1], out_uint=b["marks_uint"][s0:s1])
            align_device(frames, face2frame[s0:s1], marks, self.template, self.out_size, True, True, out=b["aligned"][s0:s1],
                         out_matrix=b["M"][s0:s1])
        return {"marks": b["marks"], "marks_uint": b["marks_uint"] if want_uint else None, "aligned": b["aligned"], "M": b["M"],
                "faceboxes": b["faceboxes"], "crops": b["crops"]}

    def capture(self, frames, boxes, face2frame, want_uint=False, lane=0):
        """Record one run over FIXED input buffers into a CUDA graph (small batches are launch-bound: ~12 kernels plus the
        Python shim per run).  Returns a CapturedRun (unpacks as (graph, results)): refill `frames` / `boxes` / `face2frame`
        in place, call `.replay()`, read `.results` (the lane's buffers).  The capture keeps every buffer its kernels point
        at alive; a lane with a live capture cannot be used for a larger batch."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):            # warm-up outside the capture: plans, workspaces, buffers get created
                self.run_device(frames, boxes, face2frame, want_uint, lane)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = self.run_device(frames, boxes, face2frame, want_uint, lane)
            cap = CapturedRun(g, res, [frames, boxes, face2frame, self, _template_device(self.template, frames.device)])
            cap._release, ws = self.model.pin_lane(self.device, self.dtype, lane, cap)
            cap._keep.append(ws)
        return cap

    def __call__(self, frames, boxes, face2frame=None):
        """NumPy in / NumPy out convenience (H2D, run, D2H)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        if face2frame is None:
            face2frame = np.zeros(len(boxes), dtype=np.int32)
        with torch.cuda.device(self.device):
            r = self.run_device(torch.from_numpy(frames).to(self.device), torch.from_numpy(boxes).to(self.device),
                                torch.from_numpy(np.ascontiguousarray(face2frame, dtype=np.int32)).to(self.device), True)
            return {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}


    def run_host(self, frames, boxes, face2frame=None, out=None):
        """Host buffers in, host buffers out, one call (the reference's boundary: host arrays, prediction.py:16-113): frames
        uint8 [F,H,W,3] BGR, boxes [B,4], face2frame [B] as NumPy arrays or CPU tensors.  Pinned inputs are copied as they are;
        pageable ones pass through a cached pinned staging buffer first.  Returns NumPy views of cached PINNED result buffers
        {"marks" float32 [B,68,2], "aligned" uint8 [B,oh,ow,3], "M" float64 [B,2,3], "faceboxes" int32 [B,4]} that the next
        run_host call on this pipeline overwrites (pass `out`, a dict of pinned CPU tensors, to own them).  For a steady stream
        of equally shaped batches use HostStream, which overlaps the copies with the compute."""
        dev = self.device
        fr = _as_cpu_tensor(frames, torch.uint8)
        if fr.dim() == 3:
            fr = fr[None]
        bx = _as_cpu_tensor(boxes, torch.int32).reshape(-1, 4)
        B = bx.shape[0]
        ff = torch.zeros(B, dtype=torch.int32) if face2frame is None else _as_cpu_tensor(face2frame, torch.int32)
        st = self._host_state
        with torch.cuda.device(dev):
            d_in = []
            for name, t in (("frames", fr), ("boxes", bx), ("f2f", ff)):
                if not t.is_pinned():
                    stage = st.get(("pin", name))
                    if stage is None or stage.numel() < t.numel():
                        stage = st[("pin", name)] = torch.empty(max(t.numel(), 1), dtype=t.dtype).pin_memory()
                    stage[:t.numel()].view(t.shape).copy_(t)
                    t = stage[:t.numel()].view(t.shape)
                d = st.get(("dev", name))
                if d is None or d.numel() < t.numel():
                    d = st[("dev", name)] = torch.empty(max(t.numel(), 1), dtype=t.dtype, device=dev)
                dv = d[:t.numel()].view(t.shape)
                dv.copy_(t, non_blocking=True)
                d_in.append(dv)
            r = self.run_device(d_in[0], d_in[1], d_in[2], lane=0)
            res = {}
            for k in ("marks", "aligned", "M", "faceboxes"):
                if out is not None and k in out:
                    h = out[k]
                else:
                    h = st.get(("out", k))
                    if h is None or h.numel() < r[k].numel() or h.dtype != r[k].dtype:
                        h = st[("out", k)] = torch.empty(max(r[k].numel(), 1), dtype=r[k].dtype).pin_memory()
                    h = h[:r[k].numel()].view(r[k].shape)
                h.copy_(r[k], non_blocking=True)
                res[k] = h
            torch.cuda.current_stream(dev).synchronize()
        return {k: v.numpy() for k, v in res.items()}


def _as_cpu_tensor(a, dtype):
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda, "host entry points take host arrays"
        return a.contiguous() if a.dtype == dtype else a.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.uint8: np.uint8, torch.int32: np.int32, torch.float32: np.float32}[dtype]))


class HostStream:
    """Streaming host-buffer entry point for equally shaped batches (video: F frames, up to B faces per step): the H2D copy of
    batch k+1, the compute of batch k and the D2H copy of batch k-1 run concurrently on three CUDA streams over `n_slots` slots
    of device buffers, pinned result buffers and (optionally) one captured CUDA graph per slot.

        hs = HostStream(pipe, batch=256, n_frames=4, frame_hw=(1080, 1920))
        t = hs.submit(frames, boxes, face2frame)        # returns at once; pinned inputs are copied as they are
        r = hs.result(t)                                # {"marks", "aligned", "M"}: NumPy views of the slot's pinned buffers,
                                                        # valid until n_slots further submits

    `hs.inputs(slot)` exposes a slot's pinned staging arrays so that a producer (decoder, detector) can write into them
    directly; submit() without arrays then sends that staging area."""

    def __init__(self, pipeline, batch, n_frames, frame_hw, n_slots=3, use_graph=True, channels=3):
        self.pipe, self.B, self.F, self.n_slots = pipeline, int(batch), int(n_frames), int(n_slots)
        dev = self.dev = pipeline.device
        H, W = frame_hw
        oh, ow = pipeline.out_size
        with torch.cuda.device(dev):
            self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.s_comp = [torch.cuda.Stream(dev) for _ in range(self.n_slots)]
            self.ev_in = [torch.cuda.Event() for _ in range(self.n_slots)]
            self.ev_comp = [torch.cuda.Event() for _ in range(self.n_slots)]
            self.ev_out = [torch.cuda.Event() for _ in range(self.n_slots)]
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            p = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self.d_in = [(e((self.F, H, W, channels), torch.uint8), e((self.B, 4), torch.int32), e((self.B,), torch.int32))
                         for _ in range(self.n_slots)]
            self.h_in = [(p((self.F, H, W, channels), torch.uint8), p((self.B, 4), torch.int32), p((self.B,), torch.int32))
                         for _ in range(self.n_slots)]
            self.h_out = [{"marks": p((self.B, 68, 2), torch.float32), "aligned": p((self.B, oh, ow, channels), torch.uint8),
                           "M": p((self.B, 2, 3), torch.float64)} for _ in range(self.n_slots)]
            for hi in self.h_in:                      # a never-filled slot must still hold valid boxes / frame indices
                hi[1][:] = torch.tensor([0, 0, 32, 32], dtype=torch.int32)
                hi[2].zero_()
            for d, hsrc in zip(self.d_in, self.h_in):
                d[0].zero_(); d[1].copy_(hsrc[1]); d[2].copy_(hsrc[2])
            torch.cuda.synchronize(dev)
            self.caps = None
            if use_graph:
                self.caps = [pipeline.capture(*self.d_in[s], lane=s) for s in range(self.n_slots)]
            self._res = [None] * self.n_slots
        self.n_submitted = 0
        self._direct = {}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.d_in[0])
        self.d2h_bytes = sum(t.numel() * t.element_size() for k, t in self.h_out[0].items() if k != "M")

    def inputs(self, slot):
        """Pinned staging arrays of a slot (NumPy views): frames [F,H,W,C], boxes [B,4], face2frame [B]."""
        return tuple(t.numpy() for t in self.h_in[slot % self.n_slots])

    def next_slot(self):
        return self.n_submitted % self.n_slots

    def submit(self, frames=None, boxes=None, face2frame=None, want_matrix=False):
        """Enqueue one batch; returns its ticket.  Arrays that are pinned CPU tensors are sent as they are, other arrays are
        first copied into the slot's pinned staging area (a host memcpy), None sends the staging area as it stands.  Fewer than
        B boxes: the remaining rows keep their previous (valid) content and their results are to be ignored."""
        k = self.n_submitted
        sl = k % self.n_slots
        src = []
        for given, stage in zip((frames, boxes, face2frame), self.h_in[sl]):
            if given is None:
                src.append(stage)
                continue
            direct = self._direct.get(id(given))              # is_pinned() asks the driver: remember the answer per tensor
            if direct is None:
                t = _as_cpu_tensor(given, stage.dtype)
                direct = t if (t.is_pinned() and t.shape == stage.shape) else False
                if isinstance(given, torch.Tensor):
                    if len(self._direct) > 256:
                        self._direct.clear()
                    self._direct[id(given)] = direct
                    self._direct_keep = getattr(self, "_direct_keep", [])[-256:] + [given]   # ids stay unique while referenced
            if direct is not False:
                src.append(direct)
            else:
                t = _as_cpu_tensor(given, stage.dtype)
                if stage.dim() == 2:
                    t = t.reshape(-1, 4)
                self.ev_in[sl].synchronize()          # the slot's previous H2D has read the staging area
                stage[:t.shape[0]].copy_(t)
                src.append(stage)
        dev = self.dev
        with torch.cuda.device(dev):
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_comp[sl])                # the slot's previous compute has consumed its inputs
                for d, h in zip(self.d_in[sl], src):
                    d.copy_(h, non_blocking=True)
                self.ev_in[sl].record(self.s_in)
            comp = self.s_comp[sl]
            with torch.cuda.stream(comp):
                comp.wait_event(self.ev_in[sl])
                comp.wait_event(self.ev_out[sl])                      # the slot's results have been read out
                if self.caps is not None:
                    r = self.caps[sl].replay()
                else:
                    r = self.pipe.run_device(*self.d_in[sl], lane=sl)
                self.ev_comp[sl].record(comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[sl])
                self.h_out[sl]["marks"].copy_(r["marks"], non_blocking=True)
                self.h_out[sl]["aligned"].copy_(r["aligned"], non_blocking=True)
                if want_matrix:
                    self.h_out[sl]["M"].copy_(r["M"], non_blocking=True)
                self.ev_out[sl].record(self.s_out)
        self._src_keep = src                                          # keep caller tensors alive until the copy was enqueued
        self.n_submitted += 1
        return k

    def result(self, ticket):
        """Wait for a ticket's D2H copy and return NumPy views of its slot's pinned result buffers."""
        if ticket < self.n_submitted - self.n_slots:
            raise N.FldError("ticket %d: its slot has been reused (%d slots, %d submitted)" % (ticket, self.n_slots, self.n_submitted))
        sl = ticket % self.n_slots
        self.ev_out[sl].synchronize()
        return {k: v.numpy() for k, v in self.h_out[sl].items()}

    def streams(self):
        return [self.s_in, self.s_out] + self.s_comp

    def drain(self):
        for ev in self.ev_out:
            ev.synchronize()


def shard_faces(n_faces, n_shards):
    """Contiguous split of a face batch over GPUs: [(start, stop)] per shard."""
    per = -(-n_faces // max(n_shards, 1))
    return [(min(i * per, n_faces), min((i + 1) * per, n_faces)) for i in range(n_shards)]


def shard_by_frames(face2frame, n_frames, n_shards):
    """Split for the alignment path ("for C4, split by frame so each frame is uploaded to exactly one GPU"):
    face2frame must be non-decreasing.  Returns [(frame_start, frame_stop, face_start, face_stop)] per shard — contiguous frame
    ranges chosen so that every shard gets about n_faces / n_shards faces and all faces of a frame stay together."""
    f2f = np.asarray(face2frame)
    n = len(f2f)
    assert n == 0 or (np.diff(f2f) >= 0).all(), "face2frame must be sorted (faces grouped by frame)"
    first = np.searchsorted(f2f, np.arange(n_frames + 1), side="left")      # first face of every frame (and n at the end)
    out, f0 = [], 0
    for s in range(n_shards):
        target = (s + 1) * n / n_shards
        f1 = n_frames if s == n_shards - 1 else int(np.clip(np.searchsorted(first, target, side="left"), f0, n_frames))
        if s < n_shards - 1 and f1 > f0 and abs(first[f1 - 1] - target) < abs(first[min(f1, n_frames)] - target):
            f1 -= 1                                                            # the nearer frame boundary
        f1 = max(f1, f0)
        out.append((f0, f1, int(first[f0]), int(first[f1])))
        f0 = f1
    return out


class MultiGpuPipeline:
    """One face batch over several GPUs of one box: the faces are split by frame (shard_by_frames), every GPU gets
    its frames and boxes from the caller's host arrays, runs crop/resize -> CNN -> decode -> fit + warp on its own stream, and
    copies its results into ITS SLICE of one pinned host buffer — the host-side gather; no collective, no peer traffic.  One
    worker thread per GPU drives its device (ctypes calls and CUDA copies release the GIL); weights are replicated at
    construction.  `devices` may name a GPU more than once (two workers then share it)."""

    def __init__(self, model, devices=None, dtype="bfloat16", out_size=(112, 112), template=None):
        import concurrent.futures
        if not torch.cuda.is_available():
            N.handle()
        self.devices = list(range(torch.cuda.device_count())) if devices is None else [torch.device("cuda", d).index if not isinstance(d, int) else d
                                                                                       for d in devices]
        self.out_size = tuple(out_size)
        # a pipeline (result buffers, lane) per worker; the model object is shared (compiled once per device)
        self.pipes = [LandmarkPipeline(model, dtype=dtype, out_size=out_size, template=template, device=d) for d in self.devices]
        # workers sharing a GPU use different lanes; lanes 16.. keep clear of the ones HostStream / captured graphs pin
        self.lanes = [16 + self.devices[:i].count(d) for i, d in enumerate(self.devices)]
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=len(self.devices))
        self._out, self._dev_in = {}, [dict() for _ in self.devices]
        self.streams = [torch.cuda.Stream(torch.device("cuda", d)) for d in self.devices]
        import threading
        self._dev_locks = {d: threading.Lock() for d in set(self.devices)}      # a GPU's handle / net are not thread-safe

    def _host_out(self, name, shape, dtype):
        t = self._out.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self._out[name] = torch.empty(max(n, 1), dtype=dtype).pin_memory()
        return t[:n].view(shape)

    def _worker(self, w, fr, bx, ff, shard, outs):
        f0, f1, i0, i1 = shard
        if i1 <= i0:
            return 0.0
        dev = torch.device("cuda", self.devices[w])
        pipe, st, cache = self.pipes[w], self.streams[w], self._dev_in[w]
        with torch.cuda.device(dev), torch.cuda.stream(st):
            def to_dev(name, src):
                d = cache.get(name)
                if d is None or d.numel() < src.numel():
                    d = cache[name] = torch.empty(src.numel(), dtype=src.dtype, device=dev)
                dv = d[:src.numel()].view(src.shape)
                dv.copy_(src, non_blocking=True)
                return dv
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            d_fr, d_bx = to_dev("frames", fr[f0:f1]), to_dev("boxes", bx[i0:i1])
            d_ff = to_dev("f2f", ff[i0:i1]) - f0
            with self._dev_locks[self.devices[w]]:                               # enqueue only: the GPU work itself runs asynchronously
                r = pipe.run_device(d_fr, d_bx, d_ff.contiguous(), lane=self.lanes[w])
            for k, h in outs.items():
                h[i0:i1].copy_(r[k], non_blocking=True)
            e1.record(st)
            st.synchronize()
            return e0.elapsed_time(e1)

    def run(self, frames, boxes, face2frame):
        """frames uint8 [F,H,W,3], boxes [B,4], face2frame [B] (non-decreasing) as host arrays (pinned CPU tensors are read in
        place by the GPUs' copy engines; NumPy / pageable arrays are staged by the driver).  Returns NumPy views of the pinned
        gather buffers {"marks", "aligned", "M", "faceboxes"} (overwritten by the next run) and sets `last_device_ms` to every
        worker's device time (copies included)."""
        fr = _as_cpu_tensor(frames, torch.uint8)
        if fr.dim() == 3:
            fr = fr[None]
        bx = _as_cpu_tensor(boxes, torch.int32).reshape(-1, 4)
        ff = _as_cpu_tensor(face2frame, torch.int32)
        B, C = bx.shape[0], fr.shape[3]
        oh, ow = self.out_size
        outs = {"marks": self._host_out("marks", (B, 68, 2), torch.float32), "aligned": self._host_out("aligned", (B, oh, ow, C), torch.uint8),
                "M": self._host_out("M", (B, 2, 3), torch.float64), "faceboxes": self._host_out("faceboxes", (B, 4), torch.int32)}
        shards = shard_by_frames(ff.numpy(), fr.shape[0], len(self.devices))
        self.last_shards = shards
        futs = [self.pool.submit(self._worker, w, fr, bx, ff, shards[w], outs) for w in range(len(self.devices))]
        self.last_device_ms = [f.result() for f in futs]
        return {k: v.numpy() for k, v in outs.items()}

    def close(self):
        self.pool.shutdown(wait=True)


# ----------------------------------------------------------------------------------------------- reference API
def detect_marks(img, model, face, dtype=None):
    """Find the 68 facial landmarks of one face (reference prediction.py:16-96).

    img: np.uint8 [H,W,3] BGR; model: landmark_regressor (or any object whose forward_device maps uint8
    [B,128,128,3] RGB to [B,>=136]); face: (x0, y0, x1, y1).  Returns np.uint (68,2) — same dtype/shape.
    Box math, crop, cv2.resize-exact resize, BGR2RGB, CNN, scale-back and the uint cast all run on the GPU.
    A box that leaves the image on the top/left makes the reference raise inside cv2.resize; here it raises
    ValueError (use detect_marks_batch for clipping behaviour)."""
    fbx = _square_box_host(face)
    if fbx[0] < 0 or fbx[1] < 0 or fbx[0] >= img.shape[1] or fbx[1] >= img.shape[0] or fbx[2] <= fbx[0]:
        raise ValueError("face box %s leaves the image on the top/left: empty crop (the reference fails in cv2.resize here)" % (fbx,))
    marks_u = detect_marks_batch(img[None], [face], [0], model, dtype=dtype)[1]
    return marks_u[0].astype(np.uint)


def detect_marks_batch(frames, faces, face2frame, model, dtype=None, device=None):
    """Batched detect_marks: frames uint8 [F,H,W,3] BGR, faces [B,4], face2frame [B] ->
    (marks float32 [B,68,2] pre-cast, marks np.uint64 [B,68,2])."""
    dev = _device(device)
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    boxes = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 4)
    f2f = np.ascontiguousarray(face2frame, dtype=np.int32)
    with torch.cuda.device(dev):
        fr = torch.from_numpy(frames).to(dev)
        crops, fb = preprocess_faces_device(fr, torch.from_numpy(boxes).to(dev), torch.from_numpy(f2f).to(dev), model.input_height)
        out = model.forward_device(crops, dtype)
        marks, marks_u = decode_regress_device(out, fb, True)
        return marks.cpu().numpy(), marks_u.cpu().numpy().astype(np.uint64)


def _square_box_host(face):
    """prediction.py:36-78 on the host (argument checking only; the kernel recomputes it)."""
    x0, y0, x1, y1 = [int(v) for v in face]
    off = int(abs((y1 - y0) * 0.1))
    y0 += off
    y1 += off
    diff = (y1 - y0) - (x1 - x0)
    delta = int(abs(diff) / 2)
    if diff > 0:
        x0 -= delta
        x1 += delta + (1 if diff % 2 == 1 else 0)
    elif diff < 0:
        y0 -= delta
        y1 += delta + (1 if diff % 2 == 1 else 0)
    return [x0, y0, x1, y1]


def align_faces(frames, marks, face2frame=None, template=None, out_size=(112, 112), mode="5pt", device=None, return_matrix=False):
    """New in this build (the reference has no alignment): similarity-align faces.

    frames uint8 [F,H,W,C] or [H,W,C]; marks [B,68,2] (mode "5pt": reduced to eye centres / nose / mouth corners and
    fitted to the 112x112 five-point template) or [B,N,2] with a caller-supplied [N,2] template (mode "full").
    Returns aligned uint8 crops [B,oh,ow,C] (and the fp64 2x3 matrices)."""
    dev = _device(device)
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim == 3:
        frames = frames[None]
    marks = np.ascontiguousarray(marks, dtype=np.float32)
    if marks.ndim == 2:
        marks = marks[None]
    if face2frame is None:
        face2frame = np.zeros(marks.shape[0], dtype=np.int32)
    five = mode == "5pt"
    if not five and template is None:
        raise ValueError("mode='full' needs a template with one row per landmark")
    with torch.cuda.device(dev):
        crops, M = align_device(torch.from_numpy(frames).to(dev), torch.from_numpy(np.ascontiguousarray(face2frame, np.int32)).to(dev),
                                torch.from_numpy(marks).to(dev), template, out_size, five, True)
        crops = crops.cpu().numpy()
        return (crops, M.cpu().numpy()) if return_matrix else crops


def video_predict(facedetector_fn, landmark_model, capture=0, max_frames=None, show=True):
    """reference prediction.py:99-113, with all faces of a frame decoded in one batched GPU call."""
    import cv2
    cap = cv2.VideoCapture(capture)
    n = 0
    while True:
        ok, img = cap.read()
        if not ok:
            break
        rects = list(facedetector_fn(img))
        if rects:
            _, marks_u = detect_marks_batch(img[None], rects, [0] * len(rects), landmark_model)
            for marks in marks_u:
                draw_marks(img, marks.astype(np.int64))
        if show:
            cv2.imshow("image", img)
            if cv2.waitKey(1) & 0xFF == ord('q'):
                break
        n += 1
        if max_frames is not None and n >= max_frames:
            break
    cap.release()
    if show:
        cv2.destroyAllWindows()


def model_from_checkpoint_path(checkpoints_path: str):
    """reference prediction.py:116-133.  The sidecar keeps the reference schema (training.py:195-200) plus the
    input_height/input_width keys the reference loader reads but its trainer never writes (App. D); when they are
    absent the builder defaults are used.  Weights are .npz (see Model.load_weights)."""
    from .networks.basic_models import LANDMARKS_MODELS
    assert (os.path.isfile(checkpoints_path + "_config.json")), "Checkpoint not found."
    model_config = json.loads(open(checkpoints_path + "_config.json", "r").read())
    latest_weights = find_latest_checkpoint(checkpoints_path)
    assert (latest_weights is not None), "Checkpoint not found."
    kwargs = {}
    if model_config.get('input_height') is not None:
        kwargs['input_height'] = model_config['input_height']
    if model_config.get('input_width') is not None:
        kwargs['input_width'] = model_config['input_width']
    model = LANDMARKS_MODELS[model_config['model_class']](model_config['n_classes'], **kwargs)
    print("loaded weights ", latest_weights)
    model.load_weights(latest_weights)
    return model


def keypts_predict(
        model=None,
        inp: typing.Union[np.ndarray, str] = None,
        out_fname: str = None,
        checkpoints_path: str = None, overlay_img: bool = False,
        class_names=None, show_legends: bool = False, colors: typing.List[tuple] = class_colors,
        pred_dim: typing.Tuple[int] = None, read_image_type=1
) -> np.ndarray:
    """reference prediction.py:158-196.  Differences, all where the reference crashes (App. D): the os.path.isdir
    test is only applied to string inputs (a directory is iterated), and the class map is returned."""
    import cv2
    if model is None and checkpoints_path is None:
        raise ValueError("Both model and checkpoint_path cannot be empty")

    if model is None and (checkpoints_path is not None):
        model = model_from_checkpoint_path(checkpoints_path)

    assert (inp is not None), "Invalid input, should be either directory, ndarray or image path"
    assert ((type(inp) is np.ndarray) or isinstance(inp, six.string_types)), \
        "Input should be the CV image or the input file name"

    args = (model.input_width, model.input_height, model.output_height, model.output_width, model.n_classes, colors,
            show_legends, class_names, pred_dim, overlay_img)
    if isinstance(inp, six.string_types) and os.path.isdir(inp):
        results = []
        for name in sorted(os.listdir(inp)):
            if os.path.splitext(name)[1].lower() in (".jpg", ".jpeg", ".png", ".bmp"):
                img = cv2.imread(os.path.join(inp, name), read_image_type)
                of = None if out_fname is None else os.path.join(out_fname, name)
                results.append(_prediction(model, img, *args, of))
        return results

    if isinstance(inp, six.string_types):
        inp = cv2.imread(inp, read_image_type)

    assert (len(inp.shape) == 3 or len(inp.shape) == 1 or len(inp.shape) == 4), "Image should be h,w,3 "
    return _prediction(model, inp, *args, out_fname)


def _prediction(
        model, inp: np.ndarray,
        input_width: int, input_height: int,
        output_height: int, output_width: int,
        n_classes: int, colors: typing.List[typing.Tuple],
        show_legends: bool, class_names: typing.List[str],
        pred_dim: typing.Tuple[int], overlay_img: bool, out_fname: str
):
    """reference prediction.py:199-222: pre-process, forward, per-pixel argmax over classes -> (oh,ow) int64.
    The softmax output stays on the GPU; only the class map is copied back."""
    import cv2
    dev = _device()
    if hasattr(model, "forward_device"):
        x = get_image_array(inp, input_width, input_height, ordering=IMAGE_ORDERING, as_tensor=True)
        with torch.cuda.device(dev):
            pr = model.forward_classmap_device(x[None].contiguous())[0].cpu().numpy()
    else:  # foreign model object with a Keras-style predict()
        x = get_image_array(inp, input_width, input_height, ordering=IMAGE_ORDERING)
        pr = model.predict(np.array([x]))[0]
        pr = pr.reshape((output_height, output_width, n_classes)).argmax(axis=2)

    seg_img = visualize_keypoints(
        pr, inp, n_classes=n_classes,
        colors=colors, overlay_img=overlay_img,
        show_legends=show_legends,
        class_names=class_names,
        pred_dim=pred_dim,
    )

    if out_fname is not None:
        cv2.imwrite(out_fname, seg_img)

    return pr
