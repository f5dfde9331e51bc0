"""Model object of the B200 build: a layer list in the vocabulary of the reference's Keras graphs,
host weights in Keras layouts, and per-(GPU, dtype) compiled fld_net instances.

It honours the duck type the reference's callers use (reference
networks/utils.py:31-37, prediction.py:84,128,181-185,208):
  attributes  output_width, output_height, n_classes, input_height, input_width, model_name
  methods     predict(np[B,H,W,C]) -> np, load_weights(path), signatures["predict"](uint8[1,128,128,3])['output']
"""
import ctypes
import json
import os

import numpy as np
import torch

from .. import _native as N

BN_EPS = 1e-3  # Keras BatchNormalization default


class Graph:
    """Records layers; tensor 0 is the input, layer i produces tensor i+1 (include/fld.h)."""

    def __init__(self, input_height, input_width, channels):
        self.input_shape = (int(input_height), int(input_width), int(channels))
        self.layers = []   # dicts: name, op, in0, in1, kh, kw, stride, pads, cout, act, pool, has_bias, has_bn, in_scale
        self.shapes = [self.input_shape]

    def _add(self, **kw):
        self.layers.append(kw)
        return len(self.layers)  # tensor id

    def conv(self, x, name, cout, k, pad=(0, 0, 0, 0), stride=1, act=N.ACT_NONE, pool=0, bias=True, bn=False, in_scale=1.0,
             bn_name=None):
        kh, kw = (k, k) if isinstance(k, int) else k
        if pad == "same":
            th, tw = kh - 1, kw - 1
            pad = (th // 2, th - th // 2, tw // 2, tw - tw // 2)
        h, w, _ = self.shapes[x]
        oh = (h + pad[0] + pad[1] - kh) // stride + 1
        ow = (w + pad[2] + pad[3] - kw) // stride + 1
        if pool:
            oh, ow = oh // 2, ow // 2
        self.shapes.append((oh, ow, cout))
        return self._add(name=name, op=N.OP_CONV, in0=x, in1=-1, kh=kh, kw=kw, stride=stride, pad=tuple(pad), cout=cout,
                         act=act, pool=pool, has_bias=bias, has_bn=bn, in_scale=in_scale,
                         bn_name=bn_name or name.replace("conv", "bn"))

    def deconv(self, x, name, cout, k, stride):
        h, w, _ = self.shapes[x]
        self.shapes.append(((h - 1) * stride + k, (w - 1) * stride + k, cout))
        return self._add(name=name, op=N.OP_DECONV, in0=x, in1=-1, kh=k, kw=k, stride=stride, pad=(0, 0, 0, 0), cout=cout,
                         act=0, pool=0, has_bias=False, has_bn=False, in_scale=1.0)

    def add(self, a, b, name, act=N.ACT_NONE):
        """crop-to-smaller + Add (fcn.py:55-86); act=ACT_RELU is the residual add -> ReLU of resnet50.py:68-69."""
        ha, wa, c = self.shapes[a]
        hb, wb, _ = self.shapes[b]
        self.shapes.append((min(ha, hb), min(wa, wb), c))
        return self._add(name=name, op=N.OP_ADD, in0=a, in1=b, kh=0, kw=0, stride=1, pad=(0, 0, 0, 0), cout=c, act=act, pool=0,
                         has_bias=False, has_bn=False, in_scale=1.0)

    def dwconv(self, x, name, k=3, pad=(1, 1, 1, 1), stride=1, act=N.ACT_NONE, bias=False, bn=True, bn_name=None):
        """ZeroPadding2D + DepthwiseConv2D(depth_multiplier=1) [+BN folded] [+ReLU6] (mobilenet.py:37-47)."""
        h, w, c = self.shapes[x]
        oh = (h + pad[0] + pad[1] - k) // stride + 1
        ow = (w + pad[2] + pad[3] - k) // stride + 1
        self.shapes.append((oh, ow, c))
        return self._add(name=name, op=N.OP_DWCONV, in0=x, in1=-1, kh=k, kw=k, stride=stride, pad=tuple(pad), cout=c, act=act,
                         pool=0, has_bias=bias, has_bn=bn, in_scale=1.0, bn_name=bn_name or name + "_bn")

    def maxpool(self, x, name, k, stride):
        """MaxPooling2D((k,k), strides) padding='valid' (resnet50.py:149)."""
        h, w, c = self.shapes[x]
        self.shapes.append(((h - k) // stride + 1, (w - k) // stride + 1, c))
        return self._add(name=name, op=N.OP_MAXPOOL, in0=x, in1=-1, kh=k, kw=k, stride=stride, pad=(0, 0, 0, 0), cout=c, act=0,
                         pool=0, has_bias=False, has_bn=False, in_scale=1.0)

    def dense(self, x, name, units, act=N.ACT_NONE):
        self.shapes.append((1, 1, units))
        return self._add(name=name, op=N.OP_DENSE, in0=x, in1=-1, kh=0, kw=0, stride=1, pad=(0, 0, 0, 0), cout=units, act=act,
                         pool=0, has_bias=True, has_bn=False, in_scale=1.0)

    def softmax(self, x, name="softmax"):
        self.shapes.append(self.shapes[x])
        return self._add(name=name, op=N.OP_SOFTMAX, in0=x, in1=-1, kh=0, kw=0, stride=1, pad=(0, 0, 0, 0), cout=self.shapes[x][2],
                         act=0, pool=0, has_bias=False, has_bn=False, in_scale=1.0)


def weight_specs(graph):
    """name -> shape for every weight array, in Keras layouts."""
    specs = {}
    for L in graph.layers:
        cin = graph.shapes[L["in0"]][2]
        n = L["name"]
        if L["op"] == N.OP_CONV:
            specs[n + "/kernel"] = (L["kh"], L["kw"], cin, L["cout"])
            if L["has_bias"]:
                specs[n + "/bias"] = (L["cout"],)
            if L["has_bn"]:
                bn = L["bn_name"]
                for s in ("gamma", "beta", "moving_mean", "moving_variance"):
                    specs[bn + "/" + s] = (L["cout"],)
        elif L["op"] == N.OP_DWCONV:
            specs[n + "/depthwise_kernel"] = (L["kh"], L["kw"], cin, 1)
            if L["has_bias"]:
                specs[n + "/bias"] = (cin,)
            if L["has_bn"]:
                for s in ("gamma", "beta", "moving_mean", "moving_variance"):
                    specs[L["bn_name"] + "/" + s] = (cin,)
        elif L["op"] == N.OP_DECONV:
            specs[n + "/kernel"] = (L["kh"], L["kw"], L["cout"], cin)
        elif L["op"] == N.OP_DENSE:
            h, w, c = graph.shapes[L["in0"]]
            specs[n + "/kernel"] = (h * w * c, L["cout"])
            specs[n + "/bias"] = (L["cout"],)
    return specs


class _Signature:
    """model.signatures["predict"](uint8[B,S,S,3]) -> {'output': float32[B,136]} (reference prediction.py:84)."""

    def __init__(self, model):
        self._m = model

    def __call__(self, x, **_):
        x = np.asarray(x)
        return {"output": self._m.predict(x)}


class Model:
    def __init__(self, graph, kind, model_name="", n_classes=None, in_dtype="float32"):
        self.graph = graph
        self.kind = kind                     # "segmentation" | "regression"
        self.model_name = model_name
        self.input_height, self.input_width, self.channels = graph.input_shape
        oh, ow, oc = graph.shapes[-1]
        if kind == "segmentation":
            self.output_height, self.output_width = oh, ow
            self.n_classes = oc if n_classes is None else n_classes
        else:
            self.output_height = self.output_width = 1
            self.n_classes = oc
        self.in_dtype = in_dtype             # "uint8" | "float32"
        # "bf16x3" (default) fp32-accurate tensor cores: 3-term bf16 split, meets the reference-precision bar (0.05 px) |
        # "float32" fp32 CUDA cores, the reference's own arithmetic (TensorFlow fp32, prediction.py:84,208), 13x slower |
        # "bfloat16" plain bf16 tensor cores (0.5 px bar), 2.7x faster than bf16x3
        self.compute_dtype = "bf16x3"
        self.weights = {}
        self._specs = weight_specs(graph)
        self._nets = {}                      # (device index, compute) -> (net handle, ...)
        self._ws = {}                        # (device index, compute, lane) -> torch.uint8 workspace
        self._pins = {}                      # (device index, compute, lane) -> number of live CUDA-graph captures using that workspace
        self._captures = []                  # weak references to live captures (invalidated when the nets are rebuilt)
        self.signatures = {"predict": _Signature(self)}

    # ------------------------------------------------------------------ weights
    def weight_specs(self):
        return dict(self._specs)

    def set_weights(self, weights):
        for k, shp in self._specs.items():
            if k not in weights:
                raise KeyError("missing weight %r" % k)
            a = np.ascontiguousarray(weights[k], dtype=np.float32)
            if tuple(a.shape) != tuple(shp):
                raise ValueError("weight %r has shape %s, expected %s" % (k, a.shape, shp))
            self.weights[k] = a
        self._release()

    def get_weights(self):
        return dict(self.weights)

    def init_weights(self, seed=0, nontrivial_bn=True):
        from .init import random_weights
        self.set_weights(random_weights(self, seed, nontrivial_bn))
        return self

    def save_weights(self, path):
        """`.safetensors` when the path says so, else `.npz` (appended if missing); keys = `layer/variable` in Keras layouts."""
        if path.endswith(".safetensors"):
            from safetensors.numpy import save_file
            save_file({k: np.ascontiguousarray(v) for k, v in self.weights.items()}, path)
        else:
            np.savez(path if path.endswith(".npz") else path + ".npz", **self.weights)

    @staticmethod
    def normalize_weight_names(arrays):
        """Map exported Keras variable names onto this build's `layer/variable` keys: drops the `:0`
        suffix, collapses h5-style `layer/layer/kernel:0` paths, keeps `depthwise_kernel`, `gamma`, `beta`,
        `moving_mean`, `moving_variance`.  Layer names are the reference's own (`block1_conv1`, `conv_dw_3`,
        `res2a_branch2a`, `bn_conv1`, ...), so a checkpoint converted elsewhere with TensorFlow drops in by name."""
        out = {}
        for k, v in arrays.items():
            parts = [p for p in k.split(":")[0].split("/") if p]
            if len(parts) >= 2:
                k2 = parts[-2] + "/" + parts[-1]
            else:
                k2 = parts[0]
            out[k2] = v
        return out

    # Keras auto-names of layers the reference leaves unnamed (fcn.py:10-51 vanilla encoder, :98-121 FCN head), by layer type
    _AUTO_NAMES = {"conv": "conv2d", "bn": "batch_normalization", "deconv": "conv2d_transpose", "dense": "dense",
                   "dwconv": "depthwise_conv2d"}

    def alias_keras_auto_names(self, arrays):
        """The reference gives no names to the vanilla encoder's and the FCN head's layers, so a checkpoint trained with it
        carries Keras' auto-generated names (`conv2d`, `conv2d_1`, ..., `batch_normalization_3`, `conv2d_transpose_2`), numbered
        in creation order with a session-dependent offset.  Layers of this build whose own name is absent from `arrays` are
        matched, per layer type and in graph (= creation) order, to the auto-named layers sorted by their numeric suffix.
        Raises KeyError when the counts do not line up."""
        import re
        have = {k.split("/")[0] for k in arrays}
        want = {"conv": [], "bn": [], "deconv": [], "dense": [], "dwconv": []}
        for L in self.graph.layers:
            kind = {N.OP_CONV: "conv", N.OP_DECONV: "deconv", N.OP_DENSE: "dense", N.OP_DWCONV: "dwconv"}.get(L["op"])
            if kind is None:
                continue
            if L["name"] not in have:
                want[kind].append(L["name"])
            if L.get("has_bn") and L["bn_name"] not in have:
                want["bn"].append(L["bn_name"])
        out = dict(arrays)
        for kind, names in want.items():
            if not names:
                continue
            pat = re.compile(r"^%s(?:_(\d+))?$" % self._AUTO_NAMES[kind])
            autos = sorted({(int(m.group(1) or 0), l) for l in have for m in [pat.match(l)] if m})
            if len(autos) != len(names):
                raise KeyError("cannot alias Keras auto-named %s layers: the file has %d (%s), the model needs %d (%s)"
                               % (kind, len(autos), [a for _, a in autos][:4], len(names), names[:4]))
            for (_, auto), name in zip(autos, names):
                for k, v in arrays.items():
                    if k.split("/")[0] == auto:
                        out[name + "/" + k.split("/", 1)[1]] = v
        return out

    def load_weights(self, path):
        """Weights container of this build: `.npz` or `.safetensors` keyed by layer name in Keras layouts (the
        reference's TF-checkpoint format cannot be read without TensorFlow)."""
        cands = [path, path + ".npz", path + ".safetensors"]
        p = next((c for c in cands if os.path.isfile(c)), None)
        if p is None:
            raise FileNotFoundError("no weights file at %s(.npz|.safetensors)" % path)
        if p.endswith(".safetensors"):
            from safetensors.numpy import load_file
            arrays = load_file(p)
        else:
            with np.load(p) as z:
                arrays = {k: z[k] for k in z.files}
        arrays = self.normalize_weight_names(arrays)
        if any(k not in arrays for k in self._specs):
            arrays = self.alias_keras_auto_names(arrays)
        self.set_weights(arrays)
        return None

    # ------------------------------------------------------------------ device side
    def _release(self):
        """Destroy the compiled nets and workspaces (new weights).  CUDA graphs captured over them hold raw pointers into
        both, so every live capture is invalidated first: replaying it afterwards raises instead of touching freed memory."""
        for ref in self._captures:
            cap = ref()
            if cap is not None:
                cap.invalidate("the model's weights / compiled nets were replaced after the capture")
        self._captures = []
        self._pins = {}
        lib = N.load_library() if self._nets else None
        for (net, _) in self._nets.values():
            lib.fld_net_destroy(net)
        self._nets.clear()
        self._ws.clear()

    def _workspace(self, dev, comp, lane, need, device):
        """The lane's activation workspace (1024-byte aligned base, usable bytes).  Grows on demand — except while a captured
        CUDA graph uses it (its kernels have the old pointer baked in): then a larger batch must use another lane."""
        key = (dev, comp, lane)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need + 1024:
            if ws is not None and self._pins.get(key, 0) > 0:
                raise N.FldError("lane %d's workspace is referenced by a captured CUDA graph and cannot grow to %d bytes: run larger "
                                 "batches on another lane or drop the capture first" % (lane, need))
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        al = (-ws.data_ptr()) % 1024
        return ws, al

    def pin_lane(self, device, dtype, lane, capture):
        """Called by LandmarkPipeline.capture: keep the lane's workspace and the net's kernel plans alive and in place while
        `capture` lives.  Returns a release callback."""
        import weakref
        dev = torch.device(device).index
        comp = self._compute_code(dtype)
        key = (dev, comp, lane)
        net = self.compiled(dev, dtype)
        lib = N.load_library()
        lib.fld_net_retain(net)
        self._pins[key] = self._pins.get(key, 0) + 1
        self._captures.append(weakref.ref(capture))
        ws = self._ws.get(key)
        nets = self._nets

        def release():
            if nets.get((dev, comp), (None,))[0] is net and self._pins.get(key, 0) > 0:   # still the same compiled net
                self._pins[key] -= 1
                lib.fld_net_release(net)
        return release, ws

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _compute_code(self, dtype=None):
        d = dtype or self.compute_dtype
        if d in ("float32", "fp32", torch.float32):
            return N.FLD_F32
        if d in ("bfloat16", "bf16", torch.bfloat16):
            return N.FLD_BF16
        if d in ("bf16x3", "bfloat16x3"):
            return N.FLD_BF16X3
        raise ValueError("compute dtype must be float32, bf16x3 or bfloat16, got %r" % (d,))

    def compiled(self, device=None, dtype=None):
        """fld_net for (device, dtype); built and weight-loaded on first use."""
        if not self.weights:
            raise N.FldError("model has no weights: call init_weights() or load_weights() first")
        lib = N.load_library()
        h = N.handle(device)
        dev = torch.cuda.current_device() if device is None else torch.device(device).index
        comp = self._compute_code(dtype)
        key = (dev, comp)
        if key in self._nets:
            return self._nets[key][0]
        g = self.graph
        descs = (N.LayerDesc * len(g.layers))()
        for i, L in enumerate(g.layers):
            d = descs[i]
            d.op, d.in0, d.in1 = L["op"], L["in0"], L["in1"]
            d.kh, d.kw, d.stride = L["kh"], L["kw"], L["stride"]
            d.pad_t, d.pad_b, d.pad_l, d.pad_r = L["pad"]
            d.cout, d.act, d.pool = L["cout"], L["act"], L["pool"]
            d.has_bias, d.has_bn, d.in_scale = int(L["has_bias"]), int(L["has_bn"]), float(L["in_scale"])
        net = N._vp()
        ih, iw, ic = g.input_shape
        with torch.cuda.device(dev):
            N.check(lib.fld_net_create(h, descs, len(g.layers), ih, iw, ic, N.FLD_U8 if self.in_dtype == "uint8" else N.FLD_F32,
                                       comp, ctypes.byref(net)))
            for i, L in enumerate(g.layers):
                n = L["name"]
                if L["op"] not in (N.OP_CONV, N.OP_DECONV, N.OP_DENSE, N.OP_DWCONV):
                    continue
                k = self.weights[n + ("/depthwise_kernel" if L["op"] == N.OP_DWCONV else "/kernel")]
                b = self.weights.get(n + "/bias") if L["has_bias"] else None
                bn = None
                if L["has_bn"]:
                    bname = L["bn_name"]
                    bn = np.ascontiguousarray(np.concatenate([self.weights[bname + "/" + s] for s in
                                                              ("gamma", "beta", "moving_mean", "moving_variance")]), dtype=np.float32)
                N.check(lib.fld_net_set_weights(net, i, N.host_ptr(k), N.host_ptr(b), N.host_ptr(bn), BN_EPS))
            N.check(lib.fld_net_finalize(net))
        self._nets[key] = (net, h)
        return net

    def set_profiling(self, enable, device=None, dtype=None):
        """Bracket every layer with CUDA events in forward_device (bench.py's live per-kernel timing)."""
        N.check(N.load_library().fld_net_set_profiling(self.compiled(device, dtype), int(enable)))

    def layer_times(self, device=None, dtype=None):
        """[(layer name, ms)] of the last profiled forward (waits for it)."""
        n = len(self.graph.layers)
        buf = (ctypes.c_float * n)()
        rc = N.load_library().fld_net_layer_times(self.compiled(device, dtype), buf, n)
        if rc < 0:
            N.check(rc)
        return [(L["name"], float(buf[i])) for i, L in enumerate(self.graph.layers)]

    def tensor_info(self, tensor, device=None, dtype=None):
        net = self.compiled(device, dtype)
        hwc = (ctypes.c_int32 * 3)()
        dt = N.load_library().fld_net_tensor_shape(net, tensor, hwc)
        if dt < 0:
            N.check(dt)
        return tuple(hwc), dt

    def input_staging(self, B, device, dtype=None, lane=0):
        """Address of the first conv layer's operand staging buffer inside the lane's workspace for batch B, or None when the net
        keeps none (fld_net_input_staging).  A producer that fills it (preprocess_faces_device(..., staging=...)) lets
        forward_device(..., staged=True) skip the network's own widening pass over the crops."""
        lib = N.load_library()
        dev = torch.device(device).index
        if self.in_dtype != "uint8" or B <= 0:
            return None
        with torch.cuda.device(dev):
            net = self.compiled(dev, dtype)
            comp = self._compute_code(dtype)
            ws, al = self._workspace(dev, comp, lane, lib.fld_net_workspace_bytes(net, B), torch.device("cuda", dev))
            stg = N._vp()
            N.check(lib.fld_net_input_staging(net, B, N._vp(ws.data_ptr() + al), ctypes.byref(stg)))
        return stg if stg.value else None

    def forward_device(self, x, dtype=None, out=None, return_workspace=False, lane=0, staged=False):
        """x: CUDA tensor [B,H,W,C] (uint8 or float32 per model.in_dtype) -> float32 CUDA tensor of the final layer.

        `lane` selects one of several independent activation workspaces, so that forwards of independent batches
        enqueued on different CUDA streams can overlap (the weights and kernel plans are shared, the activations not)."""
        lib = N.load_library()
        assert x.is_cuda and x.is_contiguous()
        exp = torch.uint8 if self.in_dtype == "uint8" else torch.float32
        if x.dtype != exp:
            raise TypeError("model expects %s input, got %s" % (exp, x.dtype))
        B = x.shape[0]
        if tuple(x.shape[1:]) != self.graph.input_shape:
            raise ValueError("input shape %s != model input %s" % (tuple(x.shape[1:]), self.graph.input_shape))
        dev = x.device.index
        with torch.cuda.device(dev):
            net = self.compiled(dev, dtype)
            comp = self._compute_code(dtype)
            need = lib.fld_net_workspace_bytes(net, B)
            ws, al = self._workspace(dev, comp, lane, need, x.device)
            base = ws.data_ptr()
            oh, ow, oc = self.graph.shapes[-1]
            if out is None:
                out = torch.empty((B, oh * ow, oc) if self.kind == "segmentation" else (B, oc), dtype=torch.float32, device=x.device)
            fwd = lib.fld_net_forward_staged if staged else lib.fld_net_forward
            N.check(fwd(net, N.ptr(x), B, N._vp(base + al), ws.numel() - al, N.ptr(out), N.stream_ptr(dev)))
        if return_workspace:
            return out, ws, al
        return out

    def forward_classmap_device(self, x, dtype=None, lane=0):
        """Segmentation models: forward + per-pixel argmax over classes (reference prediction.py:208-209) in one call ->
        int64 CUDA [B, oh, ow].  In bfloat16 mode fcn_8's argmax runs inside the last transposed conv's epilogue, so neither
        logits nor probabilities are written to HBM."""
        assert self.kind == "segmentation"
        lib = N.load_library()
        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
        B = x.shape[0]
        if tuple(x.shape[1:]) != self.graph.input_shape:
            raise ValueError("input shape %s != model input %s" % (tuple(x.shape[1:]), self.graph.input_shape))
        dev = x.device.index
        with torch.cuda.device(dev):
            net = self.compiled(dev, dtype)
            comp = self._compute_code(dtype)
            need = lib.fld_net_workspace_bytes(net, B)
            ws, al = self._workspace(dev, comp, lane, need, x.device)
            base = ws.data_ptr()
            cmap = torch.empty((B, self.output_height, self.output_width), dtype=torch.int64, device=x.device)
            N.check(lib.fld_net_forward_classmap(net, N.ptr(x), B, N._vp(base + al), ws.numel() - al, N.ptr(cmap), N.stream_ptr(dev)))
        return cmap

    def forward_landmarks_device(self, x, dtype=None, thresh=0.0, lane=0, n_points=0):
        """Segmentation models: forward + landmark decode of every class channel (reference utils/metrics.py:46-109,
        get_average_xy over transfer_target's channels; n_points < 1 = soft centroid, n_points >= 1 = top-n centroid) in one
        call -> float64 CUDA [B, 2L] = (x0, y0, x1, y1, ...).  In bfloat16 mode the soft centroid is accumulated inside the last
        transposed conv's epilogue (config C3's fused soft-argmax): the probabilities never reach HBM; top-n decodes the
        probabilities materialised in the workspace."""
        assert self.kind == "segmentation"
        lib = N.load_library()
        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
        B = x.shape[0]
        if tuple(x.shape[1:]) != self.graph.input_shape:
            raise ValueError("input shape %s != model input %s" % (tuple(x.shape[1:]), self.graph.input_shape))
        dev = x.device.index
        with torch.cuda.device(dev):
            net = self.compiled(dev, dtype)
            comp = self._compute_code(dtype)
            need = lib.fld_net_landmarks_workspace_bytes(net, B, int(n_points))
            ws, al = self._workspace(dev, comp, lane, need, x.device)
            base = ws.data_ptr()
            xy = torch.empty((B, 2 * self.graph.shapes[-1][2]), dtype=torch.float64, device=x.device)
            N.check(lib.fld_net_forward_landmarks(net, N.ptr(x), B, N._vp(base + al), ws.numel() - al, int(n_points), float(thresh),
                                                  N.ptr(xy), N.stream_ptr(dev)))
        return xy

    def intermediate(self, x, tensor, dtype=None):
        """Run forward and return intermediate tensor `tensor` ([B,h,w,c] float32 CUDA) — parity checks of the levels."""
        out, ws, al = self.forward_device(x, dtype, return_workspace=True)
        if tensor == len(self.graph.layers):   # the final tensor is written straight into `out`, not the workspace
            return out
        lib = N.load_library()
        net = self.compiled(x.device.index, dtype)
        (h, w, c), dt = self.tensor_info(tensor, x.device.index, dtype)
        off = lib.fld_net_tensor_offset(net, tensor, x.shape[0])
        B = x.shape[0]
        n = B * h * w * c
        if dt == N.FLD_BF16X3:   # SPLIT tensor of the bf16x3 mode: per pixel [hi(c) | lo(c)] bf16, value = hi + lo
            raw = ws[al + off: al + off + n * 4].view(torch.bfloat16).view(B, h, w, 2, c).float()
            return raw[:, :, :, 0] + raw[:, :, :, 1]
        tdt = {N.FLD_F32: torch.float32, N.FLD_BF16: torch.bfloat16, N.FLD_U8: torch.uint8}[dt]
        esz = {N.FLD_F32: 4, N.FLD_BF16: 2, N.FLD_U8: 1}[dt]
        raw = ws[al + off: al + off + n * esz]
        return raw.view(tdt).view(B, h, w, c).float()

    def predict(self, x, batch_size=None, dtype=None, device=None):
        """Keras-style predict on a NumPy batch [B,H,W,C]; returns NumPy (segmentation: [B, oh*ow, n_classes]
        softmax probabilities as networks/utils.py:28-30; regression: [B, 136])."""
        x = np.ascontiguousarray(x, dtype=np.uint8 if self.in_dtype == "uint8" else np.float32)
        if x.ndim == 3:
            x = x[None]
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index) \
            if torch.cuda.is_available() else None
        if dev is None:
            N.handle()  # raises the no-GPU error
        B = x.shape[0]
        if batch_size is None:
            oh, ow, oc = self.graph.shapes[-1]
            per_item = 64 * self.input_height * self.input_width * 4 + oh * ow * oc * 16
            batch_size = max(1, min(B, int(8e9 // max(per_item, 1))))
        outs = []
        for s in range(0, B, batch_size):
            xt = torch.from_numpy(x[s:s + batch_size]).to(dev, non_blocking=False)
            outs.append(self.forward_device(xt, dtype).cpu().numpy())
        return np.concatenate(outs, axis=0)

    # ------------------------------------------------------------------ config sidecar (reference training.py:187-200)
    def config_dict(self, model_class=None):
        return {"model_class": model_class or self.model_name, "n_classes": int(self.n_classes),
                "input_height": int(self.input_height), "input_width": int(self.input_width),
                "output_height": int(self.output_height), "output_width": int(self.output_width)}

    def save_config(self, checkpoints_path, model_class=None):
        with open(checkpoints_path + "_config.json", "w") as f:
            json.dump(self.config_dict(model_class), f)
