"""Host-link probe for bench.py's e2e leg: pinned H2D / D2H bandwidth alone and concurrently, at the e2e step's sizes."""
import json
import torch

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h_in = torch.empty(4 * 1080 * 1920 * 3, dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty(256 * 112 * 112 * 3, dtype=torch.uint8, device=dev)
h_out = torch.empty_like(d_out, device="cpu").pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


r = {}
t = timed(h2d); r["h2d_ms"] = t; r["h2d_GBps"] = h_in.numel() / t / 1e6
t = timed(d2h); r["d2h_ms"] = t; r["d2h_GBps"] = d_out.numel() / t / 1e6
t = timed(both); r["both_ms"] = t
print(json.dumps(r))
