"""Host-link probe: H2D bandwidth from pinned memory allocated under different CPU affinities (NUMA placement)."""
import json
import os
import subprocess
import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    for n in sorted(os.listdir("/sys/devices/system/node")):
        if n.startswith("node"):
            print(n, open("/sys/devices/system/node/%s/cpulist" % n).read().strip())
except Exception as e:
    print("no numa info", e)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
props = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
for f in ("local_cpulist", "numa_node", "current_link_speed", "current_link_width", "max_link_speed", "max_link_width"):
    try:
        print(f, open("/sys/bus/pci/devices/%s/%s" % (bus, f)).read().strip())
    except Exception as e:
        print(f, "n/a", e)


def bw(label):
    h = torch.empty(4 * 1080 * 1920 * 3, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty_like(h, device=dev)
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(json.dumps({"affinity": label, "h2d_ms": t, "GBps": h.numel() / t / 1e6}))


all_cpus = sorted(os.sched_getaffinity(0))
bw("default")
half = len(all_cpus) // 2
for label, cpus in (("first-half", all_cpus[:half]), ("second-half", all_cpus[half:])):
    os.sched_setaffinity(0, cpus)
    bw(label)
