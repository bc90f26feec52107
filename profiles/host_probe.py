"""What the host side of the e2e path can do on this box (run on the GPU box): topology, the int32 -> bit packing rate
(mcrg_host_pack_i32_colmajor) and a plain streaming read of the same pinned buffer per thread count, and the pinned
host -> device copy rate.  Output: one JSON object (also written to --json)."""
import argparse, json, os, sys, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mcrg_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
ap.add_argument("--replicas", type=int, default=16)
a = ap.parse_args()
L, R = 4096, a.replicas
out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    nodes[p.split("/")[-2]] = open(p).read().strip()
out["numa_cpulists"] = nodes
try:
    import pynvml as nv
    nv.nvmlInit()
    g = []
    for i in range(nv.nvmlDeviceGetCount()):
        h = nv.nvmlDeviceGetHandleByIndex(i)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = None
        sp = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        if os.path.exists(sp):
            node = int(open(sp).read())
        g.append({"index": i, "bus": bus, "numa_node": node})
    out["gpus"] = g
except Exception as e:
    out["gpus"] = repr(e)
host = torch.empty((R, L, L), dtype=torch.int32).pin_memory()
host.numpy()[:] = 1
host.numpy()[:, ::3, ::5] = -1
pk = torch.empty(capi.packed_words(L, R), dtype=torch.int32).pin_memory()
gb = R * L * L * 4 / 1e9
rates = {}
for nt in (1, 2, 4, 8, 16, 32, 64):
    if nt > (os.cpu_count() or 1):
        break
    best_p = best_r = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); capi.host_pack(host.data_ptr(), L, R, pk.data_ptr(), nt); best_p = min(best_p, time.perf_counter() - t0)
        t0 = time.perf_counter(); capi.host_read_probe(host.data_ptr(), R * L * L, nt); best_r = min(best_r, time.perf_counter() - t0)
    rates[nt] = {"pack_GBps": gb / best_p, "read_GBps": gb / best_r}
    print(nt, rates[nt], flush=True)
out["host_rates_by_threads"] = rates
dev = torch.empty((R, L, L), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.copy_(host, non_blocking=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
out["h2d_pinned_GBps"] = gb / best
print(json.dumps(out))
if a.json:
    json.dump(out, open(a.json, "w"), indent=1)
