"""Throughput of the hot path on the BASELINE.json configs (device-resident, CUDA-event timed).  Run on the B200:
    python profiles/configs_bench.py [--json out.json]
C1 L=8 / C2 N=64 x 4096 replicas / C3 L=1024 x 256 replicas, 4 levels / C4 L=4096 x 40 / C5 L=16384 x 1, 8 levels."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mcrg_b200

KC = float(-0.5 * np.log(1 + np.sqrt(2)))
CONFIGS = [("C1 L=8 x 65535 replicas", 8, 65535, -1), ("C2 N=64 x 4096 replicas", 64, 4096, -1),
           ("L=128 x 4096 replicas (main.cpp shape)", 128, 4096, -1), ("C3 L=1024 x 256 replicas, 4 levels", 1024, 256, 4),
           ("C4 L=4096 x 40 replicas", 4096, 40, -1), ("L=4096 x 1 replica", 4096, 1, -1),
           ("C5 L=16384 x 1 replica, 8 levels", 16384, 1, 8), ("L=16384 x 4 replicas, 8 levels", 16384, 4, 8)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--samples", type=int, default=0, help="samples / sweeps per timed call (default 64 for L <= 1024, else 32)")
    a = ap.parse_args()
    rows = []
    for name, L, R, lv in CONFIGS:
        if a.only and a.only not in name:
            continue
        with mcrg_b200.Context(L, R, seed=1) as ctx:
            ctx.set_couplings([KC])
            ctx.init_hot()
            ctx.sweep(20)
            n = a.samples if a.samples > 0 else (64 if L <= 1024 else 32)
            res = {}
            for mode, fn in (("sweep_only", lambda: ctx.sweep(n)), ("m=1", lambda: ctx.run(n, 1, lv, 0)), ("m=16", lambda: ctx.run(max(n // 8, 2), 16, lv, 0))):
                fn()
                ctx.sync()
                best = 1e9
                for _ in range(3):
                    ctx.timer_start()
                    fn()
                    best = min(best, ctx.timer_stop())
                sweeps = n if mode != "m=16" else max(n // 8, 2) * 16
                res[mode] = R * L * L * sweeps / (best * 1e-3) / 1e9
            rows.append(dict(config=name, L=L, replicas=R, G_attempts_per_s=res))
            print(f"{name:42s} sweep-only {res['sweep_only']:8.1f}   m=1 {res['m=1']:8.1f}   m=16 {res['m=16']:8.1f}  G attempts/s", flush=True)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
