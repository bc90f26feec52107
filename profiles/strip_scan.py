"""Strip-height scan on the B200: measured launch rates against the cost model that picks the strip height (capi.cu:
sweep0_cost).  python profiles/strip_scan.py [--json out.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcrg_b200

KC = float(-0.5 * np.log(1 + np.sqrt(2)))
CASES = [("C3 1024^2 x 256, 4 levels", 1024, 256, 4, [32, 64, 96, 128, 160, 192]),
         ("C4 4096^2 x 40", 4096, 40, -1, [32, 48, 56, 64, 72, 80, 96, 128]),
         ("4096^2 x 5 (C4, one replica per coupling)", 4096, 5, -1, [8, 10, 12, 14, 16, 20, 22, 24, 28, 32, 44, 64]),
         ("4096^2 x 1", 4096, 1, -1, [6, 8, 10, 12, 14, 16, 20, 28]),
         ("C5 16384^2 x 1, 8 levels", 16384, 1, 8, [8, 10, 12, 14, 16, 18, 20]),
         ("16384^2 x 4, 8 levels", 16384, 4, 8, [8, 10, 12, 16, 20])]

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
ap.add_argument("--only", default=None)
a = ap.parse_args()
out = []
for name, L, R, lv, heights in CASES:
    if a.only and a.only not in name:
        continue
    with mcrg_b200.Context(L, R, seed=1) as ctx:
        ctx.set_couplings([KC]); ctx.init_hot(); ctx.sweep(10)
        auto, auto_cost = ctx.strip_plan(1, 0)
        rows = []
        for h in sorted(set(heights + [auto])):
            try:
                _, cost = ctx.strip_plan(1, h)
                if cost < 0:
                    continue
                ctx.set_tuning(strip_rows=h)
                n = 32
                res = {}
                for mode, fn, sweeps in (("sweep", lambda: ctx.sweep(n), n), ("m1", lambda: ctx.run(n, 1, lv, 0), n)):
                    fn(); ctx.sync(); best = 1e9
                    for _ in range(3):
                        ctx.timer_start(); fn(); best = min(best, ctx.timer_stop())
                    res[mode] = R * L * L * sweeps / (best * 1e-3) / 1e9
                rows.append(dict(strip_rows=h, model_cost=cost, sweep_only=res["sweep"], m1=res["m1"], auto=(h == auto)))
                print(f"{name:44s} R={h:4d}{'*' if h == auto else ' '} model {cost:9.1f}  sweep-only {res['sweep']:8.1f}  m=1 {res['m1']:8.1f} G/s", flush=True)
            except mcrg_b200.McrgError as e:
                print(name, h, "error", e)
        out.append(dict(case=name, L=L, replicas=R, auto=auto, rows=rows))
if a.json:
    json.dump(out, open(a.json, "w"), indent=1)
