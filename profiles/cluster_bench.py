"""Device time of one Swendsen-Wang cluster update (k_sw_init + k_sw_union + k_sw_flip) next to one Metropolis sweep.
Run on the B200:  python profiles/cluster_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mcrg_b200

KC = float(-0.5 * np.log(1 + np.sqrt(2)))

for L, R in ((64, 4096), (1024, 256), (4096, 40), (16384, 2)):
    with mcrg_b200.Context(L, R, seed=3) as ctx:
        ctx.set_couplings([KC])
        ctx.init_hot()
        ctx.set_tuning(strip_rows=0 if L > 512 else 0)
        res = {}
        for mode in ("metropolis", "cluster"):
            ctx.set_update(mode)
            ctx.sweep(30)  # equilibrate a little (cluster sizes matter for the union-find)
            ctx.sync()
            n = 10
            best = 1e9
            for _ in range(3):
                ctx.timer_start()
                ctx.sweep(n)
                best = min(best, ctx.timer_stop())
            res[mode] = best / n
        sites = R * L * L
        print(f"L={L:6d} x {R:5d}: metropolis {res['metropolis']:8.3f} ms/sweep ({sites / res['metropolis'] / 1e6:8.1f} G sites/s)   "
              f"cluster {res['cluster']:8.3f} ms/update ({sites / res['cluster'] / 1e6:8.1f} G sites/s)", flush=True)
