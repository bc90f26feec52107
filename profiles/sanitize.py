import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcrg_b200
KC = float(-0.5 * np.log(1 + np.sqrt(2)))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def run(L, R, strip=0, cluster=False, samples=2):
    with mcrg_b200.Context(L, R, seed=5) as ctx:
        ctx.set_couplings([KC]); ctx.init_hot()
        if strip: ctx.set_tuning(strip_rows=strip, use_graphs=0)
        else: ctx.set_tuning(use_graphs=0)
        if cluster: ctx.set_update("cluster")
        ctx.sweep(2); ctx.run(samples, 1, -1, 0); S = ctx.measure(); ctx.sync()
        return int(S.sum())
if which in ("all", "sweep"):
    print("strip L=512 R=16 (TMA)", run(512, 2, strip=16))
    print("strip L=1024", run(1024, 1))
    print("strip L=64 R=8 (cp.async/plain)", run(64, 2, strip=8))
    print("resident L=64", run(64, 3))
    print("resident L=256 (TMA)", run(256, 1))
    print("tiny L=4", run(4, 5))
if which in ("all", "cluster"):
    print("cluster L=128", run(128, 2, cluster=True))
    print("cluster L=16", run(16, 3, cluster=True))
if which in ("all", "rgnn"):
    with mcrg_b200.Context(16, 20, seed=1) as ctx:
        ctx.init_hot(); ctx.rgnn_set_weights([[0.5, -0.2], [0.1, 0.3]]); ctx.rgnn_run(2, 1, 1e-4); print("rgnn", ctx.rgnn_sums().sum())
