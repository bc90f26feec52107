"""Tiny driver for ncu captures: the C4 batch (40 replicas of 4096^2, the five couplings), plain sweeps or measured samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcrg_b200
KS = [-0.4320459, -0.4406868, -0.4496804, -0.4688157, -0.489652]
mode = sys.argv[1] if len(sys.argv) > 1 else "sweep"
with mcrg_b200.Context(4096, 40, seed=12345) as ctx:
    ctx.set_couplings(np.repeat(KS, 8)); ctx.init_hot(); ctx.set_tuning(use_graphs=0)
    ctx.sweep(30)
    if mode == "sweep":
        ctx.sweep(8)
    else:
        ctx.run(8, 1, -1, 0)
    ctx.sync()
