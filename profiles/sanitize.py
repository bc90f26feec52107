"""compute-sanitizer workload: every kernel of the hot path on small shapes, including the CUDA-graph + side-stream pyramid
form of mcrg_run that bench.py times.  Usage (on the B200):
    compute-sanitizer --tool memcheck  python profiles/sanitize.py all
    compute-sanitizer --tool racecheck python profiles/sanitize.py sweep
    compute-sanitizer --tool racecheck python profiles/sanitize.py requeue      # only the strips tall enough for pass 2 to re-queue
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcrg_b200
KC = float(-0.5 * np.log(1 + np.sqrt(2)))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def run(L, R, strip=0, cluster=False, samples=2, graphs=0, m=1):
    with mcrg_b200.Context(L, R, seed=5) as ctx:
        ctx.set_couplings([KC]); ctx.init_hot()
        ctx.set_tuning(strip_rows=strip, use_graphs=graphs)
        if cluster: ctx.set_update("cluster")
        ctx.sweep(2); ctx.run(samples, m, -1, 0); S = ctx.measure(); ctx.sync()
        return int(S.sum())
if which in ("all", "sweep"):
    print("strip L=512 R=16 (TMA), k_sweep0 + k_tail", run(512, 2, strip=16))
    print("strip L=1024, k_sweep0 + k_level + k_tail", run(1024, 1))
    print("strip L=1024 graphs + side stream, 17 samples", run(1024, 1, samples=17, graphs=1))
    print("strip L=2048 W=32 (whole-pair rows), m=2", run(2048, 1, samples=2, m=2))
    print("strip L=4096 W=64 instantiation", run(4096, 1, samples=1))
    print("strip L=64 R=8 (plain loads)", run(64, 2, strip=8))
    print("resident L=64 (one-warp CTAs)", run(64, 3, samples=3))
    print("resident L=32", run(32, 3, samples=3))
    print("resident L=128", run(128, 2))
    print("resident L=256 (TMA)", run(256, 1))
    print("tiny L=4", run(4, 5))
    print("tiny L=8 (C1)", run(8, 9, samples=3))
if which in ("all", "sweep", "requeue"):
    # strips of 96 rows: a warp walks 24 rows of 32 words per half-sweep and queues ~80 words, i.e. pass 2 runs two one-call
    # batches that re-queue and a last batch that picks the re-queued words up (mc_half_sweep_t<.., REQUEUE = true>)
    print("strip L=4096 R=96, pass 2 with re-queueing", run(4096, 1, strip=96, samples=2))
    print("strip L=1024 R=256 (W=16: 16 row groups), pass 2 with re-queueing", run(1024, 2, strip=256, samples=2))
if which in ("all", "cluster"):
    print("cluster L=128", run(128, 2, cluster=True))
    print("cluster L=16", run(16, 3, cluster=True))
if which in ("all", "rgnn"):
    with mcrg_b200.Context(16, 20, seed=1) as ctx:
        ctx.init_hot(); ctx.rgnn_set_weights([[0.5, -0.2], [0.1, 0.3]]); ctx.rgnn_run(2, 1, 1e-4); print("rgnn", ctx.rgnn_sums().sum())
