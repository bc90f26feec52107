"""The path bench.py times, pinned to the oracle: mcrg_run on the STRIP kernels with CUDA graphs (16 samples per graph),
the pyramid on the side stream with the level-1 lattice / popcount cells double-buffered by sample parity, several graph
replays plus a remainder, a second call that takes the cached-graph branch, several couplings and replicas — every live
accumulator slot and the final configurations against cpu_run (mcrg.cpp:72-98 restated with the oracle).

The oracle's scalar sampler costs 0.17 s (L = 1024) to 3.8 s (L = 4096) per sample and replica, so the replicas are
computed on host threads (ctypes releases the GIL).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import _libs
from test_gpu_parity import KC, add_runs, assert_accumulators, cpu_run, oracle_hot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mc():
    import mcrg_b200

    assert mcrg_b200.capi.device_count() >= 1
    return mcrg_b200


# (L, couplings per replica, samples of the first call, of the second call, sweeps per sample, max_levels, strip rows)
CASES = [
    (1024, [KC, -0.43, 0.37], 35, 17, 1, -1, 0),      # two replays + 3; then cached graph + 1; K > 0 among the couplings
    (1024, [KC, -0.46, -0.30], 33, 16, 2, 4, 0),      # m = 2 (an extra k_sweep0<false> per sample), capped pyramid (C3 shape)
    (2048, [KC, -0.4497, -0.4688], 33, 16, 1, -1, 0),
    (4096, [KC, -0.4320459], 17, 0, 1, -1, 0),         # the headline shape: one graph + 1 sample
    (512, [KC, -0.42], 34, 18, 1, -1, 32),             # forced strips below the resident limit (one k_tail-only pyramid)
    (1024, [KC, -0.45], 18, 16, 1, -1, 44),            # strips that do not divide L, unaligned to the tie-coin chunks
    (1024, [KC, -0.4320459], 81, 64, 1, -1, 0),        # 64-sample graph + 16-sample graph + 1; then the cached 64-sample graph
]


@pytest.mark.parametrize("L,Ks,n1,n2,m,max_levels,strip", CASES)
def test_graph_replay_strip_run_matches_oracle(mc, L, Ks, n1, n2, m, max_levels, strip):
    seed, base, t0 = 77001, 4, (1 << 32) - 20  # the counter crosses 2^32 inside the first call
    R = len(Ks)
    with ThreadPoolExecutor(max_workers=R) as pool:
        first = [pool.submit(cpu_run, L, seed, base + r, Ks[r], t0, n1, m, max_levels, oracle_hot(L, seed, base + r)) for r in range(R)]
        with mc.Context(L, R, seed=seed, replica_base=base, n_bins=2) as ctx:
            ctx.set_tuning(strip_rows=strip, fuse_sweeps=1, use_graphs=1)  # graphs on, overlap on (default)
            ctx.set_couplings(Ks)
            ctx.init_hot()
            ctx.sweep_counter = t0
            ctx.run(n1, m, max_levels, bin=1)
            acc1, accd1 = ctx.accumulators()
            spins1 = ctx.get_spins()
            assert ctx.sweep_counter == t0 + n1 * m
            if n2:
                ctx.run(n2, m, max_levels, bin=1)  # >= 16 samples: replays the graph cached by the first call
                acc2, accd2 = ctx.accumulators()
                spins2 = ctx.get_spins()
                S_last = ctx.measure(max_levels)
        want1 = [f.result() for f in first]
        for r in range(R):
            assert_accumulators(mc, acc1[r, 1], accd1[r, 1], want1[r], (L, r, "first call"))
            assert np.array_equal(spins1[r], want1[r]["final"]), (L, r)
        assert (acc1[:, 0, :] == 0).all()
        if n2:
            second = [pool.submit(cpu_run, L, seed, base + r, Ks[r], t0 + n1 * m, n2, m, max_levels, want1[r]["final"]) for r in range(R)]
            for r in range(R):
                want2 = add_runs(want1[r], second[r].result())
                assert_accumulators(mc, acc2[r, 1], accd2[r, 1], want2, (L, r, "second call"))
                assert np.array_equal(spins2[r], want2["final"]), (L, r)
                # a measurement after the replays reads the right ping-pong buffers
                assert np.array_equal(S_last[r], _libs.pyramid(L, want2["final"], seed, base + r, t0 + (n1 + n2) * m, max_levels)), (L, r)


@pytest.mark.parametrize("slots", ["1", "3", "16"])
def test_results_do_not_depend_on_the_number_of_pyramid_slots(mc, monkeypatch, slots):
    """Samples in flight (sets of blocked lattices / popcount cells, one side stream each): with fewer slots than a graph has
    samples a measuring sweep waits for the pyramid that frees its slot; the default (64 at this size) never does inside a graph.
    Same accumulators, same configurations, same last measurement — the default is pinned to the oracle above."""
    L, R, n = 1024, 3, 85
    out = []
    for env in (None, slots):
        if env is None:
            monkeypatch.delenv("MCRG_SLOTS", raising=False)
        else:
            monkeypatch.setenv("MCRG_SLOTS", env)
        with mc.Context(L, R, seed=31, n_bins=1) as ctx:
            ctx.set_tuning(use_graphs=1)
            ctx.set_couplings([KC, -0.43, -0.46])
            ctx.init_hot()
            ctx.run(n, 1, -1, 0)
            ctx.run(17, 2, -1, 0)
            acc, accd = ctx.accumulators()
            out.append((acc, accd, ctx.measure(-1), ctx.get_spins()))
    monkeypatch.delenv("MCRG_SLOTS", raising=False)
    a, b = out
    assert (a[0] == b[0]).all() and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_graph_run_equals_plain_launches_at_full_size(mc):
    """L = 4096, 40 replicas x 5 couplings (exactly bench.py's context), 48 samples: graphs + side-stream pyramid + programmatic
    dependent launch against the same run with plain launches on one stream, and against graphs without programmatic launch —
    every accumulator and every spin identical (the oracle cannot do this volume; the plain-launch path is the one the
    oracle-pinned tests above and in test_gpu_parity.py cover at L <= 4096)."""
    import bench

    L, R, n = 4096, 40, 48
    Ks = np.repeat(bench.TRAIN_KS, 8)
    out = []
    for graphs, overlap, pdl in ((1, "1", "3"), (0, "0", "0"), (1, "1", "0"), (1, "1", "7")):  # MCRG_PDL bits: sweep, measuring sweep, pyramid
        import os

        os.environ["MCRG_OVERLAP"] = overlap
        os.environ["MCRG_PDL"] = pdl
        try:
            with mc.Context(L, R, seed=12345) as ctx:
                ctx.set_tuning(use_graphs=graphs)
                ctx.set_couplings(Ks)
                ctx.init_hot()
                ctx.sweep(2)
                ctx.run(n, 1, -1, 0)
                ctx.run(n, 1, -1, 0)
                acc, accd = ctx.accumulators()
                obs = ctx.observables()
                out.append((acc, accd, obs, ctx.get_spins(0, 2)))
        finally:
            del os.environ["MCRG_OVERLAP"]
            del os.environ["MCRG_PDL"]
    a = out[0]
    for b in out[1:]:
        assert (a[0] == b[0]).all() and np.array_equal(a[1], b[1])
        for k in ("Snn", "Snnn", "Splaq", "M"):
            assert np.array_equal(a[2][k], b[2][k])
        assert np.array_equal(a[3], b[3])
    lay = mc.capi.acc_layout()
    assert all(a[0][r, 0, lay.slot_n] == 2 * n for r in range(R))
