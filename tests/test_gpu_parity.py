"""GPU parity tests: the CUDA library (through its C ABI) against the CPU oracle on the same inputs and the same
Philox keys.  Integer / bit work => bit-exact.  Run on the B200 box: pytest -m gpu.
"""
import ctypes as C

import numpy as np
import pytest

import _libs

pytestmark = pytest.mark.gpu

KC = -0.5 * np.log(1 + np.sqrt(2))


@pytest.fixture(scope="module")
def mc():
    import mcrg_b200

    # the product must be the CUDA path: fail loudly if the library or the device is missing
    assert mcrg_b200.capi.device_count() >= 1
    return mcrg_b200


def oracle_hot(L, seed, replica):
    s = np.zeros((L, L), np.int32)
    _libs.oracle().orc_hot_start(L, seed, replica, s)
    return s


# ---------------------------------------------------------------------------------------------------------
# transport and initial conditions
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,R", [(2, 4), (4, 3), (8, 5), (16, 2), (32, 3), (64, 4), (128, 3), (256, 2), (1024, 2)])
def test_pack_unpack_roundtrip(mc, L, R):
    spins = np.stack([_libs.random_lattice(L, 10 + r) for r in range(R)])
    with mc.Context(L, R) as ctx:
        ctx.set_spins(spins)
        assert np.array_equal(ctx.get_spins(), spins)
        # partial range
        ctx.set_spins(-spins[:1], first=R - 1)
        got = ctx.get_spins()
        assert np.array_equal(got[R - 1], -spins[0])
        assert np.array_equal(got[: R - 1], spins[: R - 1])


@pytest.mark.parametrize("L,R", [(2, 3), (4, 2), (16, 5), (32, 3), (64, 2), (256, 3), (1024, 2)])
def test_packed_and_pipelined_uploads_roundtrip(mc, L, R):
    """The two faster upload paths deliver the same lattices as mcrg_set_spins_i32_colmajor: host-packed bits
    (mcrg_host_pack_i32_colmajor + mcrg_set_spins_packed) and the copy-stream pipeline (_begin / _commit)."""
    import torch

    rng = np.random.default_rng(7 * L + R)
    spins = np.where(rng.random((R, L, L)) < 0.5, 1, -1).astype(np.int32)
    pinned = torch.from_numpy(spins.copy()).pin_memory()
    packed = np.zeros(mc.capi.packed_words(L, R), np.uint32)
    mc.capi.host_pack(spins.ctypes.data, L, R, packed.ctypes.data, 2)
    with mc.Context(L, R + 1, seed=1) as ctx:
        ctx.set_spins_packed_ptr(packed.ctypes.data, R, first=1)
        ctx.sync()
        assert np.array_equal(ctx.get_spins(1, R), spins)
        assert (ctx.get_spins(0, 1) == 1).all()  # untouched replica keeps its cold start
        ctx.init_cold()
        ctx.set_spins_begin(pinned.data_ptr(), R, first=0)
        ctx.sweep(1)  # work in flight while the copy runs; replica R only matters for the check below
        ctx.set_spins_commit()
        ctx.sync()
        assert np.array_equal(ctx.get_spins(0, R), spins)
        with pytest.raises(mc.capi.McrgError):
            ctx.set_spins_commit()  # nothing in flight
        # the same pipeline with host-packed words, then an int32 upload again (the commit must pick the right unpacker)
        pk = torch.from_numpy(packed.view(np.int32).copy()).pin_memory()
        ctx.init_cold()
        ctx.set_spins_packed_begin(pk.data_ptr(), R, first=1)
        ctx.sweep(1)
        ctx.set_spins_commit()
        ctx.sync()
        assert np.array_equal(ctx.get_spins(1, R), spins)
        ctx.set_spins_begin(pinned.data_ptr(), R, first=0)
        ctx.set_spins_commit()
        ctx.sync()
        assert np.array_equal(ctx.get_spins(0, R), spins)
        # one upload of each kind in flight at once (the hybrid e2e path of bench.py): replica 0 as int32 through the copy
        # engine, replicas 1.. host-packed; one commit takes in both; a second begin of the same kind is refused
        if R >= 2:
            ctx.init_cold()
            rest = np.zeros(mc.capi.packed_words(L, R - 1), np.uint32)
            mc.capi.host_pack(spins[1:].ctypes.data, L, R - 1, rest.ctypes.data, 1)
            pk_rest = torch.from_numpy(rest.view(np.int32).copy()).pin_memory()
            ctx.set_spins_begin(pinned.data_ptr(), 1, first=0)
            ctx.set_spins_packed_begin(pk_rest.data_ptr(), R - 1, first=1)
            with pytest.raises(mc.capi.McrgError):
                ctx.set_spins_begin(pinned.data_ptr(), 1, first=0)
            ctx.sweep(1)
            ctx.set_spins_commit()
            ctx.sync()
            assert np.array_equal(ctx.get_spins(0, R), spins)


@pytest.mark.parametrize("L", [2, 4, 8, 32, 64, 128, 512])
def test_hot_start_matches_oracle(mc, L):
    with mc.Context(L, 3, seed=777, replica_base=5) as ctx:
        ctx.init_hot()
        got = ctx.get_spins()
        for r in range(3):
            assert np.array_equal(got[r], oracle_hot(L, 777, 5 + r))
        ctx.init_cold()
        assert (ctx.get_spins() == 1).all()


def test_bad_arguments_are_reported(mc):
    for L in (0, 1, 3, 6, 100, 32768):
        with pytest.raises(mc.McrgError):
            mc.Context(L, 1)
    with pytest.raises(mc.McrgError):
        mc.Context(8, 0)
    with mc.Context(8, 2) as ctx:
        with pytest.raises(mc.McrgError):
            ctx.set_spins(np.ones((3, 8, 8), np.int32))
        with pytest.raises(mc.McrgError):
            ctx.get_level_spins(0, 1)  # nothing measured yet
        with pytest.raises(mc.McrgError):
            ctx.run(1, 1, -1, bin=1)
        with pytest.raises(mc.McrgError):
            ctx.set_couplings([KC, KC, KC])
        with pytest.raises(mc.McrgError):
            ctx.set_tuning(strip_rows=3)


# ---------------------------------------------------------------------------------------------------------
# deterministic observables, block spins, pyramid   (bit-exact)
# ---------------------------------------------------------------------------------------------------------

def lattice_cases(L):
    cases = dict(_libs.pattern_lattices(L))
    cases["rand0"] = _libs.random_lattice(L, 1)
    cases["rand1"] = _libs.random_lattice(L, 2)
    cases["biased"] = _libs.random_lattice(L, 3, p_up=0.8)
    if L <= 256:
        cases["clustered"] = _libs.clustered_lattice(L, 4, n_sweeps=10)
    return cases


@pytest.mark.parametrize("L", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024])
def test_observables_match_oracle(mc, L):
    o = _libs.oracle()
    cases = lattice_cases(L)
    spins = np.stack(list(cases.values()))
    with mc.Context(L, len(cases)) as ctx:
        ctx.set_spins(spins)
        obs = ctx.observables()
    for r, (name, s) in enumerate(cases.items()):
        want = np.zeros(2, np.int64)
        o.orc_calc_interactions(L, s, want)
        assert obs["Snn"][r] == want[0], (L, name)
        assert obs["Snnn"][r] == want[1], (L, name)
        assert obs["Splaq"][r] == o.orc_plaquette(L, s), (L, name)
        assert obs["M"][r] == int(s.sum()), (L, name)


@pytest.mark.parametrize("L,strip", [(4, 0), (8, 2), (16, 0), (32, 8), (64, 0), (128, 16), (256, 0), (512, 0), (1024, 0),
                                     (2048, 32), (64, 6), (128, 10), (512, 20), (1024, 44), (2048, 18), (2048, 6)])
def test_pyramid_matches_oracle(mc, L, strip):
    """strip = 0: the library's own choice; otherwise forced strips, including heights that do not divide L (ragged last
    strip) and that are not aligned to the tie-coin chunks of the measurement loop."""
    seed, base, t = 4242, 9, (1 << 34) + 17
    cases = lattice_cases(L)
    if L >= 1024:
        cases = {k: cases[k] for k in ("rand0", "checker", "stripes_i")}
    spins = np.stack(list(cases.values()))
    with mc.Context(L, len(cases), seed=seed, replica_base=base) as ctx:
        ctx.set_tuning(strip_rows=strip)
        ctx.set_spins(spins)
        ctx.sweep_counter = t
        S = ctx.measure()
        n_lv = S.shape[1] - 1
        assert n_lv == int(np.log2(L)) - 1
        for r, (name, s) in enumerate(cases.items()):
            want_S, want_lv = _libs.pyramid(L, s, seed, base + r, t, -1, want_levels=True)
            assert np.array_equal(S[r], want_S), (L, name, S[r], want_S)
            for k in range(1, n_lv + 1):
                assert np.array_equal(ctx.get_level_spins(r, k), want_lv[k - 1]), (L, name, k)
        # the configuration itself is untouched by a measurement
        assert np.array_equal(ctx.get_spins(), spins)
        # capped pyramid
        S3 = ctx.measure(max_levels=min(2, n_lv))
        assert np.array_equal(S3, S[:, : min(2, n_lv) + 1])


def test_golden_reference_vectors(mc):
    """Correlators / energy / magnetisation / non-tie block spins produced by the reference itself
    (tests/golden/deterministic.npz, made by tests/golden/make_golden.py from the compiled reference)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "deterministic.npz"))
    by_N = {}
    for key in g["index"]:
        N = int(str(key).split("_")[0][1:])
        by_N.setdefault(N, []).append(str(key))
    for N, keys in by_N.items():
        spins = np.stack([np.where(np.unpackbits(g[k + "_bits"])[: N * N].reshape(N, N) > 0, 1, -1) for k in keys]).astype(np.int32)
        with mc.Context(N, len(keys), seed=1) as ctx:
            ctx.set_spins(spins)
            obs = ctx.observables()
            if N >= 4:
                ctx.measure(max_levels=1)
            for r, k in enumerate(keys):
                ref = g[k + "_ref"]
                assert obs["Snn"][r] == ref[0] and obs["Snnn"][r] == ref[1] and obs["Snn"][r] == ref[2], k
                assert obs["M"][r] == ref[5], k
                # IsingModel::calc_energy = K*S_nn/N^2 up to the reference's own rounding (ising.cpp:160-172)
                assert abs(KC * obs["Snn"][r] / (N * N) - ref[3]) <= 1e-12 * max(1.0, abs(ref[3])), k
                # IsingModel::calc_magnetization: integer division (ising.cpp:178)
                assert float(int(obs["M"][r] / (N * N))) == ref[4], k
                if N < 4:
                    continue
                want_blk = np.where(np.unpackbits(g[k + "_block_bits"])[: (N // 2) ** 2].reshape(N // 2, N // 2) > 0, 1, -1)
                got_blk = ctx.get_level_spins(r, 1)
                nontie = spins[r].reshape(N // 2, 2, N // 2, 2).sum(axis=(1, 3)) != 0
                assert np.array_equal(got_blk[nontie], want_blk[nontie]), k


def test_supplied_ties_reproduce_the_reference_block_lattice(mc):
    """tie_mode = supplied (mcrg_measure_supplied): with the REFERENCE's own tie draws handed to the device, the level-1
    lattice equals the reference's block_spin_transformation output on EVERY block, ties included
    (tests/golden/deterministic.npz: block lattices made by the compiled reference with its global rng, mcrg.cpp:314-348)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "deterministic.npz"))
    by_N = {}
    for key in g["index"]:
        N = int(str(key).split("_")[0][1:])
        if N >= 4:
            by_N.setdefault(N, []).append(str(key))
    for N, keys in by_N.items():
        spins = np.stack([np.where(np.unpackbits(g[k + "_bits"])[: N * N].reshape(N, N) > 0, 1, -1) for k in keys]).astype(np.int32)
        blocks = np.stack([np.where(np.unpackbits(g[k + "_block_bits"])[: (N // 2) ** 2].reshape(N // 2, N // 2) > 0, 1, -1) for k in keys])
        with mc.Context(N, len(keys), seed=1) as ctx:
            ctx.set_spins(spins)
            ctx.measure_supplied([blocks], max_levels=1)  # the reference's outputs serve as the coins of its tied blocks
            for r, k in enumerate(keys):
                assert np.array_equal(ctx.get_level_spins(r, 1), blocks[r]), k


@pytest.mark.parametrize("L,strip", [(4, 0), (8, 0), (64, 0), (256, 0), (256, 16), (1024, 0), (2048, 44)])
def test_supplied_ties_through_the_whole_pyramid(mc, L, strip):
    """Random caller-supplied coins at every level, against the oracle's orc_block_spin_supplied chained down the pyramid:
    every level's block spins and correlators (strip kernel, k_level, k_tail and the one-warp tail)."""
    o = _libs.oracle()
    rng = np.random.default_rng(L + strip)
    cases = {"rand": _libs.random_lattice(L, 5), "checker": _libs.pattern_lattices(L)["checker"], "stripes": _libs.pattern_lattices(L)["stripes_i"]}
    R = len(cases)
    n_lv = int(np.log2(L)) - 1
    coins = [np.where(rng.random((R, L >> lv, L >> lv)) < 0.5, 1, -1).astype(np.int32) for lv in range(1, n_lv + 1)]
    with mc.Context(L, R, seed=3) as ctx:
        ctx.set_tuning(strip_rows=strip)
        ctx.set_spins(np.stack(list(cases.values())))
        S = ctx.measure_supplied(coins)
        for r, (name, s) in enumerate(cases.items()):
            cur = np.ascontiguousarray(s)
            for lv in range(0, n_lv + 1):
                n = L >> lv
                want = np.zeros(2, np.int64)
                o.orc_calc_interactions(n, cur, want)
                assert S[r, lv, 0] == want[0] and S[r, lv, 1] == want[1], (L, name, lv)
                assert S[r, lv, 2] == o.orc_plaquette(n, cur) and S[r, lv, 3] == int(cur.sum()), (L, name, lv)
                if lv == n_lv:
                    break
                nxt = np.zeros((n // 2, n // 2), np.int32)
                mask = np.zeros((n // 2, n // 2), np.int32)
                o.orc_block_spin_supplied(n, 2, cur, np.ascontiguousarray(coins[lv][r]), nxt, mask)
                assert np.array_equal(ctx.get_level_spins(r, lv + 1), nxt), (L, name, lv + 1)
                cur = nxt
        # the Philox mode is untouched by a supplied-mode call
        assert np.array_equal(ctx.measure()[0], _libs.pyramid(L, list(cases.values())[0], 3, 0, 0, -1))


def test_tie_coins_are_fair_and_keyed(mc):
    L = 256
    s = _libs.pattern_lattices(L)["stripes_i"]  # every 2x2 block ties
    with mc.Context(L, 2, seed=31337) as ctx:
        ctx.set_spins(np.stack([s, s]))
        ctx.measure(max_levels=1)
        a0, a1 = ctx.get_level_spins(0, 1), ctx.get_level_spins(1, 1)
        ctx.measure(max_levels=1)
        assert np.array_equal(a0, ctx.get_level_spins(0, 1))  # same key -> same coins
        assert (a0 != a1).mean() > 0.4  # other replica -> other coins
        ctx.sweep_counter = 5
        ctx.measure(max_levels=1)
        assert (a0 != ctx.get_level_spins(0, 1)).mean() > 0.4  # other time -> other coins
        n = a0.size
        assert abs((a0 == 1).sum() - n / 2) < 5 * np.sqrt(n) / 2


# ---------------------------------------------------------------------------------------------------------
# Metropolis trajectories   (bit-exact against the scalar specification with the same Philox keys)
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,strip,fuse,n_sweeps", [(2, 0, 1, 5), (2, 2, 2, 4), (4, 0, 1, 3), (8, 2, 1, 4), (16, 0, 2, 5), (32, 8, 1, 3), (64, 0, 1, 4),
                                                   (64, 16, 3, 7), (128, 0, 2, 3), (256, 32, 1, 2), (512, 0, 1, 2),
                                                   (1024, 16, 2, 2), (2048, 0, 1, 2), (2048, 4, 2, 2), (64, 6, 1, 3), (256, 10, 2, 3),
                                                   (1024, 36, 1, 2), (2048, 22, 1, 2)])
def test_sweeps_match_scalar_metropolis(mc, L, strip, fuse, n_sweeps):
    o = _libs.oracle()
    seed, base, t0 = 0xABCDEF0123, 3, (1 << 32) - 2  # crosses the 32-bit boundary of the sweep counter
    Ks = np.array([KC, -0.3, 0.35])
    R = len(Ks)
    with mc.Context(L, R, seed=seed, replica_base=base) as ctx:
        ctx.set_tuning(strip_rows=strip, fuse_sweeps=fuse)
        ctx.set_couplings(Ks)
        ctx.init_hot()
        ctx.sweep_counter = t0
        ctx.sweep(n_sweeps)
        assert ctx.sweep_counter == t0 + n_sweeps
        got = ctx.get_spins()
        # continuing in two calls gives the same chain as one call
        ctx.sweep(2)
        got2 = ctx.get_spins()
    for r in range(R):
        want = oracle_hot(L, seed, base + r)
        o.orc_metropolis(L, want, Ks[r], seed, base + r, t0, n_sweeps)
        assert np.array_equal(got[r], want), (L, r)
        o.orc_metropolis(L, want, Ks[r], seed, base + r, t0 + n_sweeps, 2)
        assert np.array_equal(got2[r], want), (L, r)


@pytest.mark.parametrize("L,strip", [(64, 0), (128, 0), (256, 16), (1024, 0)])
def test_sweeps_match_scalar_metropolis_for_every_threshold_pattern(mc, L, strip):
    """The first Philox call of a word is compared with code specialised on the leading bits of the acceptance thresholds
    (kernels.cu: mc_compare4_nz, chosen per replica): T4 = exp(-4|K|) in [3/16, 1/4), [1/8, 3/16), [1/16, 1/8), below 1/16,
    and the general path (T4 >= 1/4), both signs of K, edges included.  All must reproduce the scalar specification."""
    o = _libs.oracle()
    seed, base, t0, n_sweeps = 99, 11, 7, 3
    edge = 0.25 * np.log(4.0)  # T4 = 1/4 exactly (up to rounding): the boundary between the general and the special path
    Ks = np.array([-0.40, 0.36, -0.4406868, 0.50, -0.60, 0.69, -0.80, 1.5, -0.2, 0.05, edge, -np.nextafter(edge, 1.0), 0.25 * np.log(8.0),
                   -0.25 * np.log(16.0)])
    R = len(Ks)
    with mc.Context(L, R, seed=seed, replica_base=base) as ctx:
        ctx.set_tuning(strip_rows=strip)
        ctx.set_couplings(Ks)
        ctx.init_hot()
        ctx.sweep_counter = t0
        ctx.sweep(n_sweeps)
        got = ctx.get_spins()
    for r in range(R):
        want = oracle_hot(L, seed, base + r)
        o.orc_metropolis(L, want, Ks[r], seed, base + r, t0, n_sweeps)
        assert np.array_equal(got[r], want), (L, r, Ks[r])


def test_sweep_is_independent_of_strip_geometry_at_full_size(mc):
    """L = 4096 and 16384: the oracle is too slow for many sweeps, so use the size-independent property that the
    result cannot depend on the strip height or on how many sweeps are fused per launch (halo recomputation with
    counter-based random numbers), plus one exact oracle sweep at 4096."""
    o = _libs.oracle()
    for L, variants in ((4096, [(0, 1), (8, 1), (64, 2), (16, 3), (14, 1), (44, 1)]), (16384, [(0, 1), (8, 2), (32, 1), (20, 1)])):
        results = []
        for strip, fuse in variants:
            with mc.Context(L, 1, seed=99) as ctx:
                ctx.set_tuning(strip_rows=strip, fuse_sweeps=fuse)
                ctx.init_hot()
                ctx.sweep(3)
                obs = ctx.observables()
                results.append((obs, ctx.get_spins() if L == 4096 else None))
        for obs, spins in results[1:]:
            for k in ("Snn", "Snnn", "Splaq", "M"):
                assert obs[k][0] == results[0][0][k][0], (L, k)
            if spins is not None:
                assert np.array_equal(spins, results[0][1])
        if L == 4096:
            want = oracle_hot(L, 99, 0)
            o.orc_metropolis(L, want, KC, 99, 0, 0, 1)
            with mc.Context(L, 1, seed=99) as ctx:
                ctx.init_hot()
                ctx.sweep(1)
                assert np.array_equal(ctx.get_spins()[0], want)


def test_full_size_pyramid_properties(mc):
    """L = 16384 (config C5: 8 levels) and 4096 (11 levels): exact known answers for ordered configurations and
    the oracle on a random one at 4096."""
    for L in (4096, 16384):
        jj, ii = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
        S = []
        with mc.Context(L, 1, seed=5) as ctx:
            for name in ("up", "checker", "stripes_i"):
                if name == "up":
                    s = np.ones((L, L), np.int32)
                elif name == "checker":
                    s = (1 - 2 * ((ii + jj) & 1)).astype(np.int32)
                else:
                    s = (1 - 2 * (ii & 1)).astype(np.int32)
                ctx.set_spins(s[None])
                S.append(ctx.measure(max_levels=8)[0])
                del s
        del jj, ii
        S = np.stack(S)
        n2 = np.array([(L >> k) ** 2 for k in range(9)], np.int64)
        # all up: every level all up
        assert np.array_equal(S[0, :, 0], 4 * n2) and np.array_equal(S[0, :, 1], 4 * n2)
        assert np.array_equal(S[0, :, 2], n2) and np.array_equal(S[0, :, 3], n2)
        # checkerboard: level 0 is the Neel state; every block ties
        assert S[1, 0, 0] == -4 * n2[0] and S[1, 0, 1] == 4 * n2[0] and S[1, 0, 2] == n2[0] and S[1, 0, 3] == 0
        # stripes along i: nn = 0 (2 aligned + 2 anti), nnn = -4 n^2, plaquettes all +1
        assert S[2, 0, 0] == 0 and S[2, 0, 1] == -4 * n2[0] and S[2, 0, 2] == n2[0] and S[2, 0, 3] == 0
    L = 4096
    s = _libs.random_lattice(L, 8)
    with mc.Context(L, 1, seed=6, replica_base=2) as ctx:
        ctx.set_spins(s[None])
        S = ctx.measure()
    assert np.array_equal(S[0], _libs.pyramid(L, s, 6, 2, 0, -1))


# ---------------------------------------------------------------------------------------------------------
# the sample loop and its accumulators   (exact 128-bit integers)
# ---------------------------------------------------------------------------------------------------------

def cpu_run(L, seed, replica, K, t0, n_samples, m, max_levels, start, update="metropolis"):
    """mcrg_run's contract restated with the oracle: per sample measure (pyramid with Philox ties keyed by the
    sweep counter), accumulate (mcrg.cpp:86-97, exact), then m sweeps."""
    o = _libs.oracle()
    s = start.copy()
    n_lv = o.orc_n_transformations(L, 2)
    if 0 <= max_levels < n_lv:
        n_lv = max_levels
    S_sum = np.zeros((n_lv + 1) * 3, np.int64)
    SS = [[0] * 9 for _ in range(n_lv + 1)]
    SB0 = [[0] * 9 for _ in range(n_lv)]
    hi1 = np.zeros(n_lv * 9, np.int64); lo1 = np.zeros(n_lv * 9, np.uint64)
    hi2 = np.zeros(n_lv * 9, np.int64); lo2 = np.zeros(n_lv * 9, np.uint64)
    absM = M2 = 0
    M4 = 0
    t = t0
    for _ in range(n_samples):
        S4 = _libs.pyramid(L, s, seed, replica, t, max_levels)
        S3 = np.ascontiguousarray(S4[:, :3])
        o.orc_accumulate_i128(n_lv, 3, S3.ravel(), S_sum, hi1, lo1, hi2, lo2)
        for lv in range(n_lv + 1):
            for b in range(3):
                for a in range(3):
                    SS[lv][b * 3 + a] += int(S3[lv, a]) * int(S3[lv, b])
                    if lv >= 1:
                        SB0[lv - 1][b * 3 + a] += int(S3[lv, a]) * int(S3[0, b])
        M = int(S4[0, 3])
        absM += abs(M); M2 += M * M; M4 += M ** 4
        (o.orc_swendsen_wang if update == "cluster" else o.orc_metropolis)(L, s, K, seed, replica, t, m)
        t += m
    SbS = [int(h) * (1 << 64) + int(l) for h, l in zip(hi1, lo1)]
    SbSb = [int(h) * (1 << 64) + int(l) for h, l in zip(hi2, lo2)]
    return dict(n=n_samples, absM=absM, M2=M2, M4=M4, S=S_sum, SS=SS, SbS=SbS, SbSb=SbSb, SB0=SB0, final=s, n_lv=n_lv)


def assert_accumulators(mc, a, ad, want, tag=None):
    """Every live accumulator slot of one (replica, bin) against cpu_run's exact integers; unused levels stay zero."""
    lay = mc.capi.acc_layout()
    n_lv = want["n_lv"]
    assert a[lay.slot_n] == want["n"], tag
    assert a[lay.slot_absm] == want["absM"] and a[lay.slot_m2] == want["M2"], tag
    assert lay.m4(a) == want["M4"], tag  # exact: three 128-bit slots (include/mcrg_b200.h: slot_m4)
    assert abs(ad[lay.dslot_m4] - float(want["M4"])) <= 1e-12 * max(1.0, float(want["M4"])), tag
    for lv in range(n_lv + 1):
        for op in range(3):
            assert a[lay.slot_s + lv * 3 + op] == int(want["S"][lv * 3 + op]), (tag, lv, op)
        for e in range(9):
            assert a[lay.slot_ss + lv * 9 + e] == want["SS"][lv][e], (tag, lv, e)
    for n in range(n_lv):
        for e in range(9):
            assert a[lay.slot_sbs + n * 9 + e] == want["SbS"][n * 9 + e], (tag, n, e)
            assert a[lay.slot_ss + (n + 1) * 9 + e] == want["SbSb"][n * 9 + e], (tag, n, e)  # Sb_Sb == SS of level n+1
            assert a[lay.slot_sb0 + n * 9 + e] == want["SB0"][n][e], (tag, n, e)
    for lv in range(n_lv + 1, mc.capi.MAX_LEVELS + 1):  # slots of levels that do not exist stay zero
        assert all(a[lay.slot_s + lv * 3 + op] == 0 for op in range(3)), tag


def add_runs(a, b):
    """Totals of two consecutive cpu_run segments (the second started from the first one's final configuration)."""
    out = dict(n=a["n"] + b["n"], absM=a["absM"] + b["absM"], M2=a["M2"] + b["M2"], M4=a["M4"] + b["M4"], S=a["S"] + b["S"],
               final=b["final"], n_lv=a["n_lv"])
    out["SS"] = [[x + y for x, y in zip(u, v)] for u, v in zip(a["SS"], b["SS"])]
    out["SB0"] = [[x + y for x, y in zip(u, v)] for u, v in zip(a["SB0"], b["SB0"])]
    out["SbS"] = [x + y for x, y in zip(a["SbS"], b["SbS"])]
    out["SbSb"] = [x + y for x, y in zip(a["SbSb"], b["SbSb"])]
    return out


@pytest.mark.parametrize("L,n_samples,m,max_levels,graphs", [(8, 5, 1, -1, 0), (16, 20, 2, -1, 1), (64, 37, 1, -1, 1),
                                                            (64, 6, 3, 2, 0), (128, 18, 1, -1, 1), (512, 3, 1, -1, 0),
                                                            (1024, 2, 2, 4, 0)])
def test_run_accumulators_match_oracle(mc, L, n_samples, m, max_levels, graphs):
    seed, base, t0 = 2024, 1, 1000
    Ks = [KC, -0.42]
    lay = mc.capi.acc_layout()
    with mc.Context(L, 2, seed=seed, replica_base=base, n_bins=2) as ctx:
        ctx.set_tuning(use_graphs=graphs)
        ctx.set_couplings(Ks)
        ctx.init_hot()
        ctx.sweep_counter = t0
        ctx.run(n_samples, m, max_levels, bin=1)
        acc, accd = ctx.accumulators()
        final = ctx.get_spins()
        limbs_host = None
        try:
            import torch

            limbs = torch.zeros(lay.n_slots * 4, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            ctx.total_limbs_to_device(limbs.data_ptr())
            ctx.sync()
            limbs_host = limbs.cpu().numpy()
        except ImportError:
            pass
    assert (acc[:, 0, :] == 0).all()  # bin 0 untouched
    for r in range(2):
        want = cpu_run(L, seed, base + r, Ks[r], t0, n_samples, m, max_levels, oracle_hot(L, seed, base + r))
        assert_accumulators(mc, acc[r, 1], accd[r, 1], want, (L, r))
        assert np.array_equal(final[r], want["final"]), (L, r)
    if limbs_host is not None:
        tot = mc.dist.limbs_to_ints(limbs_host)
        assert tot == [int(x) for x in acc.sum(axis=(0, 1))]


def test_products_beyond_int64(mc):
    """L = 16384, all spins up: S_nn = 4 L^2 = 2^30, so S(1) x S(0) ~ 2^58 per sample; 64 samples overflow int64
    (2^63) — the 128-bit accumulators must hold the exact value."""
    L, n = 16384, 64
    lay = mc.capi.acc_layout()
    with mc.Context(L, 1, seed=1) as ctx:
        ctx.init_cold()
        ctx.run(n, 0, 1, 0)  # measure only: configuration stays all-up
        acc, _ = ctx.accumulators()
    a = acc[0, 0]
    s0, s1 = 4 * L * L, 4 * (L // 2) ** 2
    assert a[lay.slot_n] == n
    assert a[lay.slot_s + 0] == n * s0 and a[lay.slot_s + 3] == n * s1
    assert a[lay.slot_ss + 0] == n * s0 * s0 and n * s0 * s0 > 2**63
    assert a[lay.slot_sbs + 0] == n * s1 * s0
    assert a[lay.slot_m2] == n * (L * L) ** 2
    assert lay.m4(a) == n * (L * L) ** 4 and n * (L * L) ** 4 > 2**117  # sum M^4: exact in its three slots


# ---------------------------------------------------------------------------------------------------------
# RGNN forward pass and finite-difference gradient (rgnn.cpp:281-339)   (bit-exact against the oracle's order)
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("L", [2, 4, 8, 16, 64])
def test_rgnn_eval_matches_oracle(mc, L):
    o = _libs.oracle()
    rng = np.random.default_rng(L)
    R = 31  # not a multiple of the 14 replicas a block handles
    spins = np.stack([_libs.random_lattice(L, 500 + r, p_up=0.5 + 0.3 * (r % 2)) for r in range(R)])
    for W in (np.array([[0.5, -0.5], [0.5, -0.5]]), 0.3 * rng.standard_normal((2, 2))):  # train.cpp:19-23 and random
        with mc.Context(L, R) as ctx:
            ctx.set_spins(spins)
            ctx.rgnn_set_weights(W)
            u, g = ctx.rgnn_eval(1e-4)
        Wcm = np.ascontiguousarray(W.T).ravel()  # column-major for the oracle
        for r in range(R):
            want = o.orc_rgnn_scalar_output(L, spins[r], 2, Wcm)
            assert u[r] == want, (L, r, u[r], want)
            gw = np.zeros(4)
            o.orc_rgnn_gradient(L, spins[r], 2, Wcm.copy(), 1e-4, gw)
            assert np.array_equal(g[r], gw.reshape(2, 2).T), (L, r)


def test_rgnn_run_accumulates_like_the_reference_loop(mc):
    """rgnn.cpp:106-123: update, then u, u^2 and the gradient of the new configuration are added up."""
    o = _libs.oracle()
    L, R, n, m, h, seed = 8, 5, 7, 2, 1e-4, 321
    W = np.array([[0.5, -0.5], [0.5, -0.5]])
    Wcm = np.ascontiguousarray(W.T).ravel()
    with mc.Context(L, R, seed=seed) as ctx:
        ctx.set_couplings([KC])
        ctx.init_hot()
        ctx.rgnn_set_weights(W)
        ctx.rgnn_run(n, m, h)
        sums = ctx.rgnn_sums()
        ctx.rgnn_reset()
        assert (ctx.rgnn_sums() == 0).all()
    for r in range(R):
        s = oracle_hot(L, seed, r)
        want = np.zeros(6)
        for k in range(n):
            o.orc_metropolis(L, s, KC, seed, r, k * m, m)
            u = o.orc_rgnn_scalar_output(L, s, 2, Wcm)
            g = np.zeros(4)
            o.orc_rgnn_gradient(L, s, 2, Wcm.copy(), h, g)
            want[0] += u
            want[1] += u * u
            want[2:] += g
        assert np.array_equal(sums[r], want), (r, sums[r], want)


# ---------------------------------------------------------------------------------------------------------
# resume and sharding: results depend only on (seed, global replica id, sweep counter)
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,update", [(64, "metropolis"), (1024, "metropolis"), (128, "cluster")])
def test_checkpoint_resume_reproduces_the_trajectory(mc, L, update):
    """A checkpoint is (configurations in the reference layout, sweep counter): a fresh context restored from it
    continues bit-identically — the counter-based RNG has no other state (checkpoint/resume)."""
    seed, base = 99, 11
    with mc.Context(L, 2, seed=seed, replica_base=base) as ctx:
        ctx.set_update(update)
        ctx.set_couplings([KC, -0.45])
        ctx.init_hot()
        ctx.sweep(7)
        want = ctx.get_spins()
        S_want = ctx.measure()
    with mc.Context(L, 2, seed=seed, replica_base=base) as ctx:
        ctx.set_update(update)
        ctx.set_couplings([KC, -0.45])
        ctx.init_hot()
        ctx.sweep(3)
        ckpt, t = ctx.get_spins(), ctx.sweep_counter
    with mc.Context(L, 2, seed=seed, replica_base=base) as ctx:  # "restart"
        ctx.set_update(update)
        ctx.set_couplings([KC, -0.45])
        ctx.set_spins(ckpt)
        ctx.sweep_counter = t
        ctx.sweep(4)
        assert np.array_equal(ctx.get_spins(), want)
        assert np.array_equal(ctx.measure(), S_want)


def test_sharded_contexts_give_the_totals_of_one_context(mc):
    """The multi-GPU layout on one device: 6 replicas in one context == 2 + 4 replicas in two contexts with
    replica_base 0 and 2; accumulator totals (as the all-reduce would form them) are identical integers."""
    L, seed, n_samples = 128, 31, 9
    Ks = [KC, -0.43, -0.45, -0.47, -0.44, -0.42]

    def totals(first, count):
        with mc.Context(L, count, seed=seed, replica_base=first) as ctx:
            ctx.set_tuning(strip_rows=32)  # the strip kernel, as on the large lattices
            ctx.set_couplings(Ks[first:first + count])
            ctx.init_hot()
            ctx.sweep(5)
            ctx.run(n_samples, 2, -1, 0)
            acc, _ = ctx.accumulators()
            return acc[:, 0, :]

    whole = totals(0, 6)
    parts = np.concatenate([totals(0, 2), totals(2, 4)], axis=0)
    assert (whole == parts).all()
    assert [int(x) for x in whole.sum(axis=0)] == [int(x) for x in parts.sum(axis=0)]


def test_resident_runs_longer_than_one_launch(mc, monkeypatch):
    """k_resident keeps 64-bit sums per launch and the host layer splits runs (at 2^22 samples; lowered here through
    MCRG_RESIDENT_MAX_SAMPLES so that the split is cheap to reach): a run across several splits equals the same run
    issued in two calls (sweep counters and accumulators continue)."""
    L, n = 16, 2 * 700 + 5
    monkeypatch.setenv("MCRG_RESIDENT_MAX_SAMPLES", "700")
    with mc.Context(L, 2, seed=8) as a:
        a.set_couplings([KC, -0.3])
        a.init_hot()
        a.run(n, 2, -1, 0)
        acc_a, _ = a.accumulators()
        spins_a, t_a = a.get_spins(), a.sweep_counter
    monkeypatch.delenv("MCRG_RESIDENT_MAX_SAMPLES")
    with mc.Context(L, 2, seed=8) as b:
        b.set_couplings([KC, -0.3])
        b.init_hot()
        b.run(1000, 2, -1, 0)
        b.run(n - 1000, 2, -1, 0)
        acc_b, _ = b.accumulators()
        lay = mc.capi.acc_layout()
        assert (acc_a == acc_b).all() and acc_a[0, 0, lay.slot_n] == n
        assert np.array_equal(spins_a, b.get_spins()) and t_a == b.sweep_counter == 2 * n


@pytest.mark.parametrize("L,forced", [(64, 96), (64, 256), (128, 256), (32, 64), (8, 128)])
def test_resident_results_do_not_depend_on_the_block_size(mc, monkeypatch, L, forced):
    """k_resident sizes its CTA by the number of column walkers (one warp at L <= 64, its own register budget); the
    trajectory and every accumulator must equal those of a forced, larger block (MCRG_RESIDENT_THREADS, read per context)."""

    def run():
        with mc.Context(L, 3, seed=77, replica_base=5) as ctx:
            ctx.set_couplings([KC, -0.3, 0.2])
            ctx.init_hot()
            ctx.sweep(3)
            ctx.run(9, 2, -1, 0)
            acc, accd = ctx.accumulators()
            return acc, accd, ctx.get_spins(), ctx.measure()

    a = run()
    monkeypatch.setenv("MCRG_RESIDENT_THREADS", str(forced))
    b = run()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
