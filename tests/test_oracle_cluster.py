"""The scalar specification of the Swendsen-Wang cluster update (oracle/mcrg_oracle.c: orc_swendsen_wang) samples the
Boltzmann distribution of the reference's model: weight exp(-K sum_<ij> s_i s_j) with the bond sum of
lattice.cpp:84-99 (every site's four neighbours, i.e. each bond twice — on the 2x2 torus each PAIR four times) and
the add probability 1 - exp(2K) of ising.cpp:9.  Checked against exact enumeration (L = 2, 4) and against the
reference's own Wolff sampler (tests/golden/statistical.json, L = 8)."""
import itertools
import json
import os

import numpy as np

import _libs

KC = -0.5 * np.log(1 + np.sqrt(2))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def exact_averages(L, K):
    """<S_nn/(4 L^2)>, <|M|>/L^2 by enumeration; S_nn is the reference's double-counted sum (lattice.cpp:84-99)."""
    n = L * L
    assert n <= 16
    states = np.array(list(itertools.product([-1, 1], repeat=n)), np.int8).reshape(-1, L, L)
    snn = np.zeros(len(states), np.int64)
    for ax in (1, 2):
        for sh in (1, -1):
            snn += (states.astype(np.int64) * np.roll(states, sh, axis=ax)).sum(axis=(1, 2))
    # S_nn counts every bond twice, so the weight exp(-K sum_bonds s s') — the one for which 1 - exp(2K) is the correct
    # cluster bond probability (ising.cpp:9) — is exp(-(K/2) S_nn)
    w = np.exp(-0.5 * K * snn.astype(np.float64))
    Z = w.sum()
    M = states.astype(np.int64).sum(axis=(1, 2))
    return float((w * snn).sum() / Z) / (4.0 * n), float((w * np.abs(M)).sum() / Z) / n


def chain_averages(update, L, K, n_chains, n_eq, n_samples, stride, seed):
    o = _libs.oracle()
    bond, absm = [], []
    for c in range(n_chains):
        s = np.zeros((L, L), np.int32)
        o.orc_hot_start(L, seed, c, s)
        update(L, s, K, seed, c, 0, n_eq)
        b = a = 0.0
        for k in range(n_samples):
            update(L, s, K, seed, c, n_eq + stride * k, stride)
            b += o.orc_calc_nn(L, s) / (4.0 * L * L)
            a += abs(int(s.sum())) / (L * L)
        bond.append(b / n_samples)
        absm.append(a / n_samples)
    return (np.mean(bond), np.std(bond, ddof=1) / np.sqrt(n_chains)), (np.mean(absm), np.std(absm, ddof=1) / np.sqrt(n_chains))


def test_cluster_update_samples_the_exact_distribution():
    o = _libs.oracle()
    for L, K in ((2, -0.3), (4, KC), (4, -0.25), (4, +0.35)):
        want_bond, want_absm = exact_averages(L, K)
        (b, be), (a, ae) = chain_averages(o.orc_swendsen_wang, L, K, n_chains=8, n_eq=50, n_samples=5000, stride=1, seed=5 + L)
        assert abs(b - want_bond) < 4 * be + 1e-12, (L, K, b, be, want_bond)
        assert abs(a - want_absm) < 4 * ae + 1e-12, (L, K, a, ae, want_absm)


def test_metropolis_samples_the_exact_distribution():
    """Same exact check for the Metropolis specification (pins the acceptance thresholds exp(-4|K|), exp(-8|K|))."""
    o = _libs.oracle()
    for L, K in ((4, KC), (4, +0.35)):
        want_bond, want_absm = exact_averages(L, K)
        (b, be), (a, ae) = chain_averages(o.orc_metropolis, L, K, n_chains=8, n_eq=100, n_samples=6000, stride=2, seed=9)
        assert abs(b - want_bond) < 4 * be, (L, K, b, be, want_bond)
        assert abs(a - want_absm) < 4 * ae, (L, K, a, ae, want_absm)


def test_cluster_update_matches_reference_wolff_sampler():
    o = _libs.oracle()
    with open(os.path.join(GOLD, "statistical.json")) as f:
        ref = next(t for t in json.load(f)["thermo"] if t["N"] == 8 and abs(t["K"] - KC) < 1e-9)
    (b, be), (a, ae) = chain_averages(o.orc_swendsen_wang, 8, KC, n_chains=8, n_eq=100, n_samples=4000, stride=1, seed=21)
    assert abs(b - ref["bond"][0]) < 4 * np.hypot(be, ref["bond"][1]), (b, be, ref["bond"])
    assert abs(a - ref["absm"][0]) < 4 * np.hypot(ae, ref["absm"][1]), (a, ae, ref["absm"])


def test_cluster_update_is_keyed_and_flips_whole_clusters():
    """Determinism in (seed, replica, t); cold configuration at strong coupling: one cluster, all or nothing."""
    o = _libs.oracle()
    L = 16
    a = _libs.random_lattice(L, 3)
    b = a.copy()
    o.orc_swendsen_wang(L, a, KC, 7, 2, 10, 3)
    o.orc_swendsen_wang(L, b, KC, 7, 2, 10, 1)
    o.orc_swendsen_wang(L, b, KC, 7, 2, 11, 2)
    assert np.array_equal(a, b)
    c = _libs.random_lattice(L, 3)
    o.orc_swendsen_wang(L, c, KC, 7, 3, 10, 3)
    assert not np.array_equal(a, c)
    seen = set()
    for t in range(40):
        s = np.ones((L, L), np.int32)
        o.orc_swendsen_wang(L, s, -20.0, 1, 0, t, 1)
        assert abs(int(s.sum())) == L * L
        seen.add(int(s[0, 0]))
    assert seen == {1, -1}
