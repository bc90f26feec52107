"""Statistical parity (north_star: 3 sigma on combined jackknife errors) of the GPU checkerboard Metropolis +
MCRG pipeline against the REFERENCE's own sampler and driver.

The reference updates with Wolff clusters (ising.cpp:87-155), this repo with Metropolis sweeps: trajectories cannot
be compared, equilibrium averages can.  Reference numbers: tests/golden/statistical.json — means and standard errors
over 16 independently seeded runs of the compiled reference (tests/golden/make_golden.py).  GPU numbers: many
independent replicas; errors by jackknife over replica groups (replicas are independent chains, so the jackknife is
valid whatever the autocorrelation time inside a chain).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KC = -0.5 * np.log(1 + np.sqrt(2))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "statistical.json")
N_SIGMA = 3.0


@pytest.fixture(scope="module")
def mc():
    import mcrg_b200

    assert mcrg_b200.capi.device_count() >= 1
    return mcrg_b200


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def run_chains(mc, L, K, n_replicas, n_eq, n_samples, m, seed, max_levels=-1):
    with mc.Context(L, n_replicas, seed=seed) as ctx:
        ctx.set_couplings([K])
        ctx.init_hot()
        ctx.sweep(n_eq)
        ctx.run(n_samples, m, max_levels, 0)
        acc, accd = ctx.accumulators()
    return acc[:, 0, :], accd[:, 0, :]


def grouped(acc, accd, n_groups):
    """Sum replicas into n_groups jackknife chunks; the double slot is appended as one more (float) column."""
    R = acc.shape[0]
    per = R // n_groups
    chunks = []
    for g in range(n_groups):
        a = acc[g * per:(g + 1) * per].sum(axis=0)
        d = accd[g * per:(g + 1) * per].sum(axis=0)
        chunks.append(list(a) + [float(d[0])])
    return chunks


@pytest.mark.parametrize("L,K,m", [(8, KC, 4), (16, KC, 8), (32, KC, 16), (16, -0.40, 4), (16, -0.48, 8),
                                   (8, -0.4320459, 4), (8, -0.4496804, 4), (8, -0.4688157, 4), (8, -0.489652, 4)])
def test_thermodynamic_averages_match_reference_sampler(mc, gold, L, K, m):
    ref = next(t for t in gold["thermo"] if t["N"] == L and abs(t["K"] - K) < 1e-9)
    lay = mc.capi.acc_layout()
    acc, accd = run_chains(mc, L, K, n_replicas=1024, n_eq=3000, n_samples=1500, m=m, seed=1000 + L, max_levels=0)
    N = L * L

    def obs(v):
        n = float(v[lay.slot_n])
        bond = float(v[lay.slot_s + 0]) / n / (4.0 * N)
        absm = float(v[lay.slot_absm]) / n / N
        m2 = float(v[lay.slot_m2]) / n / N**2
        m4 = float(v[-1]) / n / float(N) ** 4
        return np.array([bond, absm, m2, m4, 1.0 - m4 / (3.0 * m2 * m2)])

    est, err = mc.analysis.jackknife(grouped(acc, accd, 32), obs)
    for k, name in enumerate(["bond", "absm", "m2", "m4", "U4"]):
        want, want_err = ref[name]
        sigma = np.hypot(err[k], want_err)
        assert abs(est[k] - want) < N_SIGMA * sigma, (L, K, name, est[k], err[k], want, want_err, (est[k] - want) / sigma)


@pytest.mark.parametrize("L,m,n_samples", [(16, 8, 3000), (32, 16, 3000), (64, 32, 2500)])
def test_rg_eigenvalues_match_reference_driver(mc, gold, L, m, n_samples):
    """lambda per blocking level from the 2-operator (NN, NNN) matrix, as calc_critical_exponent prints it
    (mcrg.cpp:111-131), at K_c: reference mean +- error over 16 seeded runs vs GPU jackknife."""
    ref = next(t for t in gold["lambda"] if t["N"] == L)
    n_lv = int(np.log2(L)) - 1
    acc, accd = run_chains(mc, L, KC, n_replicas=1024, n_eq=4000, n_samples=n_samples, m=m, seed=77 + L)

    def lambdas(v):
        return mc.analysis.rg_eigenvalues(mc.analysis.unpack_slots(v[:-1], n_lv), ops=(0, 1))[0]

    est, err = mc.analysis.jackknife(grouped(acc, accd, 32), lambdas)
    assert len(est) == len(ref["mean"]) == n_lv
    for lv in range(n_lv):
        sigma = np.hypot(err[lv], ref["err"][lv])
        assert abs(est[lv] - ref["mean"][lv]) < N_SIGMA * sigma, (L, lv, est[lv], err[lv], ref["mean"][lv], ref["err"][lv])
    # the physics: the thermal eigenvalue approaches 2 (nu = 1) on the intermediate levels
    assert abs(est[1] - 2.0) < 0.03


EXACT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_ising.json")


@pytest.mark.parametrize("update,m", [("metropolis", 4), ("cluster", 1)])
@pytest.mark.parametrize("L,absK", [(4, 0.4406867935097715), (4, 0.40), (8, 0.4406867935097715), (8, 0.489652),
                                    (16, 0.4406867935097715), (16, 0.48), (32, 0.4406867935097715), (64, 0.4406867935097715)])
def test_energy_matches_exact_finite_size_solution(mc, update, m, L, absK):
    """<s s'> on nearest-neighbour bonds against Kaufman's exact torus partition function (tests/golden/make_exact.py):
    a known answer with no statistical error of its own, for both samplers, ferro- and antiferromagnetic sign."""
    with open(EXACT) as f:
        ex = json.load(f)
    want = next(t for t in ex["kaufman_bond"] if t["L"] == L and abs(t["absK"] - absK) < 1e-12)["bond"]
    lay = mc.capi.acc_layout()
    for sign in (-1.0, +1.0):
        n_rep = 4096 if L <= 16 else 1024
        with mc.Context(L, n_rep, seed=600 + L + (7 if sign > 0 else 0)) as ctx:
            ctx.set_update(update)
            ctx.set_couplings([sign * absK])
            ctx.init_hot()
            ctx.sweep(100 * m * (L // 4) if update == "metropolis" else 100)
            ctx.run(1000, m * max(1, L // 8) if update == "metropolis" else 1, 0, 0)
            acc, accd = ctx.accumulators()
        per = np.array([float(acc[r, 0, lay.slot_s]) / float(acc[r, 0, lay.slot_n]) / (4.0 * L * L) for r in range(n_rep)])
        mean, err = per.mean(), per.std(ddof=1) / np.sqrt(n_rep)
        assert abs(mean - (-sign) * want) < N_SIGMA * err, (update, L, sign * absK, mean, err, want, (mean + sign * want) / err)
        if L == 4 and sign < 0 and any(abs(t["absK"] - absK) < 1e-12 for t in ex["enum_4x4"]):
            e4 = next(t for t in ex["enum_4x4"] if abs(t["absK"] - absK) < 1e-12)
            absm = np.array([float(acc[r, 0, lay.slot_absm]) / float(acc[r, 0, lay.slot_n]) / 16.0 for r in range(n_rep)])
            m2 = np.array([float(acc[r, 0, lay.slot_m2]) / float(acc[r, 0, lay.slot_n]) / 256.0 for r in range(n_rep)])
            m4 = accd[:, 0, 0] / np.array([float(acc[r, 0, lay.slot_n]) for r in range(n_rep)]) / 16.0**4
            for name, v in (("absm", absm), ("m2", m2), ("m4", m4)):
                assert abs(v.mean() - e4[name]) < N_SIGMA * v.std(ddof=1) / np.sqrt(n_rep), (update, name, v.mean(), e4[name])


def test_three_operator_matrix_and_antiferromagnet(mc):
    """Extensions without a reference counterpart, pinned by physics: (i) adding the plaquette operator keeps
    lambda_t near 2; (ii) K > 0 at |K| = K_c is the same model on the bipartite lattice (staggered gauge), so the
    bond alignment flips sign and has the same magnitude."""
    L = 32
    n_lv = 4
    acc, accd = run_chains(mc, L, KC, 512, 3000, 2000, 16, seed=5)
    lam3 = mc.analysis.rg_eigenvalues(mc.analysis.unpack_slots(list(acc.sum(axis=0)), n_lv), ops=(0, 1, 2))[0]
    assert abs(lam3[1] - 2.0) < 0.06 and abs(lam3[2] - 2.0) < 0.06, lam3
    lay = mc.capi.acc_layout()
    accF, _ = run_chains(mc, 16, KC, 256, 2000, 500, 8, seed=6, max_levels=0)
    accA, _ = run_chains(mc, 16, -KC, 256, 2000, 500, 8, seed=7, max_levels=0)
    bF = float(accF[:, lay.slot_s].sum()) / float(accF[:, lay.slot_n].sum())
    bA = float(accA[:, lay.slot_s].sum()) / float(accA[:, lay.slot_n].sum())
    assert bF > 0 and bA < 0 and abs(bF + bA) < 0.01 * abs(bF)
