"""Pins the plain-C oracle to the UNMODIFIED reference (oracle/_ref/libmcrg_ref.so, built from /root/reference/src).

Everything here is CPU-only.  Skipped when the compiled reference is absent; the committed fixtures in
tests/golden/ (made from the same library) carry the same checks to machines without /root/reference.
"""
import ctypes as C
import tempfile

import numpy as np
import pytest

import _libs

pytestmark = pytest.mark.ref

KC = -0.5 * np.log(1 + np.sqrt(2))
SIZES = [2, 4, 6, 8, 16, 32, 64]


def all_cases(N):
    cases = dict(_libs.pattern_lattices(N))
    for seed in range(3):
        cases[f"rand{seed}"] = _libs.random_lattice(N, 100 + seed)
    cases["biased"] = _libs.random_lattice(N, 7, p_up=0.8)
    if N >= 4 and (N & (N - 1)) == 0:
        cases["clustered"] = _libs.clustered_lattice(N, 5)
    return cases


@pytest.mark.parametrize("N", SIZES)
def test_interactions_energy_magnetisation(N):
    o, r = _libs.oracle(), _libs.ref()
    for name, s in all_cases(N).items():
        want = np.zeros(2)
        r.ref_calc_interactions(N, s, want)
        got = np.zeros(2, np.int64)
        o.orc_calc_interactions(N, s, got)
        assert got[0] == want[0] and got[1] == want[1], (N, name)
        assert o.orc_calc_nn(N, s) == r.ref_calc_nn(N, s)
        for K in (KC, -0.3, 0.2):
            # same accumulation order => identical doubles, not just close
            assert o.orc_calc_energy(N, s, K) == r.ref_calc_energy(N, s, K), (N, name, K)
        assert o.orc_calc_magnetization(N, s) == r.ref_calc_magnetization(N, s, KC)


def test_neighbour_tables_match_identity():
    """lattice.cpp:124-151 tables vs the oracle's wrap arithmetic, including the degenerate N=2 lattice."""
    r = _libs.ref()
    for N in (2, 3, 4, 8):
        for i in range(N):
            for j in range(N):
                nn = np.zeros(8, np.int32)
                nnn = np.zeros(8, np.int32)
                r.ref_neighbors(N, i, j, nn, nnn)
                nn = nn.reshape(2, 4).T  # column-major 4x2
                nnn = nnn.reshape(2, 4).T
                ip, im, jp, jm = (i + 1) % N, (i - 1) % N, (j + 1) % N, (j - 1) % N
                assert nn.tolist() == [[ip, j], [im, j], [i, jp], [i, jm]]
                assert nnn.tolist() == [[ip, jp], [im, jp], [ip, jm], [im, jm]]


@pytest.mark.parametrize("N", [2, 4, 8, 16, 64])
def test_block_spin_teacher_forced(N):
    """mcrg.cpp:314-348.  The reference's ties come from its global mt19937_64; the oracle is handed the
    reference's choices at tie blocks and must then agree everywhere; the tie mask itself must match too."""
    o, r = _libs.oracle(), _libs.ref()
    for name, s in all_cases(N).items():
        Nb = N // 2
        want = np.zeros((Nb, Nb), np.int32)
        r.ref_seed(1234)
        a = r.ref_block_spin(N, 2, 1, s, want)
        assert a == 2
        got = np.zeros((Nb, Nb), np.int32)
        mask = np.zeros((Nb, Nb), np.int32)
        o.orc_block_spin_supplied(N, 2, s, want, got, mask)
        assert np.array_equal(got, want), (N, name)
        # tie mask is exactly "block sum == 0", and non-tie outputs do not depend on the coins
        sums = s.reshape(Nb, 2, Nb, 2).sum(axis=(1, 3))
        assert np.array_equal(mask != 0, sums == 0)
        got2 = np.zeros((Nb, Nb), np.int32)
        o.orc_block_spin_supplied(N, 2, s, -want, got2, mask)
        assert np.array_equal(got2[sums != 0], want[sums != 0])
        assert np.array_equal(got2[sums == 0], -want[sums == 0])


def test_block_spin_b3_truncates_like_reference():
    """N not divisible by b silently truncates (mcrg.cpp:316)."""
    o, r = _libs.oracle(), _libs.ref()
    s = _libs.random_lattice(8, 3)
    want = np.zeros((2, 2), np.int32)
    r.ref_seed(9)
    r.ref_block_spin(8, 3, 1, s, want)
    got = np.zeros((2, 2), np.int32)
    o.orc_block_spin_supplied(8, 3, s, want, got, np.zeros((2, 2), np.int32))
    assert np.array_equal(got, want)


def test_reference_tie_coin_is_fair():
    r = _libs.ref()
    s = _libs.pattern_lattices(64)["stripes_i"]  # every block ties
    out = np.zeros((32, 32), np.int32)
    r.ref_seed(77)
    ups = 0
    for _ in range(20):
        r.ref_block_spin(64, 2, 1, s, out)
        ups += int((out == 1).sum())
    n = 20 * 1024
    assert abs(ups - n / 2) < 5 * np.sqrt(n) / 2


def test_helpers_split_flatten_levels():
    o, r = _libs.oracle(), _libs.ref()
    for n_samples in (1, 10, 1000, 12345):
        for P in (1, 2, 7, 100):
            for rank in (0, 1, P - 1):
                assert o.orc_split_samples(rank, P, n_samples) == r.ref_split_samples(rank, P, n_samples)
    m = np.array([1.0, 2.0, 3.0, 4.0])  # column-major 2x2
    out = np.zeros(4)
    r.ref_flatten2(m, out)
    assert out.tolist() == [1.0, 2.0, 3.0, 4.0]  # flatten is the identity on column-major storage
    for N in (4, 8, 16, 32, 64, 128, 1024, 4096, 16384):
        assert o.orc_n_transformations(N, 2) == int(np.log2(N)) - 1


@pytest.mark.parametrize("N,n_samples", [(8, 3000), (16, 2000), (32, 500)])
def test_accumulation_and_rg_matrix_against_real_driver(N, n_samples):
    """The real calc_critical_exponent (mcrg.cpp:22-144) and the logged loop visit the same configurations
    for the same seed; the oracle's accumulation + RG algebra applied to the logged S must give the lambdas
    that the reference wrote to its output file (printed with 10 decimals)."""
    o, r = _libs.oracle(), _libs.ref()
    n_eq = 200
    lam_ref = np.zeros(16)
    nu_ref = np.zeros(16)
    with tempfile.TemporaryDirectory() as d:
        r.ref_seed(4242)
        n_lv = r.ref_critical_exponent(d.encode(), n_eq, n_samples, N, KC, lam_ref, nu_ref, 16)
    assert n_lv == int(np.log2(N)) - 1
    S_log = np.zeros((n_samples, n_lv + 1, 2))
    lv = C.c_int(0)
    r.ref_seed(4242)
    r.ref_mcrg_loop(n_eq, n_samples, N, KC, 0, S_log.ctypes.data, C.byref(lv))
    assert lv.value == n_lv
    S_sum = np.zeros((n_lv + 1) * 2)
    SbS = np.zeros(n_lv * 4)
    SbSb = np.zeros(n_lv * 4)
    for s in range(n_samples):
        o.orc_accumulate(n_lv, 2, np.ascontiguousarray(S_log[s]).ravel(), S_sum, SbS, SbSb)
    lam = np.zeros(n_lv)
    nu = np.zeros(n_lv)
    o.orc_rg_eigenvalues(n_lv, 2, float(n_samples), 2, S_sum, SbS, SbSb, lam, nu)
    assert np.allclose(lam, lam_ref[:n_lv], rtol=0, atol=2e-9), (lam, lam_ref[:n_lv])
    assert np.allclose(nu, nu_ref[:n_lv], rtol=0, atol=2e-8 * np.abs(nu_ref[:n_lv]).max() + 2e-9)
    # exact integer accumulation agrees with the double one while everything is < 2^53
    Si = np.rint(S_log).astype(np.int64)
    S_sum_i = np.zeros((n_lv + 1) * 2, np.int64)
    hi1 = np.zeros(n_lv * 4, np.int64); lo1 = np.zeros(n_lv * 4, np.uint64)
    hi2 = np.zeros(n_lv * 4, np.int64); lo2 = np.zeros(n_lv * 4, np.uint64)
    for s in range(n_samples):
        o.orc_accumulate_i128(n_lv, 2, np.ascontiguousarray(Si[s]).ravel(), S_sum_i, hi1, lo1, hi2, lo2)
    assert np.array_equal(S_sum_i.astype(np.float64), S_sum)
    assert np.array_equal(hi1 * 2.0**64 + lo1.astype(np.float64), SbS)
    assert np.array_equal(hi2 * 2.0**64 + lo2.astype(np.float64), SbSb)


def test_rgnn_forward_and_gradient():
    """rgnn.cpp:281-339 with the train.cpp:19-23 starting weights and random ones."""
    o, r = _libs.oracle(), _libs.ref()
    rng = np.random.default_rng(0)
    W0 = np.array([0.5, 0.5, -0.5, -0.5])  # column-major [[.5,-.5],[.5,-.5]]
    for N in (2, 4, 8, 16):
        for seed in range(4):
            s = _libs.random_lattice(N, seed)
            for W in (W0, 0.3 * rng.standard_normal(4)):
                W = np.ascontiguousarray(W)
                a = o.orc_rgnn_scalar_output(N, s, 2, W)
                b = r.ref_rgnn_scalar_output(N, s, 2, W)
                assert abs(a - b) <= 1e-12 * max(1.0, abs(b))
                ga = np.zeros(4)
                gb = np.zeros(4)
                o.orc_rgnn_gradient(N, s, 2, W.copy(), 1e-4, ga)
                r.ref_rgnn_gradient(N, s, 2, W, 1e-4, gb)
                assert np.allclose(ga, gb, rtol=0, atol=1e-7 * max(1.0, np.abs(gb).max()))


def test_write_iter_restatement_matches_reference():
    """definitions.cpp:44-68 (the log-spaced output schedule of equilibrate(write=true)) against tests/_libs.ref_write_iter_py."""
    r = _libs.ref()
    for i in list(range(1, 3000)) + list(range(9990, 10011)) + [99999, 100000, 100001, 150000, 200000, 1000000, 1234567]:
        assert bool(r.ref_write_iter(i)) == _libs.ref_write_iter_py(i), i
