"""CPU check of the device functions in mcrg_b200/csrc/{bitops,tile}.cuh.

tests/cpu_emul/emul.cpp compiles those headers with g++ and walks them with the kernels' strip/halo/phase
structure; the results must be bit-identical to the oracle's scalar code.  This is test scaffolding: the product
path is the CUDA library, exercised by the -m gpu tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _libs

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_emul", "emul.cpp")
OUT = os.path.join(HERE, "cpu_emul", "_build", "libmcrg_emul.so")
KC = -0.5 * np.log(1 + np.sqrt(2))


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(_libs.ROOT, "mcrg_b200", "csrc", f) for f in ("bitops.cuh", "tile.cuh", "mcfast.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(OUT) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT], check=True)
    e = C.CDLL(OUT)
    e.emul_sweep.argtypes = [C.c_int, _libs.i32p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                             C.c_uint32, C.c_uint32, C.c_uint64]
    e.emul_measure.argtypes = [C.c_int, _libs.i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _libs.i64p, C.c_void_p]
    e.emul_measure.restype = C.c_int
    e.emul_hot_start.argtypes = [C.c_int, C.c_uint64, C.c_uint32, _libs.i32p]
    e.emul_fast_paths.argtypes = [C.c_int, C.c_uint64]
    e.emul_fast_paths.restype = C.c_int
    e.emul_requeue_pass2.argtypes = [C.c_int, C.c_uint64]
    e.emul_requeue_pass2.restype = C.c_int
    return e


def thresholds(K):
    o = _libs.oracle()
    t4, t8 = C.c_uint32(), C.c_uint32()
    o.orc_thresholds(K, C.byref(t4), C.byref(t8))
    return t4.value, t8.value, (0xFFFFFFFF if K > 0 else 0)


@pytest.mark.parametrize("L", [2, 4, 8, 16, 32, 64, 128, 256])
def test_hot_start(emul, L):
    o = _libs.oracle()
    a = np.zeros((L, L), np.int32)
    b = np.zeros((L, L), np.int32)
    emul.emul_hot_start(L, 12345, 3, a)
    o.orc_hot_start(L, 12345, 3, b)
    assert np.array_equal(a, b)
    assert abs(a.mean()) < 5.0 / L


@pytest.mark.parametrize("L,R,fuse,n_sweeps,K", [
    (2, 2, 1, 4, KC), (2, 2, 2, 3, -0.2), (4, 4, 1, 3, KC), (4, 2, 1, 2, KC), (8, 8, 1, 3, KC), (8, 2, 2, 4, -0.3), (16, 4, 1, 2, KC), (16, 16, 3, 5, KC),
    (32, 8, 2, 3, -0.6), (64, 64, 1, 2, KC), (64, 16, 2, 3, 0.35), (128, 32, 1, 2, KC), (128, 8, 4, 4, -0.44),
    (256, 64, 2, 2, KC), (512, 32, 1, 1, KC),
])
def test_sweeps_match_scalar_metropolis(emul, L, R, fuse, n_sweeps, K):
    o = _libs.oracle()
    t4, t8, anti = thresholds(K)
    seed, replica, t0 = 0xDEADBEEF12345, 7, (1 << 33) + 5
    a = np.zeros((L, L), np.int32)
    o.orc_hot_start(L, seed, replica, a)
    b = a.copy()
    emul.emul_sweep(L, a, R, fuse, n_sweeps, seed, replica, t4, t8, anti, t0)
    o.orc_metropolis(L, b, K, seed, replica, t0, n_sweeps)
    assert np.array_equal(a, b)
    # and the chain actually moved
    c = np.zeros((L, L), np.int32)
    o.orc_hot_start(L, seed, replica, c)
    assert (a != c).mean() > 0.05


def test_sweep_from_ordered_start_low_temperature(emul):
    """A cold lattice deep in the ordered phase: almost every site has A == 0, exercising the rare-acceptance
    branch and the long lazy comparisons."""
    o = _libs.oracle()
    L, K = 64, -0.8
    t4, t8, anti = thresholds(K)
    a = np.ones((L, L), np.int32)
    b = a.copy()
    emul.emul_sweep(L, a, 16, 1, 6, 99, 0, t4, t8, anti, 0)
    o.orc_metropolis(L, b, K, 99, 0, 0, 6)
    assert np.array_equal(a, b)
    assert 0 < (a == -1).sum() < 200


@pytest.mark.parametrize("L,R,Rn", [(2, 2, 2), (4, 4, 2), (8, 2, 2), (16, 16, 4), (32, 8, 8), (64, 64, 16), (128, 16, 16),
                                    (256, 32, 32), (512, 64, 64), (1024, 64, 32), (2048, 32, 64)])
def test_measurement_pyramid(emul, L, R, Rn):
    seed, replica, t = 4242, 11, 123456789012
    cases = {"clustered": _libs.clustered_lattice(L, 3, n_sweeps=8 if L <= 256 else 2) if L <= 512 else None,
             "random": _libs.random_lattice(L, 1), "checker": _libs.pattern_lattices(L)["checker"],
             "stripes": _libs.pattern_lattices(L)["stripes_i"], "up": _libs.pattern_lattices(L)["up"]}
    for name, s in cases.items():
        if s is None:
            continue
        want_S, want_lv = _libs.pyramid(L, s, seed, replica, t, -1, want_levels=True)
        n_lv = want_S.shape[0] - 1
        S = np.zeros((n_lv + 1, 4), np.int64)
        lev = np.zeros(sum((L >> k) ** 2 for k in range(1, n_lv + 1)) or 1, np.int32)
        got = emul.emul_measure(L, s.copy(), R, Rn, 0, -1, seed, replica, 0, 0, 0, t, S, lev.ctypes.data)
        assert got == n_lv
        assert np.array_equal(S, want_S), (L, name, S, want_S)
        off = 0
        for k in range(1, n_lv + 1):
            n = L >> k
            assert np.array_equal(lev[off:off + n * n].reshape(n, n), want_lv[k - 1]), (L, name, k)
            off += n * n


def test_measure_with_fused_sweep_and_level_cap(emul):
    """k_sweep0<true> measures the incoming configuration and then sweeps it; max_levels caps the pyramid."""
    o = _libs.oracle()
    L, K = 64, KC
    t4, t8, anti = thresholds(K)
    s = _libs.clustered_lattice(L, 9)
    want_S = _libs.pyramid(L, s, 5, 2, 77, 3)
    want_next = s.copy()
    o.orc_metropolis(L, want_next, K, 5, 2, 77, 1)
    S = np.zeros((4, 4), np.int64)
    got_next = s.copy()
    n = emul.emul_measure(L, got_next, 16, 8, 1, 3, 5, 2, t4, t8, anti, 77, S, None)
    assert n == 3
    assert np.array_equal(S, want_S)
    assert np.array_equal(got_next, want_next)


def test_fast_paths_equal_the_specification(emul):
    """mcrg_b200/csrc/mcfast.cuh — what the sweep kernels actually execute per word: Philox with the word-independent part of
    rounds 0-1 shared between calls, the full-adder neighbour count, the first-call compare specialised on the leading
    threshold bits — against philox4x32_10 / metropolis_flip_mask of bitops.cuh (which tests above tie to the oracle):
    300 000 random words, keys, sweep counters (incl. the 32-bit boundary) and thresholds of every pattern, no disagreement."""
    assert emul.emul_fast_paths(300000, 2026) == 0


def test_requeue_schedule_of_pass_two(emul):
    """Pass 2 of a strip half-sweep (kernels.cu: mc_half_sweep_t<.., REQUEUE>): every batch of queued words but the last runs one
    Philox call and re-queues what is still undecided, the last batch finishes completely, a full segment finishes on the spot —
    the same flips per word as finishing every entry on its own, for queue lengths 0 .. 199, capacities 8 .. 227 and thresholds
    that keep lanes undecided over several calls (2 000 random queues)."""
    assert emul.emul_requeue_pass2(2000, 77) == 0
