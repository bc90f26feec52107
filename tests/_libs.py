"""ctypes access to the two CPU checkers (oracle/_build/libmcrg_oracle.so, oracle/_ref/libmcrg_ref.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libmcrg_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmcrg_ref.so")

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_oracle():
    """(Re)build the plain-C oracle; cheap (under a second)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)


def ref_available():
    return os.path.exists(REF_SO)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        src = os.path.join(ROOT, "oracle", "mcrg_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            build_oracle()
        o = C.CDLL(ORACLE_SO)
        o.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        o.orc_philox_keyed.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_int, u32p]
        o.orc_thresholds.argtypes = [C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        o.orc_calc_interactions.argtypes = [C.c_int, i32p, i64p]
        o.orc_calc_nn.argtypes = [C.c_int, i32p]
        o.orc_calc_nn.restype = C.c_int64
        o.orc_calc_energy.argtypes = [C.c_int, i32p, C.c_double]
        o.orc_calc_energy.restype = C.c_double
        o.orc_calc_magnetization.argtypes = [C.c_int, i32p]
        o.orc_calc_magnetization.restype = C.c_double
        o.orc_sum_spins.argtypes = [C.c_int, i32p]
        o.orc_sum_spins.restype = C.c_int64
        o.orc_plaquette.argtypes = [C.c_int, i32p]
        o.orc_plaquette.restype = C.c_int64
        o.orc_block_spin_supplied.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, i32p]
        o.orc_tie_spin.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
        o.orc_tie_spin.restype = C.c_int32
        o.orc_block_spin_philox.argtypes = [C.c_int, i32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, i32p]
        o.orc_pyramid.argtypes = [C.c_int, i32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, i64p, C.c_void_p]
        o.orc_pyramid.restype = C.c_int
        o.orc_accumulate.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, f64p]
        o.orc_accumulate_i128.argtypes = [C.c_int, C.c_int, i64p, i64p, i64p, u64p, i64p, u64p]
        o.orc_rg_eigenvalues.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, f64p, f64p, f64p, f64p, f64p]
        o.orc_split_samples.argtypes = [C.c_int, C.c_int, C.c_int]
        o.orc_split_samples.restype = C.c_int
        o.orc_n_transformations.argtypes = [C.c_int, C.c_int]
        o.orc_n_transformations.restype = C.c_int
        o.orc_hot_start.argtypes = [C.c_int, C.c_uint64, C.c_uint32, i32p]
        o.orc_metropolis.argtypes = [C.c_int, i32p, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        o.orc_swendsen_wang.argtypes = [C.c_int, i32p, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        o.orc_metropolis_timing.argtypes = [C.c_int, C.c_double, C.c_int, C.c_uint64]
        o.orc_metropolis_timing.restype = C.c_double
        o.orc_rgnn_scalar_output.argtypes = [C.c_int, i32p, C.c_int, f64p]
        o.orc_rgnn_scalar_output.restype = C.c_double
        o.orc_rgnn_gradient.argtypes = [C.c_int, i32p, C.c_int, f64p, C.c_double, f64p]
        _oracle = o
    return _oracle


def ref():
    global _ref
    if _ref is None:
        r = C.CDLL(REF_SO)
        r.ref_seed.argtypes = [C.c_uint64]
        r.ref_calc_interactions.argtypes = [C.c_int, i32p, f64p]
        r.ref_calc_nn.argtypes = [C.c_int, i32p]
        r.ref_calc_nn.restype = C.c_double
        r.ref_calc_energy.argtypes = [C.c_int, i32p, C.c_double]
        r.ref_calc_energy.restype = C.c_double
        r.ref_calc_magnetization.argtypes = [C.c_int, i32p, C.c_double]
        r.ref_calc_magnetization.restype = C.c_double
        r.ref_block_spin.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p]
        r.ref_block_spin.restype = C.c_int
        r.ref_hot_lattice.argtypes = [C.c_int, i32p]
        r.ref_wolff.argtypes = [C.c_int, i32p, C.c_double, C.c_int]
        r.ref_neighbors.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p]
        r.ref_split_samples.argtypes = [C.c_int, C.c_int, C.c_int]
        r.ref_split_samples.restype = C.c_int
        r.ref_flatten2.argtypes = [f64p, f64p]
        r.ref_rounded_str.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_int]
        r.ref_write_iter.argtypes = [C.c_int]
        r.ref_write_iter.restype = C.c_int
        r.ref_critical_exponent.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, f64p, f64p, C.c_int]
        r.ref_critical_exponent.restype = C.c_int
        r.ref_locate_critical_point.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, f64p, C.c_int, C.POINTER(C.c_double)]
        r.ref_locate_critical_point.restype = C.c_int
        r.ref_mcrg_loop.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        r.ref_mcrg_loop.restype = C.c_double
        r.ref_thermo_series.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, f64p]
        r.ref_rgnn_scalar_output.argtypes = [C.c_int, i32p, C.c_int, f64p]
        r.ref_rgnn_scalar_output.restype = C.c_double
        r.ref_rgnn_gradient.argtypes = [C.c_int, i32p, C.c_int, f64p, C.c_double, f64p]
        _ref = r
    return _ref


# ---------------------------------------------------------------------------------------------------------
# lattice generators shared by the parity tests (reference layout: int32 +-1, column-major => arr[j, i])
# ---------------------------------------------------------------------------------------------------------

def ref_write_iter_py(i):
    """definitions.cpp:44-68 restated: every iteration below 10, every 10^k below 10^(k+1), then every 10^5."""
    step = 1
    bound = 10
    while bound <= 100000:
        if i < bound:
            return i % step == 0
        step = bound
        bound *= 10
    return i % 100000 == 0


def random_lattice(N, seed, p_up=0.5):
    rng = np.random.default_rng(seed)
    return np.where(rng.random((N, N)) < p_up, 1, -1).astype(np.int32)


def pattern_lattices(N):
    """Deterministic edge-case configurations: all up, all down, stripes, checkerboard, single defect."""
    out = {}
    out["up"] = np.ones((N, N), np.int32)
    out["down"] = -np.ones((N, N), np.int32)
    s = np.ones((N, N), np.int32)
    s[::2, :] = -1
    out["stripes_j"] = s.copy()
    out["stripes_i"] = s.T.copy()
    jj, ii = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    out["checker"] = np.where((ii + jj) % 2 == 0, 1, -1).astype(np.int32)
    d = np.ones((N, N), np.int32)
    d[N // 3, N // 2] = -1
    out["defect"] = d
    if N >= 4:
        q = np.ones((N, N), np.int32)
        q[: N // 2, : N // 2] = -1
        out["quadrant"] = q
    return out


def clustered_lattice(N, seed, K=-0.4406868, n_sweeps=30):
    """A correlated (near-critical-looking) configuration made with the oracle's Metropolis from a hot start."""
    o = oracle()
    s = np.empty((N, N), np.int32)
    o.orc_hot_start(N, seed, 0, s)
    o.orc_metropolis(N, s, K, seed, 0, 0, n_sweeps)
    return s


def pyramid(N, spins, seed=1, replica=0, t=0, max_levels=-1, want_levels=False):
    o = oracle()
    n_lv = o.orc_n_transformations(N, 2)
    if 0 <= max_levels < n_lv:
        n_lv = max_levels
    S = np.zeros((n_lv + 1, 4), np.int64)
    total = sum((N >> k) ** 2 for k in range(1, n_lv + 1))
    lev = np.zeros(max(total, 1), np.int32)
    got = o.orc_pyramid(N, np.ascontiguousarray(spins, np.int32), seed, replica, t, max_levels, S,
                        lev.ctypes.data if want_levels else None)
    assert got == n_lv
    if not want_levels:
        return S
    levels = []
    off = 0
    for k in range(1, n_lv + 1):
        n = N >> k
        levels.append(lev[off:off + n * n].reshape(n, n).copy())
        off += n * n
    return S, levels
