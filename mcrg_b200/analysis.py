"""Host-side reduction of the device accumulators to RG eigenvalues and thermodynamic averages.

Mirrors the tail of MonteCarloRenormalizationGroup::calc_critical_exponent (mcrg.cpp:106-131): averages,
A = <Sb Sb^T> - <Sb><Sb>^T, B = <Sb S^T> - <Sb><S>^T, T = A^-1 B, lambda = largest real eigenvalue part,
nu = ln b / ln lambda.  The device hands over exact integers (Python ints here), so the only rounding is the
final division — unlike the reference, which accumulates products of ~2^30 numbers in doubles (mcrg.cpp:86-97).
The 2x2 / 3x3 algebra stays on the host, as in the reference.
"""
from fractions import Fraction

import numpy as np

from . import capi

NOP = capi.NOP


def unpack_slots(acc_vec, n_lv, lay=None):
    """acc_vec: sequence of n_slots exact ints (one replica-bin, or a total).  -> dict of arrays of Python ints."""
    lay = lay or capi.acc_layout()
    a = list(acc_vec)
    S = np.array([[a[lay.slot_s + lv * NOP + op] for op in range(NOP)] for lv in range(n_lv + 1)], dtype=object)
    # stored index b*NOP+a holds X_a * Y_b  ->  matrix[a, b]
    SS = np.array([[[a[lay.slot_ss + lv * NOP * NOP + b * NOP + al] for b in range(NOP)] for al in range(NOP)]
                   for lv in range(n_lv + 1)], dtype=object)
    SbS = np.array([[[a[lay.slot_sbs + n * NOP * NOP + b * NOP + al] for b in range(NOP)] for al in range(NOP)]
                    for n in range(n_lv)], dtype=object).reshape(n_lv, NOP, NOP)
    Sb0 = np.array([[[a[lay.slot_sb0 + n * NOP * NOP + b * NOP + al] for b in range(NOP)] for al in range(NOP)]
                    for n in range(n_lv)], dtype=object).reshape(n_lv, NOP, NOP)
    return dict(n=a[lay.slot_n], absM=a[lay.slot_absm], M2=a[lay.slot_m2], S=S, SS=SS, SbS=SbS, Sb0=Sb0)


def _to_float(x, n):
    return float(Fraction(int(x), int(n)))


def rg_eigenvalues(slots, ops=(0, 1), b=2):
    """-> (lambdas[n_lv], nus[n_lv]) for the operator subset `ops` ((0,1) = the reference's NN, NNN pair)."""
    n = slots["n"]
    S, SS, SbS = slots["S"], slots["SS"], slots["SbS"]
    n_lv = SbS.shape[0]
    k = len(ops)
    lam = np.full(n_lv, np.nan)
    nu = np.full(n_lv, np.nan)
    for lv in range(n_lv):
        A = np.empty((k, k))
        B = np.empty((k, k))
        for i, al in enumerate(ops):
            for j, be in enumerate(ops):
                # exact covariance numerators: n*sum(xy) - sum(x)sum(y), divided once by n^2
                A[i, j] = float(Fraction(int(n) * int(SS[lv + 1][al, be]) - int(S[lv + 1][al]) * int(S[lv + 1][be]), int(n) ** 2))
                B[i, j] = float(Fraction(int(n) * int(SbS[lv][al, be]) - int(S[lv + 1][al]) * int(S[lv][be]), int(n) ** 2))
        try:
            T = np.linalg.solve(A, B)
            ev = np.linalg.eigvals(T)
            lam[lv] = np.max(ev.real)
            nu[lv] = np.log(b) / np.log(lam[lv]) if lam[lv] > 0 else np.nan
        except np.linalg.LinAlgError:
            pass
    return lam, nu


def thermo(slots, accd_m4, L):
    """-> dict with <E>/spin... in the conventions used by the tests:
    e = <S_nn>/(2 L^2) (bond energy per site in units where each bond counts once... see below),
    absm = <|M|>/L^2, m2 = <M^2>/L^4, m4 = <M^4>/L^8, U4 = 1 - m4/(3 m2^2).
    S_nn is the reference's double-counted sum (lattice.cpp:102-120), so <S_nn>/(4 L^2) is the mean bond
    alignment in [-1, 1]."""
    n = int(slots["n"])
    N = L * L
    snn = _to_float(slots["S"][0][0], n)
    m2 = _to_float(slots["M2"], n) / N**2
    m4 = accd_m4 / n / float(N) ** 4
    return dict(bond=snn / (4.0 * N), absm=_to_float(slots["absM"], n) / N, m2=m2, m4=m4,
                U4=1.0 - m4 / (3.0 * m2 * m2) if m2 > 0 else np.nan)


def jackknife(chunks, fn):
    """chunks: list of per-chunk accumulator vectors (exact ints); fn(total_vector) -> np.ndarray.
    Returns (estimate from the total, jackknife standard error)."""
    k = len(chunks)
    total = [sum(col) for col in zip(*chunks)]
    full = np.asarray(fn(total), dtype=float)
    if k < 2:
        return full, np.full_like(full, np.nan)
    loo = np.array([np.asarray(fn([t - c for t, c in zip(total, ch)]), dtype=float) for ch in chunks])
    mean = loo.mean(axis=0)
    err = np.sqrt((k - 1) / k * ((loo - mean) ** 2).sum(axis=0))
    return full, err
