"""Exact finite-size results for the 2D Ising model on the L x L torus, used as known answers that do not depend on any
sampler (neither the reference's Wolff update nor ours):

  * Kaufman's partition function (B. Kaufman, Phys. Rev. 76, 1232 (1949)): <s_i s_j> over nearest-neighbour bonds,
    i.e. the reference's S_nn / (4 N^2) (lattice.cpp:84-99 counts every bond twice), for L = 4 ... 64 at the couplings
    used by the reference's drivers (main.cpp:13, train.cpp:25);
  * brute-force enumeration of the 4 x 4 torus: bond, <|M|>, <M^2>, <M^4> per spin power.

Run:  python tests/golden/make_exact.py   (needs mpmath; writes tests/golden/exact_ising.json)
"""
import itertools
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def ln_Z(K, m, n):
    K = mp.mpf(K)

    def gamma(l):
        if l == 0:
            return 2 * K + mp.log(mp.tanh(K))
        return mp.acosh(mp.cosh(2 * K) * mp.coth(2 * K) - mp.cos(l * mp.pi / n))

    Z = [mp.mpf(1)] * 4
    for r in range(n):
        g1, g0 = gamma(2 * r + 1), gamma(2 * r)
        Z[0] *= 2 * mp.cosh(m * g1 / 2)
        Z[1] *= 2 * mp.sinh(m * g1 / 2)
        Z[2] *= 2 * mp.cosh(m * g0 / 2)
        Z[3] *= 2 * mp.sinh(m * g0 / 2)
    return mp.log(mp.mpf(1) / 2 * (2 * mp.sinh(2 * K)) ** (mp.mpf(m * n) / 2) * sum(Z))


def bond(L, absK):
    return float(mp.diff(lambda k: ln_Z(k, L, L), mp.mpf(absK)) / (2 * L * L))


def enumerate_4x4(absK):
    L = 4
    st = np.array(list(itertools.product([-1, 1], repeat=16)), np.int64).reshape(-1, L, L)
    bonds = (st * np.roll(st, 1, 1)).sum(axis=(1, 2)) + (st * np.roll(st, 1, 2)).sum(axis=(1, 2))
    w = np.exp(absK * bonds.astype(np.float64))
    Z = w.sum()
    M = st.sum(axis=(1, 2)).astype(np.float64)
    return dict(bond=float((w * bonds).sum() / Z / 32.0), absm=float((w * np.abs(M)).sum() / Z / 16.0),
                m2=float((w * M**2).sum() / Z / 16.0**2), m4=float((w * M**4).sum() / Z / 16.0**4))


def main():
    Kc = float(mp.log(1 + mp.sqrt(2)) / 2)
    couplings = [Kc, 0.40, 0.48, 0.4320459, 0.4496804, 0.4688157, 0.489652]
    out = {"note": "absK = |K|; the reference's ferromagnet has K = -absK (ising.cpp:8-9)", "kaufman_bond": [], "enum_4x4": []}
    for L in (4, 8, 16, 32, 64):
        for aK in couplings:
            out["kaufman_bond"].append({"L": L, "absK": aK, "bond": bond(L, aK)})
    for aK in couplings[:3]:
        e = enumerate_4x4(aK)
        e["absK"] = aK
        out["enum_4x4"].append(e)
        k = next(t for t in out["kaufman_bond"] if t["L"] == 4 and t["absK"] == aK)
        assert abs(k["bond"] - e["bond"]) < 1e-12, (k, e)  # the two exact methods agree
    with open(os.path.join(HERE, "exact_ising.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote exact_ising.json:", len(out["kaufman_bond"]), "Kaufman values")


if __name__ == "__main__":
    main()
