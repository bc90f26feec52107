"""Generates the committed fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref/libmcrg_ref.so,
built from /root/reference/src by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py            # everything (a few minutes on 8 cores)
    python tests/golden/make_golden.py --fast     # deterministic vectors only

The reference has no tests or golden vectors of its own, so these are outputs of the reference
itself: deterministic ones (correlators, energy, magnetisation, block spins for a seeded global rng, S logs and
the lambdas its driver wrote) and statistical ones (equilibrium averages of its Wolff sampler and RG eigenvalues
over several seeds, with their scatter).
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402

KC = -0.5 * np.log(1 + np.sqrt(2))
TRAIN_KS = [-0.4320459, -0.4406868, -0.4496804, -0.4688157, -0.489652]  # train.cpp:25, file names of the logs


def deterministic():
    r = _libs.ref()
    out = {}
    idx = []
    for N in (2, 4, 8, 16, 32, 64, 128):
        cases = dict(_libs.pattern_lattices(N))
        for seed in range(3):
            cases[f"rand{seed}"] = _libs.random_lattice(N, 1000 + seed)
        cases["biased"] = _libs.random_lattice(N, 17, p_up=0.75)
        if N >= 4:
            cases["clustered"] = _libs.clustered_lattice(N, 21)
        for name, s in cases.items():
            key = f"N{N}_{name}"
            S = np.zeros(2)
            r.ref_calc_interactions(N, s, S)
            e = r.ref_calc_energy(N, s, KC)
            m = r.ref_calc_magnetization(N, s, KC)
            out[key + "_bits"] = np.packbits((s > 0).ravel())
            out[key + "_ref"] = np.array([S[0], S[1], r.ref_calc_nn(N, s), e, m, float(s.sum())])
            if N >= 4:
                blk = np.zeros((N // 2, N // 2), np.int32)
                r.ref_seed(1234 + N)
                r.ref_block_spin(N, 2, 1, s, blk)
                out[key + "_block_bits"] = np.packbits((blk > 0).ravel())
            idx.append(key)
    out["index"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "deterministic.npz"), **out)
    print("deterministic.npz:", len(idx), "cases")

    # S logs + the lambdas the real driver wrote for the same seed (mcrg.cpp:22-144)
    rg = {}
    for N, n_samples in ((8, 400), (16, 300)):
        lam = np.zeros(16)
        nu = np.zeros(16)
        with tempfile.TemporaryDirectory() as d:
            r.ref_seed(2020)
            n_lv = r.ref_critical_exponent(d.encode(), 100, n_samples, N, KC, lam, nu, 16)
        S_log = np.zeros((n_samples, n_lv + 1, 2))
        lv = C.c_int(0)
        r.ref_seed(2020)
        r.ref_mcrg_loop(100, n_samples, N, KC, 0, S_log.ctypes.data, C.byref(lv))
        rg[f"N{N}"] = dict(n_samples=n_samples, n_lv=n_lv, K=KC, S_log=S_log.astype(np.int64).tolist(),
                           lambdas=lam[:n_lv].tolist(), nus=nu[:n_lv].tolist())
    # RGNN forward/gradient known answers (rgnn.cpp:281-339), W0 of train.cpp:19-23
    W0 = np.array([0.5, 0.5, -0.5, -0.5])
    rgnn = []
    for N in (4, 8):
        for seed in range(3):
            s = _libs.random_lattice(N, 50 + seed)
            g = np.zeros(4)
            r.ref_rgnn_gradient(N, s, 2, W0, 1e-4, g)
            rgnn.append(dict(N=N, seed=50 + seed, out=r.ref_rgnn_scalar_output(N, s, 2, W0), grad=g.tolist()))
    rg["rgnn"] = rgnn
    with open(os.path.join(HERE, "rg_driver.json"), "w") as f:
        json.dump(rg, f)
    print("rg_driver.json written")


def _thermo_worker(args):
    N, K, seed, n_eq, n_samples, stride = args
    r = _libs.ref()
    r.ref_seed(seed)
    out = np.zeros((n_samples, 2))
    r.ref_thermo_series(N, K, n_eq, n_samples, stride, 1, out)
    snn, m = out[:, 0], out[:, 1]
    n2 = float(N * N)
    return [snn.mean() / (4 * n2), np.abs(m).mean() / n2, (m**2).mean() / n2**2, (m**4).mean() / n2**4]


def _lambda_worker(args):
    N, K, seed, n_eq, n_samples = args
    r = _libs.ref()
    lam = np.zeros(16)
    nu = np.zeros(16)
    with tempfile.TemporaryDirectory() as d:
        r.ref_seed(seed)
        n_lv = r.ref_critical_exponent(d.encode(), n_eq, n_samples, N, K, lam, nu, 16)
    return lam[:n_lv].tolist()


def statistical():
    res = {"note": "reference Wolff sampler (ising.cpp:87-155), cold start, mean and standard error over independent seeds",
           "thermo": [], "lambda": []}
    n_seeds = 16
    with mp.Pool(8) as pool:
        for N, K, n_samples in [(8, KC, 40000), (16, KC, 40000), (32, KC, 30000), (64, KC, 12000), (16, -0.40, 40000),
                                (16, -0.48, 40000)] + [(8, k, 40000) for k in TRAIN_KS if abs(k - KC) > 1e-6]:
            rows = np.array(pool.map(_thermo_worker, [(N, K, 7000 + 13 * s, 2000, n_samples, 2) for s in range(n_seeds)]))
            mean, err = rows.mean(0), rows.std(0, ddof=1) / np.sqrt(n_seeds)
            u4 = 1 - rows[:, 3] / (3 * rows[:, 2] ** 2)
            res["thermo"].append(dict(N=N, K=K, n_seeds=n_seeds, n_samples=n_samples,
                                      bond=[mean[0], err[0]], absm=[mean[1], err[1]], m2=[mean[2], err[2]],
                                      m4=[mean[3], err[3]], U4=[u4.mean(), u4.std(ddof=1) / np.sqrt(n_seeds)]))
            print("thermo", N, K, mean, err, flush=True)
        for N, K, n_samples in [(16, KC, 100000), (32, KC, 60000), (64, KC, 30000)]:
            rows = np.array(pool.map(_lambda_worker, [(N, K, 9000 + 7 * s, 3000, n_samples) for s in range(n_seeds)]))
            res["lambda"].append(dict(N=N, K=K, n_seeds=n_seeds, n_samples=n_samples, mean=rows.mean(0).tolist(),
                                      err=(rows.std(0, ddof=1) / np.sqrt(n_seeds)).tolist(),
                                      sd_single_run=rows.std(0, ddof=1).tolist()))
            print("lambda", N, rows.mean(0), rows.std(0, ddof=1) / np.sqrt(n_seeds), flush=True)
    with open(os.path.join(HERE, "statistical.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("statistical.json written")


def train_log_rows():
    """Cycle-0 rows of the training logs checked into the reference repo (train_scalar_b2_L8_K*.txt:3): the weights
    are the known W0 = [.5 -.5; .5 -.5] there (line 1), so the row pins <u_L>, Var u_L, <u_S>, Var u_S of the sampler +
    RGNN forward pass at L = 8 / 4 for five couplings, from 1e4 samples each (train.cpp:8-10)."""
    import glob

    rows = []
    for path in sorted(glob.glob("/root/reference/train_scalar_b2_L8_K*.txt")):
        with open(path) as f:
            lines = f.read().splitlines()
        w0 = [float(x) for x in lines[0].split(":")[1].split()]
        c0 = [float(x) for x in lines[2].split()]
        assert c0[0] == 0
        K = float(os.path.basename(path).split("_K")[1][:-4])
        rows.append(dict(file=os.path.basename(path), K=K, W0=w0, n_samples=10000, uL=c0[1], varL=c0[2], uS=c0[3], varS=c0[4],
                         mse=c0[5], grad_norm=c0[6]))
    with open(os.path.join(HERE, "train_logs_cycle0.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("train_logs_cycle0.json:", len(rows), "rows")


def _kc_worker(args):
    L, K0, seed, n_iter, n_eq, n_samples = args
    r = _libs.ref()
    out = np.zeros(256)
    kf = C.c_double(0)
    with tempfile.TemporaryDirectory() as d:
        r.ref_seed(seed)
        n = r.ref_locate_critical_point(d.encode(), n_iter, n_eq, n_samples, L, K0, out, 256, C.byref(kf))
    return out[:n].tolist()


def critical_point(procs=6):
    """The reference's own locate_critical_point (mcrg.cpp:146-310) over independent seeds: K after ONE iteration from
    displaced starting points (per blocking level: the response of the two-lattice estimator, sign and size), and the
    iterated fixed point.  Mean and standard error over seeds -> tests/golden/critical_point.json."""
    res = {"note": "reference locate_critical_point, Wolff sampler, hot start; rows are [iteration][level] Kc; mean and "
                   "standard error over independent seeds of the global rng", "runs": []}
    n_seeds = 16
    plan = [(16, -0.43, 1, 2000, 1000000), (16, -0.45, 1, 2000, 1000000), (16, -0.44, 3, 2000, 1000000),
            (32, -0.43, 1, 2000, 600000), (32, -0.45, 1, 2000, 600000), (32, -0.44, 3, 2000, 600000),
            (64, -0.4405, 2, 3000, 200000)]
    with mp.Pool(procs) as pool:
        for L, K0, n_iter, n_eq, n_samples in plan:
            rows = np.array(pool.map(_kc_worker, [(L, K0, 31000 + 17 * s + L, n_iter, n_eq, n_samples) for s in range(n_seeds)], chunksize=1))
            n_lv = rows.shape[1] // n_iter
            rows = rows.reshape(n_seeds, n_iter, n_lv)
            res["runs"].append(dict(L=L, K0=K0, n_iterations=n_iter, n_eq=n_eq, n_samples=n_samples, n_seeds=n_seeds,
                                    mean=rows.mean(0).tolist(), err=(rows.std(0, ddof=1) / np.sqrt(n_seeds)).tolist(),
                                    sd_single_run=rows.std(0, ddof=1).tolist()))
            print("kc", L, K0, rows.mean(0)[-1], rows.std(0, ddof=1)[-1] / np.sqrt(n_seeds), flush=True)
            with open(os.path.join(HERE, "critical_point.json"), "w") as f:
                json.dump(res, f, indent=1)
    print("critical_point.json written")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--fast", action="store_true")
    ap.add_argument("--kc", action="store_true", help="only the critical-point fixtures (about an hour on 6 cores)")
    a = ap.parse_args()
    if not _libs.ref_available():
        sys.exit("oracle/_ref/libmcrg_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    if a.kc:
        critical_point()
        sys.exit(0)
    deterministic()
    train_log_rows()
    if not a.fast:
        statistical()

