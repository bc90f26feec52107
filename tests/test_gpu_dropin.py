"""The C++ drop-in layer (mcrg_b200/host): the reference's class API over the C ABI, exercised through the apps.

ref_main is the REFERENCE's own src/main.cpp compiled unchanged against mcrg_b200/host/include (built in the
container where /root/reference exists; the binary travels to the GPU box)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _libs

pytestmark = pytest.mark.gpu

BUILD = os.path.join(_libs.ROOT, "mcrg_b200", "host", "_build")
APP = os.path.join(BUILD, "mcrg_app")
REF_MAIN = os.path.join(BUILD, "ref_main")
KC = float(-0.5 * np.log(1 + np.sqrt(2)))  # a plain float: repr() must be parseable by atof


def run(cmd, cwd, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, cwd=cwd, env=e, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("N", [2, 8, 64, 256])
def test_lattice_and_ising_methods(tmp_path, N):
    """Lattice::calc_interactions on the device == the same sums recomputed on the host from the public spins_
    member with nearest_neighbors()/next_nearest_neighbors(); calc_energy / calc_magnetization follow ising.cpp."""
    out = run([APP, "lattice", str(N), repr(KC), "20"], tmp_path, {"MCRG_QUIET": "1"})
    m = re.search(r"RESULT Snn (\S+) (\S+) Snnn (\S+) (\S+) E (\S+) M (\S+) sum (\S+)", out)
    assert m, out
    snn_d, snn_h, snnn_d, snnn_h, E, M, tot = [float(x) for x in m.groups()]
    assert snn_d == snn_h and snnn_d == snnn_h
    assert abs(E - KC * snn_d / (N * N)) < 1e-12
    assert M == float(int(tot / (N * N)))


def test_calc_critical_exponent_matches_reference_driver(tmp_path):
    with open(os.path.join(_libs.ROOT, "tests", "golden", "statistical.json")) as f:
        ref = next(t for t in json.load(f)["lambda"] if t["N"] == 32)
    out = run([APP, "exponent", "32", repr(KC), "500", "2000000"], tmp_path,
              {"MCRG_REPLICAS": "1024", "MCRG_UPDATE": "metropolis", "MCRG_SWEEPS_PER_UPDATE": "16", "MCRG_SEED": "4711"})
    # console and file formats of mcrg.cpp:12-17, 31-39, 133-141
    assert "==========  RENORMALIZATION GROUP  ==========" in out and "* Scaling factor b = 2" in out
    assert re.search(r"n = 0: lambda = \d\.\d{6}, nu = \d\.\d{6}", out)
    fn = tmp_path / f"critical_exponent_N_32_K_{KC:.7g}.txt"
    assert fn.exists(), os.listdir(tmp_path)
    lines = fn.read_text().splitlines()
    assert lines[0].startswith("# Number of parallel processes = 1024") and lines[2] == "# Number of samples = 2000000"
    rows = [l for l in lines if not l.startswith("#")]
    assert len(rows) == 4 and all(re.fullmatch(r"\s+\d+,\s+\d+\.\d{10},\s+\d+\.\d{10}", r) for r in rows)
    res = re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+) nu (\S+)", out)
    assert len(res) == 4
    for lv, lam, err, nu in res:
        lv, lam, err = int(lv), float(lam), float(err)
        sigma = np.hypot(err, ref["err"][lv])
        assert abs(lam - ref["mean"][lv]) < 3 * sigma, (lv, lam, err, ref["mean"][lv], ref["err"][lv])
        assert abs(float(rows[lv].split(",")[1]) - lam) < 1e-9


def test_calc_critical_exponent_with_cluster_updates(tmp_path):
    """MCRG_UPDATE=cluster: the driver keeps the reference's schedule — ONE (cluster) update per sample, mcrg.cpp:72-98
    — and reproduces the reference driver's lambda per level."""
    with open(os.path.join(_libs.ROOT, "tests", "golden", "statistical.json")) as f:
        ref = next(t for t in json.load(f)["lambda"] if t["N"] == 32)
    out = run([APP, "exponent", "32", repr(KC), "300", "2000000"], tmp_path,
              {"MCRG_REPLICAS": "1024", "MCRG_UPDATE": "cluster", "MCRG_SEED": "99", "MCRG_QUIET": "1"})
    res = re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+) nu (\S+)", out)
    assert len(res) == 4
    for lv, lam, err, nu in res:
        lv, lam, err = int(lv), float(lam), float(err)
        sigma = np.hypot(err, ref["err"][lv])
        assert abs(lam - ref["mean"][lv]) < 3 * sigma, (lv, lam, err, ref["mean"][lv], ref["err"][lv])


def _kc_golden():
    with open(os.path.join(_libs.ROOT, "tests", "golden", "critical_point.json")) as f:
        return json.load(f)["runs"]


@pytest.mark.parametrize("L,K0", [(16, -0.43), (16, -0.45), (32, -0.43), (32, -0.45), (16, -0.44), (32, -0.44), (64, -0.4405)])
def test_locate_critical_point_matches_reference(tmp_path, L, K0):
    """Two-lattice matching (mcrg.cpp:146-310) against the COMPILED REFERENCE's own locate_critical_point: K per blocking
    level after one iteration from the same starting point, reference mean and standard error over 16 seeds
    (tests/golden/critical_point.json, made by tests/golden/make_golden.py --kc) against ours with the jackknife error over
    groups of chains; 3 sigma combined.  The starting points are displaced from the fixed point on both sides (-0.43, -0.45:
    one iteration must move K the right way by the right amount at every level) and near it (-0.44, -0.4405)."""
    ref = next(r for r in _kc_golden() if r["L"] == L and abs(r["K0"] - K0) < 1e-12)
    out = run([APP, "kc", str(L), repr(K0), "1", "1000", "8000000"], tmp_path, {"MCRG_REPLICAS": "4096", "MCRG_QUIET": "1", "MCRG_SEED": str(7 + L)})
    res = re.findall(r"RESULT level (\d+) Kc (\S+) err (\S+)", out)
    n_lv = len(ref["mean"][0])
    assert len(res) == n_lv == int(np.log2(L)) - 1
    for lv, kc, err in res:
        lv, kc, err = int(lv), float(kc), float(err)
        want, werr = ref["mean"][0][lv], ref["err"][0][lv]
        assert abs(kc - want) < 3 * np.hypot(err, werr), (L, K0, lv, kc, err, want, werr)
        if abs(K0 - (-0.4407)) > 5e-3:  # displaced start: the step points towards the critical coupling
            assert (kc - K0) * (-0.4406868 - K0) > 0, (L, K0, lv, kc)
    rows = [l for l in (tmp_path / f"critical_point_L_{L}_K_{K0:.7g}.txt").read_text().splitlines() if not l.startswith("#")]
    assert len(rows) == n_lv  # 1 iteration x n_lv blocking levels, "%25i, %25i, %25.10lf, %25.10lf"
    assert all(re.fullmatch(r"\s+\d+,\s+\d+,\s+-?\d+\.\d{10},\s+-?\d+\.\d{10}", r) for r in rows)


def test_locate_critical_point_iterates(tmp_path):
    """Three iterations from K0 = -0.44 at L = 16, as the reference run in the golden file: the K returned after every
    iteration feeds the next one; the final K per level against the reference's (its scatter over seeds includes the
    noise accumulated over the iterations; ours is taken 3 times the last iteration's jackknife error)."""
    ref = next(r for r in _kc_golden() if r["L"] == 16 and r["n_iterations"] == 3)
    out = run([APP, "kc", "16", "-0.44", "3", "1000", "8000000"], tmp_path, {"MCRG_REPLICAS": "4096", "MCRG_QUIET": "1", "MCRG_SEED": "99"})
    rows = [l for l in (tmp_path / "critical_point_L_16_K_-0.44.txt").read_text().splitlines() if not l.startswith("#")]
    assert len(rows) == 3 * 3
    t = np.array([[float(x) for x in r.split(",")] for r in rows])
    assert np.array_equal(t[:, 0], np.repeat([1, 2, 3], 3)) and np.array_equal(t[:, 1], np.tile([0, 1, 2], 3))
    assert t[0, 2] == -0.44 and t[3, 2] == t[2, 3] and t[6, 2] == t[5, 3]  # starting K of an iteration = last level's Kc of the one before
    res = {int(lv): (float(kc), float(err)) for lv, kc, err in re.findall(r"RESULT level (\d+) Kc (\S+) err (\S+)", out)}
    for lv in range(3):
        kc, err = res[lv]
        assert abs(kc - ref["mean"][2][lv]) < 3 * np.hypot(3 * err, ref["err"][2][lv]), (lv, kc, err, ref["mean"][2][lv], ref["err"][2][lv])


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="ref_main is built only where /root/reference exists")
def test_reference_main_cpp_runs_unchanged(tmp_path):
    """main.cpp:8-17: b=2, N=128, 1e4 equilibration updates, 1e6 samples, at K=-0.44 and at K_c, with the drop-in's defaults:
    at N = 128 those are cluster updates (the reference's update family), one per sample as in mcrg.cpp:75, so the 1e4
    equilibration updates equilibrate and the chains decorrelate like the reference's.  lambda per level must sit where the
    2D Ising RG puts it: 1.95 at the first level (the compiled reference gives 1.951 +- 0.003 at N = 64,
    tests/golden/statistical.json), 2.0 on the inner levels, 2.02 on the last one (the 2 x 2 lattice)."""
    out = run([REF_MAIN], tmp_path, {"MCRG_REPLICAS": "4096"}, timeout=900)
    assert out.count("* Critical exponent: nu =") == 2
    lam = {}
    for key, name in (("off", "critical_exponent_N_128_K_-0.44.txt"), ("Kc", f"critical_exponent_N_128_K_{KC:.7g}.txt")):
        rows = [l for l in (tmp_path / name).read_text().splitlines() if not l.startswith("#")]
        assert len(rows) == 6  # floor(log 128 / log 2) - 1 blocking levels
        lam[key] = [float(r.split(",")[1]) for r in rows]
    # at K_c every level sits at the fixed point
    assert 1.93 < lam["Kc"][0] < 1.97, lam
    assert all(1.97 < x < 2.02 for x in lam["Kc"][1:5]), lam
    assert 1.99 < lam["Kc"][5] < 2.06, lam
    # K = -0.44 is 0.16 % above T_c: the first levels see the fixed point, the deep ones have flowed away from it (the
    # deviation doubles per level), so their eigenvalue is no longer 2 — only a sanity window there
    assert 1.93 < lam["off"][0] < 1.97 and all(1.97 < x < 2.02 for x in lam["off"][1:3]), lam
    assert all(1.85 < x < 2.1 for x in lam["off"][3:]), lam


def test_equilibrate_log_is_averaged_over_chains(tmp_path):
    """IsingModel::equilibrate(write=true), ising.cpp:22-74: the log averages E, |M|, E^2, M^2 over the ranks; here over
    MCRG_REPLICAS device chains.  File name, header and row format are the reference's; rows appear at its write_iter
    schedule; sigma_E, C, chi are non-zero; E converges to the reference sampler's equilibrium value; MCRG_COMPAT=1
    (default) reproduces the reference's integer-division magnetisation and second division by n_spins."""
    with open(os.path.join(_libs.ROOT, "tests", "golden", "statistical.json")) as f:
        ref = next(t for t in json.load(f)["thermo"] if t["N"] == 32 and abs(t["K"] - KC) < 1e-9)
    N, n_eq, R = 32, 2000, 2048
    logs = {}
    for compat in ("1", "0"):
        d = tmp_path / f"compat{compat}"
        d.mkdir()
        run([APP, "equilibrate", str(N), repr(KC), str(n_eq)], d,
            {"MCRG_REPLICAS": str(R), "MCRG_SEED": "31", "MCRG_COMPAT": compat, "MCRG_QUIET": "1"})
        fn = d / f"equilibrate_N_{N}_K_{KC:.7g}.txt"
        assert fn.exists(), os.listdir(d)
        logs[compat] = fn.read_text().splitlines()
    for compat, lines in logs.items():
        assert lines[0] == f"# Nearest neighbor coupling K = {KC:.6f}" and lines[1] == f"# Temperature T = {-1 / KC:.6f}"
        assert lines[2] == f"# Number of lattice sites = {N * N}" and lines[3] == "# Lattice spacing = 1"
        assert lines[4] == f"# Using {R} parallel processes"
        assert lines[5] == "# Iteration, Avg E/spin, Stddev E/spin, Heat Capacity, Avg |M|/spin, Stddev |M|/spin, Susceptibility"
        rows = [l for l in lines[6:] if l and not l.startswith("#")]
        iters = [int(r.split(",")[0]) for r in rows]
        assert iters == [n for n in range(1, n_eq + 1) if _libs.ref_write_iter_py(n)]
        assert all(re.fullmatch(r"\d+(, -?\d\.\d{7}e[+-]\d\d){6}", r) for r in rows), rows[:2]
        spins = [l for l in lines if l.startswith("# ") and l.rstrip().endswith(",")]
        assert len(spins) == N and all(len(l[2:].split(",")) == N + 1 for l in spins)  # write_spins, lattice.cpp:58-70
    # np.genfromtxt-style parse (comment lines skipped), as the reference's plotting scripts read such tables
    fixed = np.genfromtxt(logs["0"], delimiter=",", comments="#")
    compat = np.genfromtxt(logs["1"], delimiter=",", comments="#")
    assert fixed.shape == compat.shape and fixed.shape[1] == 7
    last = fixed[-1]
    assert last[2] > 0 and last[3] > 0 and last[5] > 0 and last[6] > 0  # sigma_E, C, sigma_M, chi: averages over > 1 chain
    e_ref, e_err = 4 * KC * ref["bond"][0], 4 * abs(KC) * ref["bond"][1]
    assert abs(last[1] - e_ref) < 4 * np.hypot(e_err, last[2] / np.sqrt(R)), (last[1], e_ref)
    assert abs(last[4] - ref["absm"][0]) < 4 * np.hypot(ref["absm"][1], last[5] / np.sqrt(R)), (last[4], ref["absm"])
    # same seed => the same device chains (chain 0, the caller's Lattice(N), starts from the host's non-deterministic rng
    # like the reference's, so 1 chain in 2048 differs): the compat log is the same energy divided by n_spins once more
    # (ising.cpp:72), and its magnetisation column is the integer division of ising.cpp:178 (0 unless a chain is fully ordered)
    assert np.allclose(compat[:, 1] * N * N, fixed[:, 1], rtol=2e-3)
    assert (compat[:, 4] == 0).all()


def test_test_scalar_output_file(tmp_path):
    """RenormalizationGroupNeuralNetwork::test_scalar_output, rgnn.cpp:192-278: 101 couplings K0 - DeltaK .. K0 + DeltaK in steps
    of DeltaK / 50, one row "K T <u_L> Var <u_S> Var mse" each in the reference's format; at K0 the row agrees with the
    cycle-0 row of the reference's own training log at that coupling (same W0, same observable)."""
    with open(os.path.join(_libs.ROOT, "tests", "golden", "train_logs_cycle0.json")) as f:
        row = next(r for r in json.load(f) if abs(r["K"] - (-0.4406868)) < 1e-9)
    K0, DK = -0.4406868, 0.02
    out = run([APP, "test", "8", repr(K0), repr(DK), "400000", "500"], tmp_path, {"MCRG_REPLICAS": "2048", "MCRG_SWEEPS_PER_UPDATE": "4", "MCRG_QUIET": "1"})
    assert "RESULT done" in out
    lines = (tmp_path / f"test_scalar_b2_L8_K{K0:.7g}.txt").read_text().splitlines()
    assert lines[0] == "" and lines[1] == "# Coupling K, Temperature T, Avg Output L, Var Output L, Avg Output S, Var Output S, MSE"
    rows = lines[2:]
    assert len(rows) in (100, 101)  # the reference's floating-point loop bound (K <= K0 + DeltaK after 100 additions)
    assert all(re.fullmatch(r"(\s*-?\d\.\d{7}e[+-]\d\d){7}", r) for r in rows)
    t = np.array([[float(x) for x in r.split()] for r in rows])
    assert np.allclose(t[:, 1], -1.0 / t[:, 0], rtol=1e-6) and np.allclose(np.diff(t[:, 0]), DK / 50, rtol=1e-4)
    assert np.allclose(t[:, 6], (t[:, 2] - t[:, 4]) ** 2, rtol=1e-4, atol=1e-12)
    mid = t[50]
    assert abs(mid[0] - K0) < 1e-9
    n_eff = 400000 / 4  # generous: Metropolis at L = 8 with 4 sweeps per sample
    assert abs(mid[2] - row["uL"]) < 3 * np.hypot(1.5 * np.sqrt(row["varL"] / row["n_samples"]), np.sqrt(mid[3] / n_eff)), (mid, row)
    assert abs(mid[4] - row["uS"]) < 3 * np.hypot(1.5 * np.sqrt(row["varS"] / row["n_samples"]), np.sqrt(mid[5] / n_eff)), (mid, row)
    # the output moves monotonically with the coupling across the window (it measures disorder inside the 2 x 2 blocks)
    assert abs(np.corrcoef(t[:, 0], t[:, 2])[0, 1]) > 0.9


def test_rgnn_cycle0_matches_reference_training_logs(tmp_path):
    """Known answers from the reference's own checked-in logs (train_scalar_b2_L8_K*.txt:3, via
    tests/golden/train_logs_cycle0.json): with W0 = [.5 -.5; .5 -.5], <u_L=8>, Var, <u_S=4>, Var at five couplings.
    The logged row is a 1e4-sample average, so its statistical error is sqrt(Var/1e4) (inflated x1.5 for
    autocorrelation); ours uses 4e6 samples.  3 sigma."""
    with open(os.path.join(_libs.ROOT, "tests", "golden", "train_logs_cycle0.json")) as f:
        rows = json.load(f)
    assert len(rows) == 5
    for row in rows:
        d = tmp_path / f"K{row['K']}"
        d.mkdir()
        out = run([APP, "train", "8", repr(float(row["K"])), "0", "4000000", "2000"], d,
                  {"MCRG_REPLICAS": "4096", "MCRG_SWEEPS_PER_UPDATE": "4", "MCRG_QUIET": "1"})
        assert "RESULT final_mse" in out
        log = (d / f"train_scalar_b2_L8_K{row['K']:.7g}.txt").read_text().splitlines()
        assert log[0].startswith("# Initial Weights:") and [float(x) for x in log[0].split(":")[1].split()] == row["W0"]
        assert log[1] == "# Cycles, Avg Output L, Var Output L, Avg Output S, Var Output S, MSE, || MSE Gradient ||"
        c0 = [float(x) for x in log[2].split()]
        assert re.fullmatch(r"\s+0(\s+-?\d\.\d{7}e[+-]\d\d){6}", log[2])
        for got, want, var in ((c0[1], row["uL"], row["varL"]), (c0[3], row["uS"], row["varS"])):
            assert abs(got - want) < 3 * 1.5 * np.sqrt(var / row["n_samples"]), (row["K"], got, want)
        for got, want in ((c0[2], row["varL"]), (c0[4], row["varS"])):
            assert abs(got - want) < 0.06 * want, (row["K"], got, want)  # variance of a variance: ~sqrt(2/1e4) x kurtosis
        assert any(l.startswith("# Final Weights:") for l in log) and any(l.startswith("# Example Flow:") for l in log)


def test_rgnn_training_reduces_the_cost(tmp_path):
    """train.cpp's experiment in miniature: minimise (<u_L> - <u_{L/2}>)^2 with ADAM from W0; the cost must fall and
    the weights keep the symmetry the reference's finished runs show (W00 = W10, W01 = W11: its Final Weights rows)."""
    out = run([APP, "train", "8", repr(KC), "300", "20000", "1000"], tmp_path, {"MCRG_REPLICAS": "4096", "MCRG_QUIET": "1"})
    m = re.search(r"RESULT final_mse (\S+) W (\S+) (\S+) (\S+) (\S+)", out)
    mse, w00, w01, w10, w11 = [float(x) for x in m.groups()]
    log = [l for l in (tmp_path / f"train_scalar_b2_L8_K{KC:.7g}.txt").read_text().splitlines() if l and not l.startswith("#")]
    assert len(log) == 301
    first, last = float(log[0].split()[5]), np.mean([float(l.split()[5]) for l in log[-20:]])
    assert last < 0.5 * first, (first, last)
    assert abs(w00 - w10) < 0.05 and abs(w01 - w11) < 0.05, (w00, w01, w10, w11)


@pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "ref_train")), reason="ref_train is built only where /root/reference exists")
def test_reference_train_cpp_runs_unchanged(tmp_path):
    """train.cpp:7-34 as shipped: 6 temperatures x 1e4 cycles x 1e4 samples at L = 8 (hours on the reference's cluster)."""
    out = run([os.path.join(BUILD, "ref_train")], tmp_path, {"MCRG_REPLICAS": "2048", "MCRG_QUIET": "1"}, timeout=1500)
    rows = [l.split() for l in out.splitlines() if re.fullmatch(r"\d\.\d+e[+-]\d+ \d\.\d+e[+-]\d+", l.strip())]
    assert len(rows) == 6  # printf("%e %e\n", T, final_mse_), train.cpp:33
    logs = sorted(p.name for p in tmp_path.glob("train_scalar_b2_L8_K*.txt"))
    assert len(logs) == 6
    for name in logs:
        lines = (tmp_path / name).read_text().splitlines()
        assert sum(1 for l in lines if l and not l.startswith("#")) == 10001
