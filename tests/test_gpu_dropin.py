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
              {"MCRG_REPLICAS": "1024", "MCRG_SWEEPS_PER_UPDATE": "16", "MCRG_SEED": "4711"})
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


def test_locate_critical_point(tmp_path):
    """Two-lattice matching (mcrg.cpp:146-310).  The reference's own result for L = 16 is K_c(16) = -0.440414806
    (main.cpp:26); it is a fixed point of the iteration, so start there and expect to stay within errors."""
    out = run([APP, "kc", "16", "-0.4404", "2", "2000", "4000000"], tmp_path,
              {"MCRG_REPLICAS": "2048", "MCRG_SWEEPS_PER_UPDATE": "8", "MCRG_QUIET": "1"})
    kc = float(re.search(r"RESULT Kc (\S+)", out).group(1))
    assert abs(kc - (-0.440414806)) < 1.5e-3, kc
    rows = [l for l in (tmp_path / "critical_point_L_16_K_-0.4404.txt").read_text().splitlines() if not l.startswith("#")]
    assert len(rows) == 2 * 3  # 2 iterations x 3 blocking levels, "%25i, %25i, %25.10lf, %25.10lf"
    assert all(re.fullmatch(r"\s+\d+,\s+\d+,\s+-?\d+\.\d{10},\s+-?\d+\.\d{10}", r) for r in rows)


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="ref_main is built only where /root/reference exists")
def test_reference_main_cpp_runs_unchanged(tmp_path):
    """main.cpp:8-17: b=2, N=128, 1e4 equilibration updates, 1e6 samples, at K=-0.44 and at K_c."""
    out = run([REF_MAIN], tmp_path, {"MCRG_REPLICAS": "4096", "MCRG_SWEEPS_PER_UPDATE": "2"}, timeout=900)
    assert out.count("* Critical exponent: nu =") == 2
    for name in ("critical_exponent_N_128_K_-0.44.txt", f"critical_exponent_N_128_K_{KC:.7g}.txt"):
        rows = [l for l in (tmp_path / name).read_text().splitlines() if not l.startswith("#")]
        assert len(rows) == 6  # floor(log 128 / log 2) - 1 blocking levels
        lam = [float(r.split(",")[1]) for r in rows]
        assert all(1.7 < x < 2.3 for x in lam), lam
