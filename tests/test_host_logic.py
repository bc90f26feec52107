"""CPU-only checks of the host-side logic and of the C-ABI library as a binary (no compute calls: no GPU here)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import _libs

ROOT = _libs.ROOT
KC = -0.5 * np.log(1 + np.sqrt(2))


@pytest.fixture(scope="module")
def mc():
    from mcrg_b200 import build

    build.build_library()  # nvcc cross-compiles sm_100a without a GPU
    import mcrg_b200

    return mcrg_b200


def test_library_exports_every_declared_symbol(mc):
    header = open(os.path.join(ROOT, "include", "mcrg_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mcrg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = C.CDLL(mc.capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mcrg_b200.h but not exported"
    # and nothing CUDA- or torch-typed leaks through the signatures
    assert "cudaStream_t" not in header and "torch" not in header


def test_library_is_sm100a_and_has_no_cpu_fallback(mc):
    out = subprocess.run(["cuobjdump", "-lelf", mc.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # without a device every compute entry point fails loudly
    with pytest.raises(mc.McrgError):
        mc.Context(64, 1)
    lay = mc.capi.acc_layout()
    assert lay.n_slots == 3 + 3 * 16 + 9 * 16 + 9 * 15 + 9 * 15 + 3 and lay.slot_m4 == lay.n_slots - 3  # + the three exact parts of sum M^4
    assert mc.capi.levels_full(128) == 6 and mc.capi.levels_full(4096) == 11 and mc.capi.levels_full(16384) == 13


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may reference oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mcrg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may cite the oracle as the specification; code may not include, import, link or load it
                for pat in (r'#\s*include\s*[<"][^>"]*oracle', r"\bimport\s+_libs\b", r"libmcrg_oracle", r"libmcrg_ref",
                            r"mcrg_oracle\.h", r"-I\S*oracle", r"\borc_[a-z_0-9]+\s*\(", r"\bref_[a-z_0-9]+\s*\("):
                    assert not re.search(pat, text), (os.path.join(dirpath, f), pat)


def _fake_acc(mc, S_log):
    """Build an accumulator vector (exact ints) from a list of per-sample S matrices [n_lv+1][3]."""
    lay = mc.capi.acc_layout()
    acc = [0] * lay.n_slots
    n_lv = S_log.shape[1] - 1
    for S in S_log:
        acc[lay.slot_n] += 1
        for lv in range(n_lv + 1):
            for a in range(3):
                acc[lay.slot_s + lv * 3 + a] += int(S[lv, a])
                for b in range(3):
                    acc[lay.slot_ss + lv * 9 + b * 3 + a] += int(S[lv, a]) * int(S[lv, b])
                    if lv >= 1:
                        acc[lay.slot_sbs + (lv - 1) * 9 + b * 3 + a] += int(S[lv, a]) * int(S[lv - 1, b])
                        acc[lay.slot_sb0 + (lv - 1) * 9 + b * 3 + a] += int(S[lv, a]) * int(S[0, b])
    return acc


def test_rg_analysis_reproduces_reference_lambdas(mc):
    """analysis.rg_eigenvalues on exact accumulators built from the reference's own S logs gives the lambdas the
    reference's driver printed (tests/golden/rg_driver.json), and agrees with the oracle's restatement."""
    with open(os.path.join(ROOT, "tests", "golden", "rg_driver.json")) as f:
        rg = json.load(f)
    for key in ("N8", "N16"):
        c = rg[key]
        S2 = np.array(c["S_log"], np.int64)
        S3 = np.concatenate([S2, np.zeros_like(S2[:, :, :1])], axis=2)  # plaquette column unused by ops=(0,1)
        acc = _fake_acc(mc, S3)
        slots = mc.analysis.unpack_slots(acc, c["n_lv"])
        lam, nu = mc.analysis.rg_eigenvalues(slots, ops=(0, 1))
        assert np.allclose(lam, c["lambdas"], rtol=0, atol=5e-9), (lam, c["lambdas"])
        assert np.allclose(nu, c["nus"], rtol=1e-7)


def test_three_operator_analysis_matches_oracle(mc):
    o = _libs.oracle()
    rng = np.random.default_rng(3)
    n, n_lv = 400, 3
    base = rng.normal(size=(n, 1, 1))
    S = np.rint(1000 * (base * np.array([1.0, 0.8, 0.5]) / (2.0 ** np.arange(n_lv + 1))[None, :, None]
                        + 0.3 * rng.normal(size=(n, n_lv + 1, 3))) + 5000).astype(np.int64)
    acc = _fake_acc(mc, S)
    lam, _ = mc.analysis.rg_eigenvalues(mc.analysis.unpack_slots(acc, n_lv), ops=(0, 1, 2))
    S_sum = np.zeros((n_lv + 1) * 3)
    SbS = np.zeros(n_lv * 9)
    SbSb = np.zeros(n_lv * 9)
    for s in S:
        o.orc_accumulate(n_lv, 3, np.ascontiguousarray(s, np.float64).ravel(), S_sum, SbS, SbSb)
    want = np.zeros(n_lv)
    o.orc_rg_eigenvalues(n_lv, 3, float(n), 2, S_sum, SbS, SbSb, want, np.zeros(n_lv))
    assert np.allclose(lam, want, rtol=1e-8)


def test_jackknife(mc):
    rng = np.random.default_rng(0)
    chunks = [[int(x), 100] for x in rng.normal(5000, 50, size=40)]
    full, err = mc.analysis.jackknife(chunks, lambda t: np.array([t[0] / t[1]]))
    vals = np.array([c[0] / c[1] for c in chunks])
    assert abs(full[0] - vals.mean()) < 1e-9
    assert abs(err[0] - vals.std(ddof=1) / np.sqrt(len(vals))) < 1e-9


def test_limb_encoding_roundtrip(mc):
    vals = [0, 1, -1, 2**63, -(2**63) - 5, 2**100 + 12345, -(2**120), 2**32 - 1, -(2**32)]
    limbs = mc.dist.ints_to_limbs(vals)
    assert mc.dist.limbs_to_ints(limbs) == vals
    # sums of limb vectors (what the all-reduce does) decode to sums of values
    assert mc.dist.limbs_to_ints(limbs * 7) == [7 * v for v in vals]
    assert mc.dist.limbs_to_ints(limbs + limbs[::-1]) == [a + b for a, b in zip(vals, vals[::-1])]


@pytest.mark.parametrize("isa", ["", "avx2", "scalar"])
@pytest.mark.parametrize("L", [2, 4, 16, 32, 64, 256])
def test_host_pack_matches_numpy(mc, monkeypatch, L, isa):
    """mcrg_host_pack_i32_colmajor (host-side format conversion, no device): bit k of word w of row y = spin (32w+k, y).
    Every instruction-set path (widest available, AVX2, scalar; MCRG_HOSTPACK_ISA is read per call)."""
    if isa:
        monkeypatch.setenv("MCRG_HOSTPACK_ISA", isa)
    rng = np.random.default_rng(L)
    n = 3
    spins = np.where(rng.random((n, L, L)) < 0.5, 1, -1).astype(np.int32)
    out = np.zeros(mc.capi.packed_words(L, n), np.uint32)
    for threads in (1, 3):
        out[:] = 0
        mc.capi.host_pack(spins.ctypes.data, L, n, out.ctypes.data, threads)
        bits = min(32, L)
        flat = (spins > 0).reshape(-1, bits)
        want = np.zeros(flat.shape[0], np.uint32)
        for k in range(bits):
            want |= flat[:, k].astype(np.uint32) << np.uint32(k)
        assert np.array_equal(out, want)


def test_shard_replicas(mc):
    for n, w in ((40, 1), (40, 8), (5, 2), (7, 4), (4096, 8), (3, 8)):
        seen = []
        for r in range(w):
            first, count = mc.dist.shard_replicas(n, w, r)
            seen += list(range(first, first + count))
        assert seen == list(range(n))


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C2", "--ref-samples", "50", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "G spin-flip attempts/s"
