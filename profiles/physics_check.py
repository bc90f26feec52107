"""End-to-end physics on the B200 through the drop-in drivers, next to the reference's own numbers.  Run on the GPU box:
    python profiles/physics_check.py | tee gpurun_out/physics_r1.txt

  (1) K_c(L) from the two-lattice matching (MonteCarloRenormalizationGroup::locate_critical_point, mcrg.cpp:146-310)
      against the values the reference's author recorded in main.cpp:26-29;
  (2) lambda per blocking level from calc_critical_exponent (mcrg.cpp:72-131) at N = 64 against the compiled reference
      (tests/golden/statistical.json), Metropolis and cluster updates;
  (3) the same at N = 4096 with cluster updates: sizes the reference cannot reach (its own largest run is N = 128).
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "mcrg_b200", "host", "_build", "mcrg_app")
KC = -0.44068679350977147
REF_KC = {16: -0.440414806, 32: -0.440619, 64: -0.440675, 128: -0.440683}  # main.cpp:26-29


def run(args, env):
    e = dict(os.environ, MCRG_QUIET="1", **env)
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        out = subprocess.run([APP] + [str(a) for a in args], cwd=d, env=e, capture_output=True, text=True, timeout=1800)
        dt = time.time() - t0
    if out.returncode != 0:
        raise RuntimeError(out.stdout[-1500:] + out.stderr[-1500:])
    return out.stdout, dt


def main():
    print("(1) K_c(L): two-lattice matching (locate_critical_point), default updates, from K0 = -0.4400, 4096 chains, 8e6 samples per")
    print("    iteration; K per iteration = the last blocking level's estimate, +- jackknife error over 32 groups of chains of the last")
    print("    iteration.  Against the COMPILED reference (16 seeds, tests/golden/critical_point.json) the estimator agrees at 3 sigma per")
    print("    level (tests/test_gpu_dropin.py); the values quoted in the comment of main.cpp:26-29 are single runs of unknown length.")
    for L in (16, 32, 64, 128):
        ks = []
        err = None
        for n_it in (1, 2, 3, 4, 6):
            out, dt = run(["kc", L, -0.4400, n_it, 500, 8000000], dict(MCRG_REPLICAS="4096", MCRG_SEED=str(10 + L)))
            ks.append(float(re.search(r"RESULT Kc (\S+)", out).group(1)))
            err = [float(e) for _, _, e in re.findall(r"RESULT level (\d+) Kc (\S+) err (\S+)", out)][-1]
        print(f"    L={L:4d}: K after 1, 2, 3, 4, 6 iterations = " + ", ".join(f"{k:.6f}" for k in ks) + f"  (+- {err:.6f});  main.cpp comment: {REF_KC[L]:.6f}", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "statistical.json")) as f:
        ref = next(t for t in json.load(f)["lambda"] if t["N"] == 64)
    print("(2) lambda per level at N = 64, K_c: ours (jackknife error) | compiled reference (mean +- error over 16 runs)")
    for name, env, n in (("Metropolis, 32 sweeps per sample", dict(MCRG_SWEEPS_PER_UPDATE="32"), 4000000),
                         ("cluster, 1 update per sample", dict(MCRG_UPDATE="cluster"), 4000000)):
        out, dt = run(["exponent", 64, KC, 1000, n], dict(MCRG_REPLICAS="2048", MCRG_SEED="5", **env))
        res = re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+)", out)
        print(f"    {name} ({dt:.1f} s)")
        for lv, lam, err in res:
            lv = int(lv)
            print(f"      n={lv}: {float(lam):.4f} +- {float(err):.4f} | {ref['mean'][lv]:.4f} +- {ref['err'][lv]:.4f}")
    print("(3) lambda per level at N = 4096, K_c, cluster updates, 64 chains x 400 samples (no reference run exists at this size)")
    out, dt = run(["exponent", 4096, KC, 150, 64 * 400], dict(MCRG_REPLICAS="64", MCRG_UPDATE="cluster", MCRG_SEED="9"))
    for lv, lam, err in re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+)", out):
        print(f"      n={int(lv):2d}: {float(lam):.4f} +- {float(err):.4f}")
    print(f"    ({dt:.1f} s)")


if __name__ == "__main__":
    sys.exit(main())
