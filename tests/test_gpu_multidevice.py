"""Several GPUs inside one process through the C ABI (mcrg_comm_init_all / mcrg_allreduce_accumulators: one NCCL
all-reduce of exact integer limbs = the MPI_Allreduce of mcrg.cpp:101-103) and through the drop-in (MCRG_DEVICES).
Results depend only on (seed, global replica id, sweep counter), so any split over devices gives identical integers.
The two-device tests skip on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multidevice.py` runs them."""
import os
import re
import subprocess

import numpy as np
import pytest

import _libs

pytestmark = pytest.mark.gpu

KC = float(-0.5 * np.log(1 + np.sqrt(2)))
APP = os.path.join(_libs.ROOT, "mcrg_b200", "host", "_build", "mcrg_app")


@pytest.fixture(scope="module")
def mc():
    import mcrg_b200

    assert mcrg_b200.capi.device_count() >= 1
    return mcrg_b200


def test_one_context_per_device_is_enforced(mc):
    with mc.Context(16, 2, seed=1) as a, mc.Context(16, 2, seed=1, replica_base=2) as b:
        with pytest.raises(mc.capi.McrgError, match="share device"):
            mc.capi.comm_init_all([a, b])
        with pytest.raises(mc.capi.McrgError, match="no communicator"):
            mc.capi.allreduce_accumulators([a])


def test_single_device_group_reduces_to_its_own_totals(mc):
    """n = 1 is a valid group (a communicator of one rank): the all-reduce returns the context's own totals."""
    with mc.Context(64, 6, seed=5) as ctx:
        ctx.init_hot()
        ctx.sweep(3)
        ctx.run(7, 1, -1, 0)
        mc.capi.comm_init_all([ctx])
        tot = mc.capi.allreduce_accumulators([ctx])
        acc, _ = ctx.accumulators()
        assert tot == [int(x) for x in acc.sum(axis=(0, 1))]
        mc.capi.comm_destroy_all([ctx])


@pytest.mark.parametrize("L,strip", [(64, 0), (256, 32)])
def test_two_devices_give_the_totals_of_one(mc, L, strip):
    if mc.capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    seed, n_samples = 17, 11
    Ks = [KC, -0.43, -0.45, -0.47, -0.44, -0.42]

    def prepare(ctx, first, count):
        if strip:
            ctx.set_tuning(strip_rows=strip)
        ctx.set_couplings(Ks[first:first + count])
        ctx.init_hot()
        ctx.sweep(4)
        ctx.run(n_samples, 2, -1, 0)

    with mc.Context(L, 6, seed=seed) as whole:
        prepare(whole, 0, 6)
        acc, _ = whole.accumulators()
        want = [int(x) for x in acc.sum(axis=(0, 1))]
    with mc.Context(L, 2, seed=seed, device=0, replica_base=0) as a, mc.Context(L, 4, seed=seed, device=1, replica_base=2) as b:
        mc.capi.comm_init_all([a, b])
        prepare(a, 0, 2)  # asynchronous: both devices work at the same time
        prepare(b, 2, 4)
        got = mc.capi.allreduce_accumulators([a, b])
        assert got == want
        b.run(3, 1, -1, 0)  # the group stays usable
        again = mc.capi.allreduce_accumulators([a, b])
        lay = mc.capi.acc_layout()
        assert again[lay.slot_n] == want[lay.slot_n] + 4 * 3
        mc.capi.comm_destroy_all([a, b])


def test_dropin_driver_on_two_devices_prints_the_same_result(tmp_path):
    import mcrg_b200

    if mcrg_b200.capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    for n_dev in ("1", "2"):
        d = tmp_path / n_dev
        d.mkdir()
        env = dict(os.environ, MCRG_REPLICAS="256", MCRG_SWEEPS_PER_UPDATE="4", MCRG_SEED="99", MCRG_QUIET="1", MCRG_DEVICES=n_dev)
        out = subprocess.run([APP, "exponent", "32", repr(KC), "200", "200000"], cwd=d, env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        outs.append(re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+) nu (\S+)", out.stdout))
    assert len(outs[0]) == 4 and outs[0] == outs[1]


def test_dropin_totals_beyond_64_bits_on_two_devices(tmp_path):
    """N = 4096 with 64 chains x 400 samples: the slot totals (sum S_nn^2 ~ 1.4e15 per sample) pass 2^64, where a running
    long-double sum of the per-chain values no longer equals the exact all-reduced total.  The drop-in compares the two in exact
    integers, so the two-device run must go through (round 1 threw "all-reduced totals differ") and print what one device prints."""
    import mcrg_b200

    if mcrg_b200.capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    for n_dev in ("1", "2"):
        d = tmp_path / n_dev
        d.mkdir()
        env = dict(os.environ, MCRG_REPLICAS="64", MCRG_SEED="5", MCRG_QUIET="1", MCRG_DEVICES=n_dev, MCRG_UPDATE="metropolis")
        out = subprocess.run([APP, "exponent", "4096", repr(KC), "50", str(64 * 400)], cwd=d, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        outs.append(re.findall(r"RESULT level (\d+) lambda (\S+) err (\S+) nu (\S+)", out.stdout))
    assert len(outs[0]) == 11 and outs[0] == outs[1]
