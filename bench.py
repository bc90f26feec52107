#!/usr/bin/env python
"""bench.py — throughput of the MCRG hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference's own CPU code, host cores

Workload (BASELINE.json metric: spin-flip attempts/s at L=4096 with MCRG correlators; config C4):
  L = 4096 bit-packed lattices, the 5 couplings of train.cpp:25 x `--replicas-per-k` replicas on every GPU,
  one full measurement (correlators at all 12 levels of the b=2 pyramid + cross-correlator accumulation) after
  EVERY sweep (m = 1, as the reference measures after every update, mcrg.cpp:75-97).
  One "step" = one measurement block: `--samples` (default 128) x (measure + sweep) on every replica, then the ONE collective
  of the path — an int64 all-reduce of the accumulator totals (mcrg.cpp:101-103).
  value = attempts of all ranks / max-over-ranks device time (CUDA events, state resident in HBM).
  e2e   = the same block through the C ABI with HOST buffers: every step uploads the replicas' configurations in
          the reference's layout (int32 column-major `imat`, 4 B/spin) from pinned memory, runs the block, and
          reads the reduced accumulators back.  Reported pipelined (the upload of the next step's configurations runs on
          a copy stream while the current block computes; 2.7 GB per step make it PCIe-bound), unpipelined, and with
          the configurations bit-packed on the host first (32x fewer PCIe bytes); `value` is the fastest, all are listed.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_KS = [-0.4320459, -0.4406868, -0.4496804, -0.4688157, -0.489652]  # train.cpp:25 / checked-in log names
UNIT = "G spin-flip attempts/s"
METRIC = "spin-flip attempts/s at L=4096 with MCRG correlators at every level (m=1)"
# algorithmic bytes at 1 bit/spin.  The dominant kernel (k_sweep0<measure>) reads level 0 once,
# writes level 0 once and writes the level-1 block spins (1/4 bit): 2.25 bits = 0.28125 B per site per launch.
# The whole sample (sweep + full pyramid) is 0.25 + 0.2083 = 0.4583 B per attempt.
BYTES_PER_SITE_DOMINANT = 0.28125
BYTES_PER_ATTEMPT_SAMPLE = 0.25 + (1.0 + 2.0 / 3.0) / 8.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that rather than inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# -------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own loop (mcrg.cpp:72-98) on the host cores
# -------------------------------------------------------------------------------------------------------------

def _ref_worker(args):
    L, K, n_samples, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs

    kind = "reference" if _libs.ref_available() else "port"
    if kind == "reference":
        r = _libs.ref()
        r.ref_seed(seed)
        secs = r.ref_mcrg_loop(0, n_samples, L, K, 0, None, None)
    else:  # the plain-C oracle: scalar Metropolis sweep + full pyramid measurement per sample
        import numpy as np

        o = _libs.oracle()
        s = np.zeros((L, L), np.int32)
        o.orc_hot_start(L, seed, 0, s)
        t0 = time.perf_counter()
        for k in range(n_samples):
            o.orc_metropolis(L, s, K, seed, 0, k, 1)
            _libs.pyramid(L, s, seed, 0, k)
        secs = time.perf_counter() - t0
    return secs, kind


def host_procs(L):
    n = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
        per_proc = 6 * 4 * L * L + (64 << 20)  # the reference copies the imat several times per blocking level
        n = max(1, min(n, int(avail_kb * 1024 * 0.5 / per_proc)))
    except Exception:
        pass
    return n


def cpu_reference_rate(L, n_samples, procs, pool=None):
    """-> (G site-updates/s over all processes, kind, slowest process's loop seconds)."""
    import multiprocessing as mp

    Kc = TRAIN_KS[1]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(procs)
    try:
        res = pool.map(_ref_worker, [(L, Kc, n_samples, 1000 + p) for p in range(procs)], chunksize=1)
    finally:
        if own:
            pool.close()
            pool.join()
    loop = max(r[0] for r in res)
    return procs * n_samples * L * L / loop / 1e9, res[0][1], loop


def run_reference(args, rank, world):
    """The reference arm: rank 0 alone times the reference's CPU code with all host threads; other ranks exit."""
    if rank != 0:
        return
    import math
    import multiprocessing as mp

    L = args.L
    procs = host_procs(L)
    per_step = args.ref_samples
    kind = "reference"
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        for _ in range(min(args.warmup, 1)):  # one warm-up pass is enough to page the library in on every worker
            cpu_reference_rate(L, 1, procs, pool)
        for _ in range(args.steps):
            _, kind, loop = cpu_reference_rate(L, per_step, procs, pool)
            times.append(loop)
    total = sum(times)
    value = procs * per_step * args.steps * L * L / total / 1e9
    sample = (f"{procs} independent processes (the reference's own parallel model, mcrg.cpp:42-50) x {per_step} sample(s)/step "
              f"of the loop mcrg.cpp:72-98 at L={L}: one Wolff cluster update (ising.cpp:87-155; the reference has no "
              f"Metropolis) + calc_interactions at all {int(math.log2(L))} levels; each sample is counted as L^2 attempts")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic (hot start)",
            "config": {"workload": f"C4 shape on host cores: L={L}, K=Kc, full pyramid, 1 update per measurement", "L": L},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mcrg_b200
    from mcrg_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L, S, m = args.L, args.samples, args.sweeps_per_sample
    n_loc = len(TRAIN_KS) * args.replicas_per_k
    lay = capi.acc_layout()
    ctx = mcrg_b200.Context(L, n_loc, seed=12345, device=local_rank, replica_base=rank * n_loc, n_bins=1)
    ctx.set_couplings(np.repeat(TRAIN_KS, args.replicas_per_k))
    if args.strip_rows or args.fuse_sweeps != 1 or not args.graphs:
        ctx.set_tuning(args.strip_rows, args.fuse_sweeps, int(args.graphs))
    ctx.init_hot()
    ctx.sweep(10)
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=torch.device("cuda", local_rank))
    limbs = torch.zeros(lay.n_slots * 4, dtype=torch.int64, device="cuda")
    n_lv = capi.levels_full(L)
    n_level_kernels = sum(1 for lv in range(1, n_lv + 1) if (L >> lv) > 256)
    # per sample: k_sweep0<measure>, k_level per large level, k_tail, further sweep launches; per step: the sweep
    # counter update(s) (one per 16-sample graph + one for the rest) and the limb-total kernel
    extra_sweeps = 0 if m <= 1 else -(-(m - 1) // max(1, args.fuse_sweeps))
    launches_per_step = S * (2 + n_level_kernels + extra_sweeps) + ((S // 16 + (1 if S % 16 else 0)) if args.graphs else 1) + 1

    def block():
        ctx.run(S, m, -1, 0)
        ctx.total_limbs_to_device(limbs.data_ptr())
        if world > 1:
            dist.all_reduce(limbs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            block()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            block()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        attempts_per_step = world * n_loc * L * L * S * m
        value = attempts_per_step * args.steps / (ms * 1e-3) / 1e9

        # per-kernel device times, live, same state (events between the launches, no graph)
        prof = ctx.profile_kernels(min(S, 32), m, -1)
        barrier()

        # also the sweep-only rate and one measurement every 16 sweeps (same state, device timers, this rank)
        def rate(fn, sweeps):
            fn()
            ctx.sync()
            best = 1e30
            for _ in range(3):
                ctx.timer_start()
                fn()
                best = min(best, ctx.timer_stop())
            return n_loc * L * L * sweeps / (best * 1e-3) / 1e9

        other = {"sweep_only": rate(lambda: ctx.sweep(32), 32), "m16": rate(lambda: ctx.run(4, 16, -1, 0), 64)}
        philox_calls_per_s = ctx.probe_philox_rate()  # live: the instruction-issue ceiling of the update's arithmetic core
        barrier()

        # ---- end to end through the C ABI with host buffers
        e2e_steps = max(5, args.steps // 2)
        host = torch.empty((n_loc, L, L), dtype=torch.int32).pin_memory()
        host_np = host.numpy()
        for r0 in range(0, n_loc, 4):  # current configurations as the uploaded inputs (valid +-1 data)
            host_np[r0:r0 + 4] = ctx.get_spins(r0, min(4, n_loc - r0))
        result = torch.empty(lay.n_slots * 4, dtype=torch.int64).pin_memory()

        def e2e_block():
            ctx.set_spins_ptr(host.data_ptr(), n_loc)
            block()
            result.copy_(limbs, non_blocking=True)
            stream.synchronize()

        e2e_block()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_block()
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_sync_value = attempts_per_step * e2e_steps / float(e2e_s.item()) / 1e9

        # the same, pipelined: the upload of step s+1 (copy stream, second pinned buffer) overlaps the kernels of step s
        # (mcrg_set_spins_i32_colmajor_begin / mcrg_set_spins_commit); every step's upload is inside the timed region
        host2 = torch.empty((n_loc, L, L), dtype=torch.int32).pin_memory()
        host2.copy_(host)
        bufs = [host, host2]

        def e2e_pipelined(n):
            ctx.set_spins_begin(bufs[0].data_ptr(), n_loc)
            for s in range(n):
                ctx.set_spins_commit()
                if s + 1 < n:
                    ctx.set_spins_begin(bufs[(s + 1) & 1].data_ptr(), n_loc)
                block()
                result.copy_(limbs, non_blocking=True)
                stream.synchronize()

        e2e_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined(e2e_steps)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_pipe_value = attempts_per_step * e2e_steps / float(e2e_s.item()) / 1e9

        # third form: pack on the host (1 bit/spin, all host threads) while the previous block runs on the GPU, upload
        # the packed words (32x fewer PCIe bytes).  Same host inputs (int32 imat in pinned memory), same results.
        n_thr = max(1, (os.cpu_count() or 1) // world)  # ranks share the host cores
        pk = [torch.empty(capi.packed_words(L, n_loc), dtype=torch.int32).pin_memory() for _ in range(2)]

        def e2e_packed(n):
            capi.host_pack(bufs[0].data_ptr(), L, n_loc, pk[0].data_ptr(), n_thr)
            for s in range(n):
                ctx.set_spins_packed_ptr(pk[s & 1].data_ptr(), n_loc)
                block()
                if s + 1 < n:  # the GPU is busy with block s: pack the next step's configurations meanwhile
                    capi.host_pack(bufs[(s + 1) & 1].data_ptr(), L, n_loc, pk[(s + 1) & 1].data_ptr(), n_thr)
                result.copy_(limbs, non_blocking=True)
                stream.synchronize()

        e2e_packed(2)
        barrier()
        t0 = time.perf_counter()
        e2e_packed(e2e_steps)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_packed_value = attempts_per_step * e2e_steps / float(e2e_s.item()) / 1e9
        # fourth form: as the third, with the packed upload of step s+1 on the copy stream under block s (begin / commit)
        def e2e_packed_pipelined(n):
            capi.host_pack(bufs[0].data_ptr(), L, n_loc, pk[0].data_ptr(), n_thr)
            ctx.set_spins_packed_begin(pk[0].data_ptr(), n_loc)
            for s in range(n):
                ctx.set_spins_commit()
                block()
                if s + 1 < n:
                    capi.host_pack(bufs[(s + 1) & 1].data_ptr(), L, n_loc, pk[(s + 1) & 1].data_ptr(), n_thr)
                    ctx.set_spins_packed_begin(pk[(s + 1) & 1].data_ptr(), n_loc)
                result.copy_(limbs, non_blocking=True)
                stream.synchronize()

        e2e_packed_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        e2e_packed_pipelined(e2e_steps)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_packed_pipe_value = attempts_per_step * e2e_steps / float(e2e_s.item()) / 1e9
        e2e_packed_sync_value = e2e_packed_value
        e2e_packed_value = max(e2e_packed_value, e2e_packed_pipe_value)
        use_packed = e2e_packed_value > e2e_pipe_value
        e2e_value = max(e2e_packed_value, e2e_pipe_value)
        e2e_h2d = pk[0].numel() * 4 if use_packed else n_loc * L * L * 4
        del host2, bufs, pk

    peak, peak_src = measured_peak()
    dom_ms = prof["sweep_measure"]
    sample_ms = sum(prof.values())
    achieved = BYTES_PER_SITE_DOMINANT * n_loc * L * L / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("k_sweep0_measure_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 bit-planes (1 bit/spin), int64/int128 sums", "data": "synthetic (Philox hot start, seed 12345, 10 warm-up sweeps)",
        "config": {"workload": f"C4: L={L} bit-packed, 5 couplings K in [-0.4897,-0.4320] x {args.replicas_per_k} replicas per GPU, "
                               f"measurement at all {n_lv + 1} levels after every sweep", "L": L, "replicas_per_gpu": n_loc,
                   "samples_per_step": S, "sweeps_per_sample": m, "levels": n_lv + 1, "parallelism": f"replica-sharded x{world}",
                   "collective": f"one int64 all-reduce of {lay.n_slots * 4} limbs per step",
                   "l2": f"state is double-buffered: {2 * n_loc * L * L // 8 >> 20} MiB resident per GPU vs 126 MB L2"
                         + (" (inputs larger than L2)" if 2 * n_loc * L * L // 8 > 126e6 else " (L2-resident by design: 1 bit/spin)"),
                   "cuda_graphs": bool(args.graphs)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": lay.n_slots * 4 * 8,
                "steps": e2e_steps, "input_layout": "int32 column-major imat (reference Lattice::spins_), pinned",
                "host_input_bytes_per_step": n_loc * L * L * 4,
                "path": ("host-side bit packing on %d threads (mcrg_host_pack_i32_colmajor) overlapped with the previous "
                         "block, packed upload" % n_thr) if use_packed else
                        "int32 upload of step s+1 on a copy stream overlaps the kernels of step s (_begin/_commit)",
                "variants": {"unpipelined_int32_upload": e2e_sync_value, "pipelined_int32_upload": e2e_pipe_value,
                             "host_packed_upload": e2e_packed_sync_value, "host_packed_pipelined_upload": e2e_packed_pipe_value}},
        "gpu_launches": launches_per_step * args.steps,
        "other_schedules_per_gpu": {"unit": UNIT, "sweep_only": other["sweep_only"], "one_measurement_per_16_sweeps": other["m16"]},
        "roofline": {"bound": "hbm", "kernel": "k_sweep0<MEASURE> (level-0 correlators + block to level 1 + Metropolis sweep)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_SITE_DOMINANT * n_loc * L * L,
                     "kernel_ms": dom_ms, "kernel_share_of_sample": dom_ms / sample_ms if sample_ms > 0 else None,
                     "per_sample_ms": prof,
                     "compute_bound": {"what": "Philox4x32-10 + 4-plane lazy compare alone, measured live on this GPU by "
                                               "mcrg_probe_philox_rate (see also profiles/microbench_pipes_r1.txt); a sweep draws 2 calls "
                                               "per 32 sites in pass 1 and ~0.13 in pass 2, so ceiling = calls/s * 32 / 2.1 (the kernel "
                                               "shares the word-independent products of rounds 0-1 between calls, which the probe does not)",
                                       "philox_T_calls_per_s": philox_calls_per_s / 1e12,
                                       "ceiling_G_sites_per_s": philox_calls_per_s / 1e9 * 32 / 2.1,
                                       "achieved_G_sites_per_s": n_loc * L * L / (dom_ms * 1e-3) / 1e9,
                                       "frac": (n_loc * L * L / (dom_ms * 1e-3)) / (philox_calls_per_s * 32 / 2.1)},
                     "note": "1 bit/spin makes the compulsory traffic tiny: the kernel is INT/Philox-issue bound, not HBM bound "
                             "(see DESIGN.md section 3.1 and profiles/)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = host_procs(L)
        rate, kind, loop = cpu_reference_rate(L, args.ref_samples, procs)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": procs, "kind": kind,
                                "sample": f"{procs} processes x {args.ref_samples} sample(s) of the reference loop mcrg.cpp:72-98 at L={L} "
                                          f"(Wolff update + correlators at all levels), each sample counted as L^2 attempts; loop {loop:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # release everything that was used on the context's stream before the stream is destroyed
    del host, host_np, result, limbs, e0, e1
    torch.cuda.synchronize()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--replicas-per-k", type=int, default=8)
    ap.add_argument("--samples", type=int, default=128,
                    help="measurement samples per step = one measurement block between collectives (the reference "
                         "takes 1e4 samples per rank between its all-reduces, main.cpp:10-11 / mcrg.cpp:72-103)")
    ap.add_argument("--sweeps-per-sample", type=int, default=1)
    ap.add_argument("--strip-rows", type=int, default=0)
    ap.add_argument("--fuse-sweeps", type=int, default=1)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--ref-samples", type=int, default=2, help="reference samples per process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
