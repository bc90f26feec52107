#!/usr/bin/env python
"""bench.py — throughput of the MCRG hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4]     # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...           # the reference's own CPU code, host cores

Workloads = BASELINE.json configs (`--config`, default C4 — the one the metric is quoted on):
  C1  L=8,     K=-0.4406868, 65535 replicas per GPU              (train_scalar_b2_L8 shape; resident kernel)
  C2  N=64,    K_c, 4096 replicas per GPU                        (train_temp_b2_N64 shape; resident kernel)
  C3  L=1024,  K_c, 256 replicas per GPU, 4 RG levels            (full correlator + cross-correlator matrix)
  C4  L=4096,  the 5 couplings of train.cpp:25 x 8 replicas per GPU, all 12 levels        <- headline
  C5  L=16384, K_c, one replica per GPU, 8 RG levels, replicas sharded over the GPUs
In every config one full measurement (correlators at every level of the b=2 pyramid + cross-correlator accumulation)
follows EVERY sweep (m = 1, as the reference measures after every update, mcrg.cpp:75-97).
One "step" = one measurement block: `--samples` x (measure + sweep) on every replica, then the ONE collective of the path —
an int64 all-reduce of the accumulator totals (mcrg.cpp:101-103).
  value = attempts of all ranks / max-over-ranks device time (CUDA events, state resident in HBM).
  e2e   = the same block through the C ABI with HOST buffers: every step takes the replicas' configurations in the reference's
          layout (int32 column-major `imat`, 4 B/spin) from pinned host memory, runs the block, and reads the reduced
          accumulators back.  Forms measured (all listed under e2e.variants, the fastest is e2e.value):
            pipelined_int32     the int32 arrays go through the copy engine under the previous block (_begin/_commit)
            packed_pipelined    the host packs them to 1 bit/spin on its threads (mcrg_host_pack_i32_colmajor) under the
                                previous block, 32x fewer PCIe bytes
            hybrid              part of the replicas each way at once, split by the measured rates (what a host with few
                                cores per GPU needs: the copy engine reads host memory without occupying a core)
          e2e.host_roofline states what the host side can deliver at all (measured streaming-read rate of this rank's
          threads + the copy engine's rate, all ranks at once) and the attempt rate that bound implies.
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
KC = -0.44068679350977147
UNIT = "G spin-flip attempts/s"
# name -> (L, couplings, replicas per coupling per GPU, max_levels, samples per step, description)
CONFIGS = {
    "C1": (8, [-0.4406868], 65535, -1, 512, "C1: L=8, K=-0.4406868, b=2 blocking + MCRG correlators (train_scalar_b2_L8 shape)"),
    "C2": (64, [KC], 4096, -1, 2048, "C2: N=64 at K_c, 4096 independent replicas per GPU (train_temp_b2_N64 shape)"),
    "C3": (1024, [KC], 256, 4, 256, "C3: L=1024, 4 RG levels, 256 replicas per GPU, full correlator + cross-correlator matrix"),
    "C4": (4096, TRAIN_KS, 8, -1, 128, "C4: L=4096 bit-packed, 5 couplings K in [-0.4897,-0.4320] x 8 replicas per GPU"),
    "C5": (16384, [KC], 1, 8, 256, "C5: L=16384, one replica per GPU, 8 RG levels, replica-sharded"),
}
# algorithmic bytes at 1 bit/spin.  The dominant strip kernel (k_sweep0<measure>) reads level 0 once,
# writes level 0 once and writes the level-1 block spins (1/4 bit): 2.25 bits = 0.28125 B per site per launch.
# The whole sample (sweep + full pyramid) is 0.25 + 0.2083 = 0.4583 B per attempt.
BYTES_PER_SITE_DOMINANT = 0.28125
BYTES_PER_ATTEMPT_SAMPLE = 0.25 + (1.0 + 2.0 / 3.0) / 8.0


def metric_name(L):
    return f"spin-flip attempts/s at L={L} with MCRG correlators at every level (m=1)"


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
                "samples": len(s), "sm_mhz_min": s[0] if s else None}


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


def ref_samples_default(L):
    """Reference samples per process per step: about a second of CPU work per process (the loop costs ~0.25 us per site)."""
    return max(1 if L > 4096 else 2, min(20000, int(4e6 / (L * L))))


def cpu_reference_rate(L, n_samples, procs, pool=None):
    """-> (G site-updates/s over all processes, kind, slowest process's loop seconds)."""
    import multiprocessing as mp

    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(procs)
    try:
        res = pool.map(_ref_worker, [(L, KC, n_samples, 1000 + p) for p in range(procs)], chunksize=1)
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

    L = CONFIGS[args.config][0]
    procs = host_procs(L)
    per_step = args.ref_samples or ref_samples_default(L)
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
    line = {"impl": "reference", "metric": metric_name(L), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic (hot start)",
            "config": {"workload": f"{args.config} shape on host cores: L={L}, K=Kc, full pyramid, 1 update per measurement", "L": L},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------------------

def parse_cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def bind_host_threads(local_rank, world):
    """Give this rank its own share of the host cores — on the NUMA node its GPU hangs off when the box says which —
    BEFORE any pinned buffer is allocated (first touch places the pages) and before any packing thread starts (threads
    inherit the mask).  -> (cpus of this rank, how they were chosen)."""
    avail = sorted(os.sched_getaffinity(0))
    if world <= 1:
        return avail, "all host cores (single rank)"
    how = "even slices of the host cores"
    share = None
    try:
        import pynvml as nv

        nv.nvmlInit()
        nodes = []
        for i in range(world):
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(i)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()[-12:]
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                nodes.append(int(f.read()))
        if all(n >= 0 for n in nodes):
            with open(f"/sys/devices/system/node/node{nodes[local_rank]}/cpulist") as f:
                node_cpus = [c for c in parse_cpulist(f.read()) if c in avail]
            peers = [r for r in range(world) if nodes[r] == nodes[local_rank]]
            k = peers.index(local_rank)
            per = len(node_cpus) // len(peers)
            if per >= 1:
                share = node_cpus[k * per:(k + 1) * per]
                how = f"NUMA node {nodes[local_rank]} of the GPU, split among its {len(peers)} ranks"
    except Exception:
        share = None
    if not share:
        per = max(1, len(avail) // world)
        share = avail[(local_rank * per) % len(avail):][:per] or avail
    try:
        os.sched_setaffinity(0, share)
    except OSError:
        how += " (affinity not settable)"
    return share, how


def run_ours(args, rank, world, local_rank):
    cpus, cpu_how = bind_host_threads(local_rank, world)
    import numpy as np
    import torch
    import torch.distributed as dist

    import mcrg_b200
    from mcrg_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L, Ks, per_k, max_lv, S_default, desc = CONFIGS[args.config]
    if args.replicas_per_k:
        per_k = args.replicas_per_k
    S, m = args.samples or S_default, args.sweeps_per_sample
    n_loc = len(Ks) * per_k
    lay = capi.acc_layout()
    ctx = mcrg_b200.Context(L, n_loc, seed=12345, device=local_rank, replica_base=rank * n_loc, n_bins=1)
    ctx.set_couplings(np.repeat(Ks, per_k))
    if args.strip_rows or args.fuse_sweeps != 1 or not args.graphs:
        ctx.set_tuning(args.strip_rows, args.fuse_sweeps, int(args.graphs))
    ctx.init_hot()
    ctx.sweep(10)
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=torch.device("cuda", local_rank))
    limbs = torch.zeros(lay.n_slots * 4, dtype=torch.int64, device="cuda")
    n_lv = capi.levels_full(L) if max_lv < 0 else min(max_lv, capi.levels_full(L))
    resident = L <= 512 and not args.strip_rows
    n_level_kernels = sum(1 for lv in range(1, n_lv + 1) if (L >> lv) > (512 if n_loc >= 16 else 256))  # capi.cu: tail_start_size
    # strips: per sample k_sweep0<measure>, k_level per large level, k_tail, further sweep launches; per step the sweep-counter
    # update(s) (one per graph + one for the rest) and the limb-total kernel.  resident: one launch per block.
    extra_sweeps = 0 if m <= 1 else -(-(m - 1) // max(1, args.fuse_sweeps))
    if resident:
        launches_per_step = 1 + 1 + 1
    else:
        n_graphs = S // 64 + (S % 64) // 16 + (1 if S % 16 else 0)  # capi.cu mcrg_run: 64-sample graphs, 16-sample graphs, the rest
        launches_per_step = S * (2 + n_level_kernels + extra_sweeps) + (n_graphs if args.graphs else 1) + 1

    def block():
        ctx.run(S, m, max_lv, 0)
        ctx.total_limbs_to_device(limbs.data_ptr())
        if world > 1:
            dist.all_reduce(limbs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
        ms = max_over_ranks(e0.elapsed_time(e1))
        attempts_per_step = world * n_loc * L * L * S * m
        value = attempts_per_step * args.steps / (ms * 1e-3) / 1e9
        block_ms = ms / args.steps

        # per-kernel device times, live, same state (events between the launches, no graph); strip kernels only
        prof = None if resident else ctx.profile_kernels(min(S, 32), m, max_lv)
        if prof is not None and m <= 1:
            prof["sweep_only"] = None  # no further sweep launches per sample at m = 1 (the event pair brackets nothing)
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

        n_sw = 32 if L >= 1024 else 256
        other = {"sweep_only": rate(lambda: ctx.sweep(n_sw), n_sw), "m16": rate(lambda: ctx.run(max(2, n_sw // 8), 16, max_lv, 0), max(2, n_sw // 8) * 16)}
        philox_calls_per_s = ctx.probe_philox_rate()  # live: the instruction-issue ceiling of the update's arithmetic core
        barrier()

        # ---- end to end through the C ABI with host buffers -------------------------------------------------------------
        e2e_steps = args.e2e_steps or max(5, args.steps)
        n_thr = max(1, len(cpus))
        host_bytes = n_loc * L * L * 4
        bufs = [torch.empty((n_loc, L, L), dtype=torch.int32).pin_memory() for _ in range(2)]
        chunk = max(4, (64 << 20) // (L * L * 4))
        for r0 in range(0, n_loc, chunk):  # current configurations as the uploaded inputs (valid +-1 data)
            bufs[0].numpy()[r0:r0 + chunk] = ctx.get_spins(r0, min(chunk, n_loc - r0))
        bufs[1].copy_(bufs[0])
        pk = [torch.empty(capi.packed_words(L, n_loc), dtype=torch.int32).pin_memory() for _ in range(2)]
        result = torch.empty(lay.n_slots * 4, dtype=torch.int64).pin_memory()
        per_rep_ints, per_rep_words = L * L, capi.packed_words(L, 1)

        # what the host side can deliver, all ranks at once: streaming read and packing on this rank's threads, and the copy engine
        def host_rate(fn):
            fn()
            best = 1e30
            for _ in range(2):
                barrier()
                t0 = time.perf_counter()
                fn()
                best = min(best, max_over_ranks(time.perf_counter() - t0))
            return host_bytes / best / 1e9

        read_gbs = host_rate(lambda: capi.host_read_probe(bufs[0].data_ptr(), n_loc * L * L, n_thr))
        pack_gbs = host_rate(lambda: capi.host_pack(bufs[0].data_ptr(), L, n_loc, pk[0].data_ptr(), n_thr))
        dev_tmp = torch.empty(min(n_loc, max(1, (1 << 30) // (L * L * 4))) * L * L, dtype=torch.int32, device="cuda")

        def dma():
            dev_tmp.copy_(bufs[0].view(-1)[:dev_tmp.numel()], non_blocking=True)
            stream.synchronize()

        dma_gbs = host_rate(dma) * dev_tmp.numel() * 4 / host_bytes

        # both at once — what the hybrid form asks of the host: the copy engine and the threads read the same memory system, so
        # their rates do not add up on a box whose memory bandwidth is the limit (8 ranks on one host)
        n_dma = dev_tmp.numel()
        n_read = int(min(n_loc * L * L, max(1 << 20, n_dma * read_gbs / max(dma_gbs, 1e-9))))

        def both():
            dev_tmp.copy_(bufs[0].view(-1)[:n_dma], non_blocking=True)
            capi.host_read_probe(bufs[1].data_ptr(), n_read, n_thr)
            stream.synchronize()

        combined_gbs = host_rate(both) * (n_dma + n_read) * 4 / host_bytes
        del dev_tmp

        def e2e_run(n, k_int32):
            """n steps; replicas [0, k_int32) travel as int32 through the copy engine, the rest host-packed; every step's
            inputs are taken from the pinned int32 buffers inside the timed region; the upload of step s+1 overlaps block s."""
            k = k_int32

            def begin(s):
                b = bufs[s & 1]
                if k > 0:
                    ctx.set_spins_begin(b.data_ptr(), k, first=0)
                if k < n_loc:
                    capi.host_pack(b.data_ptr() + k * per_rep_ints * 4, L, n_loc - k, pk[s & 1].data_ptr(), n_thr)
                    ctx.set_spins_packed_begin(pk[s & 1].data_ptr(), n_loc - k, first=k)

            begin(0)
            for s in range(n):
                ctx.set_spins_commit()
                block()
                if s + 1 < n:
                    begin(s + 1)
                result.copy_(limbs, non_blocking=True)
                stream.synchronize()

        def e2e_time(n, k):
            barrier()
            t0 = time.perf_counter()
            e2e_run(n, k)
            barrier()
            return max_over_ranks(time.perf_counter() - t0)

        def e2e_sync_int32(n):
            for _ in range(n):
                ctx.set_spins_ptr(bufs[0].data_ptr(), n_loc)
                block()
                result.copy_(limbs, non_blocking=True)
                stream.synchronize()

        variants = {}
        e2e_run(2, 0)
        t_packed = e2e_time(e2e_steps, 0)
        variants["packed_pipelined"] = attempts_per_step * e2e_steps / t_packed / 1e9
        e2e_run(2, n_loc)
        t_int32 = e2e_time(e2e_steps, n_loc)
        variants["pipelined_int32"] = attempts_per_step * e2e_steps / t_int32 / 1e9
        # hybrid: split by the measured rates, refined by a short search around it (each rank its own split)
        k_star = n_loc * dma_gbs / (dma_gbs + pack_gbs)
        cands = sorted({int(round(k_star * f)) for f in (0.6, 0.8, 1.0, 1.2)} - {0, n_loc})
        cands = [k for k in cands if 0 < k < n_loc]
        best_k, best_t = None, None
        for k in cands:
            e2e_run(1, k)
            barrier()
            t0 = time.perf_counter()
            e2e_run(4, k)
            t = time.perf_counter() - t0  # this rank's own time: the split is per rank
            if best_t is None or t < best_t:
                best_k, best_t = k, t
        if best_k is not None:
            t_hyb = e2e_time(e2e_steps, best_k)
            variants["hybrid"] = attempts_per_step * e2e_steps / t_hyb / 1e9
        barrier()
        t0 = time.perf_counter()
        e2e_sync_int32(min(e2e_steps, 5))
        barrier()
        variants["unpipelined_int32_upload"] = attempts_per_step * min(e2e_steps, 5) / max_over_ranks(time.perf_counter() - t0) / 1e9
        e2e_name = max(variants, key=variants.get)
        e2e_value = variants[e2e_name]
        k_used = {"packed_pipelined": 0, "pipelined_int32": n_loc, "unpipelined_int32_upload": n_loc}.get(e2e_name, best_k)
        e2e_h2d = k_used * per_rep_ints * 4 + (n_loc - k_used) * per_rep_words * 4
        # the host bound: every step the host must deliver host_bytes per rank; its threads read at read_gbs (packing runs
        # at that rate: pack_gbs), the copy engine adds dma_gbs; a step cannot be shorter than the block itself
        host_gbs = max(combined_gbs, read_gbs, dma_gbs)
        t_host = host_bytes / (host_gbs * 1e9)
        host_roofline = attempts_per_step / max(t_host, block_ms * 1e-3) / 1e9
        del bufs, pk

    # the other named configurations, device-resident, a few steps each (same timing rules: warm-up, barrier + max over ranks,
    # CUDA events on the context's stream) so that one default run shows every BASELINE shape; `--config Cn` gives the full line
    others = {}
    if args.config == "C4" and not args.no_other_configs:
        for name in ("C1", "C2", "C3", "C5"):
            oL, oKs, oper, olv, oS, odesc = CONFIGS[name]
            on = len(oKs) * oper
            octx = mcrg_b200.Context(oL, on, seed=12345, device=local_rank, replica_base=rank * on, n_bins=1)
            octx.set_couplings(np.repeat(oKs, oper))
            octx.init_hot()
            octx.sweep(10)
            ostream = torch.cuda.ExternalStream(octx.stream_handle, device=torch.device("cuda", local_rank))
            with torch.cuda.stream(ostream):
                def oblock():
                    octx.run(oS, 1, olv, 0)
                    octx.total_limbs_to_device(limbs.data_ptr())
                    if world > 1:
                        dist.all_reduce(limbs)
                for _ in range(3):
                    oblock()
                barrier()
                o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_steps = 5
                o0.record(ostream)
                for _ in range(n_steps):
                    oblock()
                o1.record(ostream)
                barrier()
                oms = max_over_ranks(o0.elapsed_time(o1))
            others[name] = {"workload": odesc, "value": world * on * oL * oL * oS * n_steps / (oms * 1e-3) / 1e9, "unit": UNIT,
                            "ms_per_step": oms / n_steps, "steps": n_steps, "warmup": 3, "replicas_per_gpu": on, "samples_per_step": oS,
                            "levels": (capi.levels_full(oL) if olv < 0 else min(olv, capi.levels_full(oL))) + 1}
            del o0, o1
            torch.cuda.synchronize()
            octx.close()

    peak, peak_src = measured_peak()
    sites = n_loc * L * L
    if prof is not None:
        dom_ms, dom_name = prof["sweep_measure"], "k_sweep0<MEASURE> (level-0 correlators + block to level 1 + Metropolis sweep)"
        dom_bytes = BYTES_PER_SITE_DOMINANT * sites
        sample_ms = sum(v for v in prof.values() if v)
    else:  # resident kernel: one launch = the whole block of samples, state in shared memory throughout
        dom_ms, dom_name = block_ms, "k_resident<MEASURE> (one launch = S samples: measure + pyramid + accumulate + sweep, state in shared memory)"
        dom_bytes = BYTES_PER_ATTEMPT_SAMPLE * sites * S * m
        sample_ms = block_ms
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.config == "C4":
        try:
            with open(tp) as f:
                traffic = json.load(f).get("k_sweep0_measure_dram_bytes_per_launch")
        except Exception:
            traffic = None
    ws_mib = 2 * n_loc * L * L // 8 >> 20
    line = {
        "metric": metric_name(L), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": block_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 bit-planes (1 bit/spin), int64/int128 sums", "data": "synthetic (Philox hot start, seed 12345, 10 warm-up sweeps)",
        "config": {"workload": f"{desc}, measurement at all {n_lv + 1} levels after every sweep", "name": args.config, "L": L,
                   "replicas_per_gpu": n_loc, "samples_per_step": S, "sweeps_per_sample": m, "levels": n_lv + 1,
                   "parallelism": f"replica-sharded x{world}", "collective": f"one int64 all-reduce of {lay.n_slots * 4} limbs per step",
                   "kernel_path": "resident (whole replica in one CTA's shared memory)" if resident else f"strips of {ctx.strip_plan(1, args.strip_rows)[0]} rows",
                   "l2": f"state is double-buffered: {ws_mib} MiB resident per GPU vs 126 MB L2"
                         + (" (inputs larger than L2)" if 2 * n_loc * L * L // 8 > 126e6 else " (L2-resident by design: 1 bit/spin)"),
                   "cuda_graphs": bool(args.graphs)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": lay.n_slots * 4 * 8,
                "steps": e2e_steps, "input_layout": "int32 column-major imat (reference Lattice::spins_), pinned",
                "host_input_bytes_per_step": host_bytes, "path": e2e_name, "replicas_sent_as_int32": k_used, "variants": variants,
                "pipeline_fill": "the first step's conversion/upload is not overlapped and is inside the timed region",
                "host_threads": n_thr, "host_threads_how": cpu_how,
                "host_roofline": {"value": host_roofline, "unit": UNIT,
                                  "what": "per step every rank must take host_input_bytes_per_step out of host memory: its threads "
                                          "stream at read_GBps (the bit packing runs at pack_GBps), the copy engine alone at dma_GBps, both "
                                          "at once at combined_GBps (they share the host's memory system); all ranks measured at once; "
                                          "value = attempts per step / max(host bytes / best of these rates, device step)",
                                  "read_GBps_per_rank": read_gbs, "pack_GBps_per_rank": pack_gbs, "dma_GBps_per_rank": dma_gbs,
                                  "combined_GBps_per_rank": combined_gbs, "host_GBps_all_ranks": host_gbs * world,
                                  "host_ms_per_step": t_host * 1e3, "device_ms_per_step": block_ms,
                                  "e2e_over_roofline": e2e_value / host_roofline}},
        "gpu_launches": launches_per_step * args.steps,
        "other_schedules_per_gpu": {"unit": UNIT, "sweep_only": other["sweep_only"], "one_measurement_per_16_sweeps": other["m16"]},
        "other_configs": others,
        "roofline": {"bound": "hbm", "kernel": dom_name,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/traffic.json <- profiles/ncu_summary_r2_final.json (one ncu --set full capture; refresh with profiles/run_profiles.sh + summarize.py when the kernel changes)" if traffic else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                     "kernel_ms": dom_ms, "kernel_share_of_sample": dom_ms / sample_ms if sample_ms > 0 else None,
                     "per_sample_ms": prof,
                     "whole_step_frac": value / world * BYTES_PER_ATTEMPT_SAMPLE / peak,
                     "compute_bound": {"what": "Philox4x32-10 + 4-plane lazy compare alone, measured live on this GPU by "
                                               "mcrg_probe_philox_rate (see also profiles/microbench_pipes_r1.txt); a sweep draws 2 calls "
                                               "per 32 sites in pass 1 and ~0.13 in pass 2, so ceiling = calls/s * 32 / 2.1 (the kernel "
                                               "shares the word-independent products of rounds 0-1 between calls, which the probe does not)",
                                       "philox_T_calls_per_s": philox_calls_per_s / 1e12,
                                       "ceiling_G_sites_per_s": philox_calls_per_s / 1e9 * 32 / 2.1,
                                       "achieved_G_sites_per_s": sites * (1 if prof is not None else S * m) / (dom_ms * 1e-3) / 1e9,
                                       "frac": (sites * (1 if prof is not None else S * m) / (dom_ms * 1e-3)) / (philox_calls_per_s * 32 / 2.1)},
                     "note": "1 bit/spin makes the compulsory traffic tiny: the kernel is INT/Philox-issue bound, not HBM bound "
                             "(see DESIGN.md section 3.1 and profiles/)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = host_procs(L)
        n_ref = args.ref_samples or ref_samples_default(L)
        rate_, kind, loop = cpu_reference_rate(L, n_ref, procs)
        line["cpu_baseline"] = {"value": rate_, "unit": UNIT, "cores": procs, "kind": kind,
                                "sample": f"{procs} processes x {n_ref} sample(s) of the reference loop mcrg.cpp:72-98 at L={L} "
                                          f"(Wolff update + correlators at all levels), each sample counted as L^2 attempts; loop {loop:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # release everything that was used on the context's stream before the stream is destroyed
    del result, limbs, e0, e1
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
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--replicas-per-k", type=int, default=0, help="override the config's replicas per coupling per GPU")
    ap.add_argument("--samples", type=int, default=0,
                    help="measurement samples per step = one measurement block between collectives (the reference "
                         "takes 1e4 samples per rank between its all-reduces, main.cpp:10-11 / mcrg.cpp:72-103); 0 = the config's default")
    ap.add_argument("--sweeps-per-sample", type=int, default=1)
    ap.add_argument("--strip-rows", type=int, default=0)
    ap.add_argument("--fuse-sweeps", type=int, default=1)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of every end-to-end variant (default: --steps)")
    ap.add_argument("--ref-samples", type=int, default=0, help="reference samples per process per step (0 = about a second of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short device-resident runs of C1, C2, C3, C5 in the default line")
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
