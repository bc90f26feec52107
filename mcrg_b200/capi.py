"""ctypes binding of include/mcrg_b200.h.  Thin: every method is one C call; no compute happens here.

There is no CPU path: if libmcrg_b200.so is missing, or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCRG_LIB") or os.path.join(HERE, "libmcrg_b200.so")  # MCRG_LIB: A/B builds of the same library

MAX_LEVELS = 15
NOP = 3
NOBS = 4


class McrgError(RuntimeError):
    pass


class AccLayout(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_slots", "slot_n", "slot_absm", "slot_m2", "slot_s", "slot_ss",
                                       "slot_sbs", "slot_sb0", "slot_m4")]
    M4_SPLIT_BITS = 20
    dslot_m4 = 0  # index of sum M^4 in the float64 companion array Context.accumulators() derives from the exact slots

    def m4(self, slots):
        """Exact sum of M^4 from the three slots at slot_m4 (sum h*h, h*l, l*l with M^2 = h*2^20 + l), as a Python int."""
        hh, hl, ll = (int(slots[self.slot_m4 + k]) for k in range(3))
        return (hh << (2 * self.M4_SPLIT_BITS)) + (hl << (self.M4_SPLIT_BITS + 1)) + ll


_lib = None


def lib():
    """Loads the CUDA library (never builds it, never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise McrgError(f"{LIB_PATH} is missing: run `python -m mcrg_b200.build` (nvcc, sm_100a); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        P = C.POINTER
        vp = C.c_void_p
        l.mcrg_last_error.restype = C.c_char_p
        l.mcrg_device_count.argtypes = [P(C.c_int)]
        l.mcrg_version.restype = C.c_int
        l.mcrg_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int, P(vp)]
        l.mcrg_ctx_destroy.argtypes = [vp]
        l.mcrg_sync.argtypes = [vp]
        l.mcrg_stream_handle.argtypes = [vp]
        l.mcrg_stream_handle.restype = C.c_uint64
        l.mcrg_timer_start.argtypes = [vp]
        l.mcrg_timer_stop.argtypes = [vp, P(C.c_float)]
        l.mcrg_levels_full.argtypes = [C.c_int]
        l.mcrg_set_tuning.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        l.mcrg_set_update.argtypes = [vp, C.c_int]
        l.mcrg_strip_plan.argtypes = [vp, C.c_int, C.c_int, P(C.c_int), P(C.c_double)]
        l.mcrg_set_couplings.argtypes = [vp, vp, C.c_int]
        l.mcrg_init_hot.argtypes = [vp]
        l.mcrg_init_cold.argtypes = [vp]
        l.mcrg_set_spins_i32_colmajor.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_get_spins_i32_colmajor.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_set_spins_i32_colmajor_begin.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_set_spins_commit.argtypes = [vp]
        l.mcrg_set_spins_packed_begin.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_packed_words.argtypes = [C.c_int, C.c_int]
        l.mcrg_packed_words.restype = C.c_size_t
        l.mcrg_host_pack_i32_colmajor.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        l.mcrg_set_spins_packed.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_host_read_probe.argtypes = [vp, C.c_size_t, C.c_int, P(C.c_int64)]
        l.mcrg_get_level_spins_i32_colmajor.argtypes = [vp, C.c_int, C.c_int, vp]
        l.mcrg_get_sweep_counter.argtypes = [vp, P(C.c_uint64)]
        l.mcrg_set_sweep_counter.argtypes = [vp, C.c_uint64]
        l.mcrg_sweep.argtypes = [vp, C.c_int]
        l.mcrg_measure.argtypes = [vp, C.c_int, vp, P(C.c_int)]
        l.mcrg_observables.argtypes = [vp, vp, vp, vp, vp]
        l.mcrg_tie_words.argtypes = [C.c_int, C.c_int]
        l.mcrg_tie_words.restype = C.c_size_t
        l.mcrg_measure_supplied.argtypes = [vp, C.c_int, vp, vp, P(C.c_int)]
        l.mcrg_run.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.mcrg_profile_kernels.argtypes = [vp, C.c_int, C.c_int, C.c_int, P(C.c_float)]
        l.mcrg_probe_philox_rate.argtypes = [vp, P(C.c_double)]
        l.mcrg_rgnn_set_weights.argtypes = [vp, vp]
        l.mcrg_rgnn_eval.argtypes = [vp, C.c_double, vp, vp]
        l.mcrg_rgnn_run.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        l.mcrg_rgnn_accumulators_reset.argtypes = [vp]
        l.mcrg_rgnn_accumulators_get.argtypes = [vp, vp]
        l.mcrg_accumulators_layout.argtypes = [P(AccLayout)]
        l.mcrg_accumulators_reset.argtypes = [vp]
        l.mcrg_accumulators_get.argtypes = [vp, vp, vp]
        l.mcrg_accumulators_total_limbs_device.argtypes = [vp, vp]
        l.mcrg_comm_init_all.argtypes = [C.c_int, P(vp)]
        l.mcrg_allreduce_accumulators.argtypes = [C.c_int, P(vp), vp, vp]
        l.mcrg_comm_destroy_all.argtypes = [C.c_int, P(vp)]
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise McrgError(f"mcrg_b200 error {rc}: {lib().mcrg_last_error().decode()}")


def device_count():
    n = C.c_int(0)
    _check(lib().mcrg_device_count(C.byref(n)))
    return n.value


def acc_layout():
    lay = AccLayout()
    _check(lib().mcrg_accumulators_layout(C.byref(lay)))
    return lay


def levels_full(L):
    return lib().mcrg_levels_full(L)


def packed_words(L, count):
    return lib().mcrg_packed_words(L, count)


def host_pack(spins_ptr, L, count, packed_ptr, n_threads=1):
    """int32 column-major host configurations -> 1 bit/spin transport words, on the host (mcrg_host_pack_i32_colmajor)."""
    _check(lib().mcrg_host_pack_i32_colmajor(C.c_void_p(spins_ptr), L, count, C.c_void_p(packed_ptr), n_threads))


def host_read_probe(ptr, n_ints, n_threads=1):
    """Streams n_ints int32 through n_threads host threads (no conversion); returns their sum.  Time it from outside."""
    s = C.c_int64(0)
    _check(lib().mcrg_host_read_probe(C.c_void_p(ptr), n_ints, n_threads, C.byref(s)))
    return s.value


def _handles(ctxs):
    return (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])


def comm_init_all(ctxs):
    """One NCCL communicator over the contexts (one per device) of THIS process: mcrg_comm_init_all."""
    _check(lib().mcrg_comm_init_all(len(ctxs), _handles(ctxs)))


def allreduce_accumulators(ctxs):
    """-> exact Python ints [n_slots]: totals over all replicas and bins of all contexts (one ncclAllReduce per device)."""
    lay = acc_layout()
    hi = np.zeros(lay.n_slots, np.int64)
    lo = np.zeros(lay.n_slots, np.uint64)
    _check(lib().mcrg_allreduce_accumulators(len(ctxs), _handles(ctxs), hi.ctypes.data, lo.ctypes.data))
    return [int(h) * (1 << 64) + int(l) for h, l in zip(hi, lo)]


def comm_destroy_all(ctxs):
    _check(lib().mcrg_comm_destroy_all(len(ctxs), _handles(ctxs)))


class Context:
    """A batch of `n_replicas` L x L lattices on one device (mcrg_ctx)."""

    def __init__(self, L, n_replicas, seed=12345, device=0, replica_base=0, n_bins=1):
        self._h = C.c_void_p()
        self.L, self.n_replicas, self.n_bins, self.device = L, n_replicas, n_bins, device
        rc = lib().mcrg_ctx_create(device, L, n_replicas, seed, replica_base, n_bins, C.byref(self._h))
        if rc != 0:
            msg = lib().mcrg_last_error().decode()
            if self._h:
                lib().mcrg_ctx_destroy(self._h)
                self._h = C.c_void_p()
            raise McrgError(f"mcrg_ctx_create failed ({rc}): {msg}")

    def close(self):
        if self._h:
            lib().mcrg_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state
    def sync(self):
        _check(lib().mcrg_sync(self._h))

    @property
    def stream_handle(self):
        return lib().mcrg_stream_handle(self._h)

    def timer_start(self):
        _check(lib().mcrg_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        _check(lib().mcrg_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_tuning(self, strip_rows=0, fuse_sweeps=1, use_graphs=1):
        _check(lib().mcrg_set_tuning(self._h, strip_rows, fuse_sweeps, use_graphs))

    def strip_plan(self, fused_sweeps=1, strip_rows=0):
        """-> (rows per strip, estimated launch cost): the library's own choice for strip_rows = 0, else that height's estimate."""
        r, c = C.c_int(0), C.c_double(0)
        _check(lib().mcrg_strip_plan(self._h, fused_sweeps, strip_rows, C.byref(r), C.byref(c)))
        return r.value, c.value

    def set_update(self, mode):
        """'metropolis' (default) or 'cluster' (Swendsen-Wang); applies to sweep(), run() and rgnn_run()."""
        _check(lib().mcrg_set_update(self._h, {"metropolis": 0, "cluster": 1}[mode]))

    def set_couplings(self, K):
        K = np.ascontiguousarray(np.atleast_1d(np.asarray(K, np.float64)))
        _check(lib().mcrg_set_couplings(self._h, K.ctypes.data, K.size))

    def init_hot(self):
        _check(lib().mcrg_init_hot(self._h))

    def init_cold(self):
        _check(lib().mcrg_init_cold(self._h))

    def set_spins(self, spins, first=0):
        """spins: [count, L, L] int32 in the reference layout (arr[r, j, i])."""
        spins = np.ascontiguousarray(spins, np.int32).reshape(-1, self.L, self.L)
        _check(lib().mcrg_set_spins_i32_colmajor(self._h, first, spins.shape[0], spins.ctypes.data))

    def set_spins_ptr(self, host_ptr, count, first=0):
        """Same, from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        _check(lib().mcrg_set_spins_i32_colmajor(self._h, first, count, C.c_void_p(host_ptr)))

    def set_spins_begin(self, host_ptr, count, first=0):
        """Start an asynchronous upload from PINNED host memory (copy stream); pair with set_spins_commit()."""
        _check(lib().mcrg_set_spins_i32_colmajor_begin(self._h, first, count, C.c_void_p(host_ptr)))

    def set_spins_packed_begin(self, packed_ptr, count, first=0):
        """Start an asynchronous upload of host-packed configurations from PINNED memory; pair with set_spins_commit()."""
        _check(lib().mcrg_set_spins_packed_begin(self._h, first, count, C.c_void_p(packed_ptr)))

    def set_spins_commit(self):
        _check(lib().mcrg_set_spins_commit(self._h))

    def set_spins_packed_ptr(self, packed_ptr, count, first=0):
        """Upload host-packed configurations (host_pack); stream-ordered, asynchronous for pinned memory."""
        _check(lib().mcrg_set_spins_packed(self._h, first, count, C.c_void_p(packed_ptr)))

    def get_spins(self, first=0, count=None):
        count = self.n_replicas - first if count is None else count
        out = np.empty((count, self.L, self.L), np.int32)
        _check(lib().mcrg_get_spins_i32_colmajor(self._h, first, count, out.ctypes.data))
        return out

    def get_level_spins(self, replica, level):
        n = self.L >> level
        out = np.empty((n, n), np.int32)
        _check(lib().mcrg_get_level_spins_i32_colmajor(self._h, replica, level, out.ctypes.data))
        return out

    @property
    def sweep_counter(self):
        t = C.c_uint64(0)
        _check(lib().mcrg_get_sweep_counter(self._h, C.byref(t)))
        return t.value

    @sweep_counter.setter
    def sweep_counter(self, t):
        _check(lib().mcrg_set_sweep_counter(self._h, t))

    # ---- hot path
    def sweep(self, n):
        _check(lib().mcrg_sweep(self._h, n))

    def measure(self, max_levels=-1):
        """-> S[replica, level, {nn, nnn, plaq, sum}] of the current configurations."""
        n_lv = min(levels_full(self.L), max_levels) if max_levels >= 0 else levels_full(self.L)
        S = np.zeros((self.n_replicas, n_lv + 1, NOBS), np.int64)
        got = C.c_int(0)
        _check(lib().mcrg_measure(self._h, max_levels, S.ctypes.data, C.byref(got)))
        assert got.value == n_lv
        return S

    def measure_supplied(self, tie_levels, max_levels=-1):
        """measure() with caller-supplied tie coins.  tie_levels[lv - 1]: int array [replica, Ln, Ln] (reference layout, arr[r, j, i]) of
        +-1 (or 0/1) coins for the blocks of level lv = 1 .. n_lv; only the entries of tied blocks matter."""
        n_lv = min(levels_full(self.L), max_levels) if max_levels >= 0 else levels_full(self.L)
        per = lib().mcrg_tie_words(self.L, n_lv)
        bits = np.zeros((self.n_replicas, max(per, 1)), np.uint32)
        off = 0
        for lv in range(1, n_lv + 1):
            Ln = self.L >> lv
            Wn = max(1, Ln // 32)
            t = (np.asarray(tie_levels[lv - 1]).reshape(self.n_replicas, Ln, Ln) > 0)
            for w in range(Wn):
                chunk = t[:, :, 32 * w:32 * w + min(32, Ln)].astype(np.uint64)
                words = (chunk << np.arange(chunk.shape[2], dtype=np.uint64)).sum(axis=2).astype(np.uint32)  # [replica, row]
                bits[:, off + np.arange(Ln) * Wn + w] = words
            off += Ln * Wn
        S = np.zeros((self.n_replicas, n_lv + 1, NOBS), np.int64)
        got = C.c_int(0)
        _check(lib().mcrg_measure_supplied(self._h, max_levels, bits.ctypes.data, S.ctypes.data, C.byref(got)))
        assert got.value == n_lv
        return S

    def observables(self):
        out = [np.zeros(self.n_replicas, np.int64) for _ in range(4)]
        _check(lib().mcrg_observables(self._h, *[o.ctypes.data for o in out]))
        return dict(Snn=out[0], Snnn=out[1], Splaq=out[2], M=out[3])

    def run(self, n_samples, sweeps_per_sample=1, max_levels=-1, bin=0):
        _check(lib().mcrg_run(self._h, n_samples, sweeps_per_sample, max_levels, bin))

    def profile_kernels(self, n_samples, sweeps_per_sample=1, max_levels=-1):
        """-> dict of average device ms per sample for each kernel class (events between the launches)."""
        out = (C.c_float * 4)()
        _check(lib().mcrg_profile_kernels(self._h, n_samples, sweeps_per_sample, max_levels, out))
        return dict(sweep_measure=out[0], sweep_only=out[1], level=out[2], tail=out[3])

    def probe_philox_rate(self):
        """-> measured Philox4x32-10 + 4-plane compare calls per second on this device (instruction-issue ceiling)."""
        v = C.c_double(0)
        _check(lib().mcrg_probe_philox_rate(self._h, C.byref(v)))
        return v.value

    # ---- RGNN
    def rgnn_set_weights(self, W):
        """W: 2x2, W[r, k] (set_weights of rgnn.cpp:38-42); sent column-major."""
        W = np.ascontiguousarray(np.asarray(W, np.float64).reshape(2, 2).T)  # column-major bytes
        _check(lib().mcrg_rgnn_set_weights(self._h, W.ctypes.data))

    def rgnn_eval(self, h=1e-4):
        """-> (u[replica], grad[replica, r, k]) of the current configurations."""
        u = np.zeros(self.n_replicas, np.float64)
        g = np.zeros((self.n_replicas, 4), np.float64)
        _check(lib().mcrg_rgnn_eval(self._h, h, u.ctypes.data, g.ctypes.data))
        return u, g.reshape(-1, 2, 2).transpose(0, 2, 1).copy()  # column-major -> [r, k]

    def rgnn_run(self, n_samples, sweeps_per_sample=1, h=1e-4):
        _check(lib().mcrg_rgnn_run(self._h, n_samples, sweeps_per_sample, h))

    def rgnn_reset(self):
        _check(lib().mcrg_rgnn_accumulators_reset(self._h))

    def rgnn_sums(self):
        """-> [replica, 6] = sum u, sum u^2, sum grad (column-major 2x2)."""
        out = np.zeros((self.n_replicas, 6), np.float64)
        _check(lib().mcrg_rgnn_accumulators_get(self._h, out.ctypes.data))
        return out

    # ---- accumulators
    def reset_accumulators(self):
        _check(lib().mcrg_accumulators_reset(self._h))

    def accumulators(self):
        """-> (acc, accd): acc is an object array [replica, bin, slot] of exact Python ints; accd[replica, bin, 0] is the sum of
        M^4 as float64, derived here from its three exact slots (AccLayout.m4) for callers that want a plain number."""
        lay = acc_layout()
        shape = (self.n_replicas, self.n_bins, lay.n_slots)
        hi = np.zeros(shape, np.int64)
        lo = np.zeros(shape, np.uint64)
        _check(lib().mcrg_accumulators_get(self._h, hi.ctypes.data, lo.ctypes.data))
        acc = hi.astype(object) * (1 << 64) + lo.astype(object)
        d = np.array([[[float(lay.m4(acc[r, b]))] for b in range(self.n_bins)] for r in range(self.n_replicas)], np.float64)
        return acc, d

    def total_limbs_to_device(self, dev_ptr):
        _check(lib().mcrg_accumulators_total_limbs_device(self._h, C.c_void_p(dev_ptr)))
