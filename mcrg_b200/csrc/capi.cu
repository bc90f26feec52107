// extern "C" layer of libmcrg_b200.so (include/mcrg_b200.h): context management and launch sequencing.
// No compute happens on the host; every function either enqueues kernels from kernels.cu or moves bytes.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "capi_internal.cuh"

using namespace mcrg;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(MCRG_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2h(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}

struct GraphKey {
    int m, n_lv, bin, chunk, cur, R, fuse, mode;
    bool operator<(const GraphKey &o) const {
        return std::tie(m, n_lv, bin, chunk, cur, R, fuse, mode) < std::tie(o.m, o.n_lv, o.bin, o.chunk, o.cur, o.R, o.fuse, o.mode);
    }
};

}  // namespace

// Samples in flight: sample s measures into slot s % n_slots (its own set of blocked lattices and popcount cells) and its
// pyramid runs on that slot's side stream, so up to n_slots pyramids overlap each other and the sweeps that follow.  The slot
// count is chosen per context (mcrg_ctx_create): as many as the largest graph holds samples when the blocked lattices are small
// enough, so that inside a graph no measuring sweep ever waits for a pyramid to free its slot.
#ifndef MCRG_PYR_SLOTS
#define MCRG_PYR_SLOTS 64
#endif
constexpr int PYR_SLOTS = MCRG_PYR_SLOTS;  // upper bound (array sizes)
// Samples per CUDA graph of mcrg_run: the large one while that many samples remain, then the small one, then plain launches.
// Every graph ends with a join (the side streams come back, the sweep counter advances): fewer, longer graphs lose less there.
constexpr int GRAPH_CHUNK_LARGE = 64, GRAPH_CHUNK_SMALL = 16;

struct mcrg_ctx {
    int device = 0, L = 0, W = 0, bits = 0, n_replicas = 0, n_bins = 1, full_levels = 0;
    uint64_t seed = 0;
    uint32_t replica_base = 0;
    cudaStream_t stream = nullptr;   // sweeps (and everything else)
    cudaStream_t stream2[PYR_SLOTS] = {};  // blocked-level pyramid of sample s (stream2[s % c->n_slots]), overlapped with the sweeps of
                                           // the next samples and with the pyramids of its neighbours (their accumulation excepted)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_meas[PYR_SLOTS] = {}, ev_pyr[PYR_SLOTS] = {};
    bool pyr_pending[PYR_SLOTS] = {};
    int n_slots = 4;      // slots in use (<= PYR_SLOTS)
    int overlap = 1;      // run the pyramids on the side streams
    int pdl = 3;          // programmatic dependent launch (MCRG_PDL), bits: 1 k_sweep0<false>, 2 k_sweep0<MEASURE>, 4 k_level / k_tail
                          // (4 measured harmful at m = 1: their early CTAs take slots from the sweeps, profiles/r2/pdl_ab.txt)
    int resident_cap = 1 << 22;  // samples per resident launch (64-bit in-launch sums stay exact); MCRG_RESIDENT_MAX_SAMPLES lowers it
    int resident = 1;     // lattices up to RESIDENT_MAX_L: whole replica in one CTA's shared memory, one launch per call
    int resident_threads = 0;    // 0: one thread per column walker (kernels.cu: resident_threads); MCRG_RESIDENT_THREADS forces a block size
    int last_parity = 0;  // which slot (blocked lattices / popcount cells) the last measurement used
    size_t levels_words = 0, cnt_cells = 0;  // words of one set of blocked lattices (n_slots sets)
    uint32_t *planes[2] = {nullptr, nullptr};
    int cur = 0;
    uint32_t *levels = nullptr;
    size_t level_off[MAX_LEVELS + 1] = {0};
    size_t *d_level_off = nullptr;
    unsigned long long *cnt = nullptr;
    long long *S_out = nullptr;
    unsigned long long *acc_lo = nullptr;
    long long *acc_hi = nullptr;
    uint32_t *T4 = nullptr, *T8 = nullptr, *anti = nullptr, *TP = nullptr;
    void *comm = nullptr;        // ncclComm_t of mcrg_comm_init_all (comm.cu), with its device limb buffer
    void *comm_limbs = nullptr;
    int update_mode = MCRG_UPDATE_METROPOLIS;
    int *sw_parent = nullptr;  // union-find forest of the cluster update, allocated by mcrg_set_update
    uint32_t *sw_coins = nullptr;  // cluster coin bitmap, 1 bit per site (rounded up to 128 per replica)
    unsigned long long *d_t = nullptr;
    unsigned long long t_host = 0;
    double *rgnn_W = nullptr, *rgnn_acc = nullptr, *rgnn_u = nullptr, *rgnn_grad = nullptr;
    int32_t *stage = nullptr;
    size_t stage_ints = 0;
    // pipelined uploads (mcrg_set_spins_i32_colmajor_begin / mcrg_set_spins_packed_begin / _commit): two lanes — int32
    // configurations and host-packed ones — each with its own copy stream and device buffer, so that one upload of each
    // kind can be in flight at once (a driver whose host cores cannot pack fast enough sends part of a batch as int32
    // through the copy engine while its threads pack the rest)
    struct UploadLane {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev_copy = nullptr, ev_consumed = nullptr;
        int32_t *buf = nullptr;
        size_t ints = 0;
        int first = 0, count = 0;
        bool pending = false, consumed_recorded = false;
    } up[2];  // [0]: int32 imat, [1]: packed bits
    int last_levels = 0;
    bool measured = false;
    int strip_rows = 0, fuse_sweeps = 1, use_graphs = 1;
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<int, int> auto_R;  // halo depth -> strip height chosen by choose_R
    const uint32_t *ties = nullptr;  // caller-supplied tie coins of the measurement being enqueued (mcrg_measure_supplied)
    size_t tie_stride = 0;
    int n_sm = 148;
};

int mcrg_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
int mcrg_ctx_device(const mcrg_ctx *c) { return c->device; }
cudaStream_t mcrg_ctx_stream(const mcrg_ctx *c) { return c->stream; }
void *mcrg_ctx_comm(const mcrg_ctx *c) { return c->comm; }
void *mcrg_ctx_limbs(const mcrg_ctx *c) { return c->comm_limbs; }
void mcrg_ctx_set_comm(mcrg_ctx *c, void *comm, void *limbs) {
    c->comm = comm;
    c->comm_limbs = limbs;
}

namespace {

// Largest blocked lattice that k_tail (one CTA per replica) takes over from k_level (strips): 512^2 when the batch has a CTA for
// every few SMs — one launch less per sample —, 256^2 for a handful of replicas, whose pyramid is a latency-bound chain already.
int tail_start_size(const mcrg_ctx *c) {
    if (const char *e = getenv("MCRG_TAIL_START")) {
        const int v = atoi(e);
        if (v == 256 || v == 512) return v;
    }
    return c->n_replicas >= 16 ? TAIL_MAX_L : 256;
}

// Row steps a half-sweep over `nrows` rows costs a CTA of `threads` threads (the row distribution of mc_half_sweep_t): a thread
// owns one column, the threads / W row groups take whole row pairs (W >= 32) or equal even chunks (W < 32).
int half_sweep_steps(int nrows, int W, int threads) {
    const int n_grp = threads / W > 0 ? threads / W : 1;
    if (W >= 32) return 2 * ((nrows / 2 + n_grp - 1) / n_grp);
    const int chunk = (nrows + n_grp - 1) / n_grp;
    return (chunk + 1) & ~1;
}

// Estimated duration of one k_sweep0<MEASURE> launch with strips of R rows, in row steps of one SM.  The kernel is bound by
// instruction issue, so a CTA costs its row steps (plus the measurement at ~0.6 of a row step per pair row of a thread and a
// fixed part: staging, barriers, the queue pass and the set-up of every half-sweep), an SM issues at a rate that grows with the
// resident warps up to the 24 the register budget allows, and the CTAs are handed out in waves of n_sm x occupancy — the last
// wave is as slow as its fullest SM.  What this captures: wave quantisation (C5: 1024 CTAs of 16 rows on 444 slots = 3 waves
// for 2.3 waves of work), idle row groups in short strips, halo rows.  Calibrated on profiles/r2/strip_rows_scan.txt.
double sweep0_cost(const mcrg_ctx *c, int R, int H, int nsw) {
    const int L = c->L, W = c->W;
    const int threads = sweep0_threads(L, R, H), occ = sweep0_occupancy(L, R, H);
    if (occ < 1) return 1e300;
    const int rows = R + 2 * H;
    double work = 3.0;
    for (int h = 0; h < 2 * (nsw > 0 ? nsw : 0); ++h) work += 1.0 + half_sweep_steps(rows - 2 - 2 * h, W, threads);
    const int di = threads / W > 0 ? threads / W : 1;
    work += 0.6 * ((R / 2 + di - 1) / di);
    auto rate = [&](int n_ctas_on_sm) { const double r = 0.1 + 0.0375 * (n_ctas_on_sm * threads / 32.0); return r < 1.0 ? r : 1.0; };
    const long long ctas = (long long)c->n_replicas * ((L + R - 1) / R), slots = (long long)c->n_sm * occ;
    const long long full = ctas / slots, rest = ctas % slots;
    double t = (double)full * occ / rate(occ);
    if (rest > 0) {
        const int per_sm = (int)((rest + c->n_sm - 1) / c->n_sm);
        t += per_sm / rate(per_sm);
    }
    return t * work;
}

// Rows per strip: the explicit setting (mcrg_set_tuning / MCRG_STRIP_ROWS) if it fits, else the even height with the lowest
// estimated launch time (sweep0_cost); cached per halo depth, the choice does not change during the life of a context.
int choose_R(mcrg_ctx *c, int H) {
    const int L = c->L;
    const int max_smem = sweep0_max_smem();
    auto fits = [&](int R) { return (long long)sweep0_smem_bytes(L, R, H) <= (long long)max_smem; };
    if (c->strip_rows >= 2 && c->strip_rows <= L && (c->strip_rows & 1) == 0 && fits(c->strip_rows)) return c->strip_rows;
    auto it = c->auto_R.find(H);
    if (it != c->auto_R.end()) return it->second;
    const int nsw = H / 2;  // callers stage H = 2 * (sweeps fused per launch) halo rows (2 for a measurement without a sweep)
    int best = 2;
    double best_cost = 1e300;
    for (int R = 2; R <= L && R <= 256; R += 2) {
        if (!fits(R)) break;
        if (L % R != 0 && (L / R) < 8) continue;  // a ragged last strip is fine among many strips, wasteful among few
        const double cost = sweep0_cost(c, R, H, nsw);
        if (cost < best_cost * (1.0 - 1e-9)) {
            best_cost = cost;
            best = R;
        }
    }
    c->auto_R[H] = best;
    return best;
}

int choose_Rn(int Ln) {
    const int Wn = nat_words(Ln);
    int R = 4096 / Wn;
    static const int cap = [] { const char *e = getenv("MCRG_LEVEL_ROWS"); const int v = e ? atoi(e) : 0; return (v >= 2 && v <= 256 && (v & (v - 1)) == 0) ? v : 64; }();
    if (R > cap) R = cap;
    if (R < 2) R = 2;
    if (R > Ln) R = Ln;
    return R;
}

int ensure_stage(mcrg_ctx *c, size_t ints) {
    if (c->stage_ints >= ints) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_ints = 0;
    CK(cudaMalloc(&c->stage, ints * sizeof(int32_t)));
    c->stage_ints = ints;
    return 0;
}

// level-1 lattice and popcount cells are double-buffered by sample parity (see enqueue_sample)
uint32_t *level_ptr(const mcrg_ctx *c, int lv, int parity) {
    return c->levels + (size_t)parity * c->levels_words + c->level_off[lv];
}
unsigned long long *cnt_ptr(const mcrg_ctx *c, int parity) { return c->cnt + (size_t)parity * c->cnt_cells; }

SweepArgs sweep_args(const mcrg_ctx *c, int R, int nsw, unsigned long long t_off, int parity = 0) {
    SweepArgs a;
    a.src = c->planes[c->cur];
    a.dst = c->planes[1 - c->cur];
    a.level1 = level_ptr(c, 1, parity);
    a.ties = c->ties;
    a.tie_stride = c->tie_stride;
    a.cnt = cnt_ptr(c, parity);
    a.T4 = c->T4;
    a.T8 = c->T8;
    a.anti = c->anti;
    a.d_t = c->d_t;
    a.t_off = t_off;
    a.seed = c->seed;
    a.replica_base = c->replica_base;
    a.L = c->L;
    a.W = c->W;
    a.bits = c->bits;
    a.R = R;
    a.H = nsw > 0 ? 2 * nsw : 2;
    a.nsw = nsw;
    a.strips = (c->L + R - 1) / R;
    return a;
}

// enqueue nsw sweeps (no measurement), fused `fuse` per launch; flips c->cur per launch
void enqueue_sweeps(mcrg_ctx *c, int n, unsigned long long t_off) {
    int done = 0;
    int fuse = c->fuse_sweeps;  // the halo of k fused sweeps is 2k rows on each side: fuse only as deep as shared memory allows
    while (fuse > 1 && (long long)sweep0_smem_bytes(c->L, 2, 2 * fuse) > (long long)sweep0_max_smem()) --fuse;
    while (done < n) {
        const int k = (n - done) < fuse ? (n - done) : fuse;
        const int R = choose_R(c, 2 * k);
        SweepArgs a = sweep_args(c, R, k, t_off + done);
        launch_sweep0(a, c->n_replicas, false, c->stream, (c->pdl & 1) != 0);
        c->cur ^= 1;
        done += k;
    }
}

// one Swendsen-Wang cluster update per sweep-counter tick, in place on the current buffer
void enqueue_cluster_updates(mcrg_ctx *c, int n, unsigned long long t_off) {
    SwArgs a;
    a.planes = c->planes[c->cur];
    a.parent = c->sw_parent;
    a.coins = c->sw_coins;
    a.TP = c->TP;
    a.anti = c->anti;
    a.d_t = c->d_t;
    a.seed = c->seed;
    a.replica_base = c->replica_base;
    a.L = c->L;
    a.W = c->W;
    a.bits = c->bits;
    for (int k = 0; k < n; ++k) {
        a.t_off = t_off + (unsigned long long)k;
        launch_sw_update(a, c->n_replicas, c->stream);
    }
}

// n updates of the configured kind (mcrg_set_update)
void enqueue_updates(mcrg_ctx *c, int n, unsigned long long t_off) {
    if (c->update_mode == MCRG_UPDATE_CLUSTER) enqueue_cluster_updates(c, n, t_off);
    else enqueue_sweeps(c, n, t_off);
}

// make the main stream wait for every pyramid still in flight on the side streams (no-op when nothing is pending)
void join_pyramids(mcrg_ctx *c) {
    for (int p = 0; p < PYR_SLOTS; ++p)
        if (c->pyr_pending[p]) {
            cudaStreamWaitEvent(c->stream, c->ev_pyr[p], 0);
            c->pyr_pending[p] = false;
        }
}

// enqueue: measure the current configuration at levels 0..n_lv (+ accumulate), fused with the first of
// `m` sweeps; then the remaining m-1 sweeps.  The blocked-level kernels of sample s (k_level, k_tail) only read
// the level-1 lattice and the popcount cells written by k_sweep0<MEASURE>; both are double-buffered by `parity`,
// so they run on a side stream while the main stream already sweeps towards sample s+1.  ALL blocked lattices are
// double-buffered by parity and the two parities have their own side stream, so the pyramid of sample s+1 does not queue
// behind that of sample s either (a single replica's pyramid is a chain of small, latency-bound launches longer than its
// sweep); the accumulation needs no order: k_tail adds into the 128-bit sums with atomics (exact in any interleaving).
void enqueue_sample(mcrg_ctx *c, int n_lv, int m, int accumulate, int bin, unsigned long long t_off, int parity,
                    cudaEvent_t *probe = nullptr) {
    const bool overlap = c->overlap && probe == nullptr;
    cudaStream_t s_pyr = overlap ? c->stream2[parity] : c->stream;
    if (c->pyr_pending[parity]) {  // the pyramid of sample s-2 used this buffer pair
        cudaStreamWaitEvent(c->stream, c->ev_pyr[parity], 0);
        c->pyr_pending[parity] = false;
    }
    const bool cluster = c->update_mode == MCRG_UPDATE_CLUSTER;
    const int first = (m > 0 && !cluster) ? 1 : 0;  // a Metropolis sweep is fused into the measuring kernel
    const int R = choose_R(c, 2);
    SweepArgs a = sweep_args(c, R, first, t_off, parity);
    if (probe) cudaEventRecord(probe[0], c->stream);
    launch_sweep0(a, c->n_replicas, true, c->stream, (c->pdl & 2) != 0);
    if (probe) cudaEventRecord(probe[1], c->stream);
    if (first) c->cur ^= 1;
    if (overlap) {
        cudaEventRecord(c->ev_meas[parity], c->stream);
        cudaStreamWaitEvent(s_pyr, c->ev_meas[parity], 0);
    }
    int lv = 1;
    while (lv <= n_lv && (c->L >> lv) > tail_start_size(c)) {
        LevelArgs la;
        la.in = level_ptr(c, lv, parity);
        la.out = lv < n_lv ? level_ptr(c, lv + 1, parity) : nullptr;
        la.ties = c->ties;
        la.tie_stride = c->tie_stride;
        la.L = c->L;
        la.cnt = cnt_ptr(c, parity);
        la.d_t = c->d_t;
        la.t_off = t_off;
        la.seed = c->seed;
        la.replica_base = c->replica_base;
        la.Ln = c->L >> lv;
        la.level = lv;
        la.R = choose_Rn(la.Ln);
        la.strips = la.Ln / la.R;
        launch_level(la, c->n_replicas, s_pyr, (c->pdl & 4) != 0);
        ++lv;
    }
    TailArgs ta;
    ta.in = lv <= n_lv ? level_ptr(c, lv, parity) : nullptr;
    ta.levels_out = c->levels + (size_t)parity * c->levels_words;
    ta.ties = c->ties;
    ta.tie_stride = c->tie_stride;
    ta.level_off = c->d_level_off;
    ta.cnt = cnt_ptr(c, parity);
    ta.S_out = c->S_out;
    ta.acc_lo = c->acc_lo;
    ta.acc_hi = c->acc_hi;
    ta.d_t = c->d_t;
    ta.t_off = t_off;
    ta.seed = c->seed;
    ta.replica_base = c->replica_base;
    ta.L = c->L;
    ta.start = lv;
    ta.n_levels = n_lv;
    ta.n_bins = c->n_bins;
    ta.bin = bin;
    ta.accumulate = accumulate;
    if (probe) cudaEventRecord(probe[2], c->stream);
    launch_tail(ta, c->n_replicas, s_pyr, (c->pdl & 4) != 0);
    if (probe) cudaEventRecord(probe[3], c->stream);
    if (overlap) {
        cudaEventRecord(c->ev_pyr[parity], s_pyr);
        c->pyr_pending[parity] = true;
    }
    c->last_levels = n_lv;
    c->last_parity = parity;
    c->measured = true;
    if (cluster) enqueue_cluster_updates(c, m, t_off);
    else if (m > 1) enqueue_sweeps(c, m - 1, t_off + 1);
    if (probe) cudaEventRecord(probe[4], c->stream);
}

// an explicit strip height (mcrg_set_tuning / MCRG_STRIP_ROWS) asks for the strip kernel
bool use_resident(const mcrg_ctx *c) {
    return c->resident && c->strip_rows == 0 && c->L <= RESIDENT_MAX_L && c->update_mode == MCRG_UPDATE_METROPOLIS;
}

void enqueue_resident_launch(mcrg_ctx *c, bool measure, int n_samples, int m, int n_lv, int accumulate, int bin, unsigned long long t_off);

// a resident run, split into launches of at most 2^22 samples (the kernel keeps 64-bit sums per launch, see k_resident)
void enqueue_resident(mcrg_ctx *c, bool measure, int n_samples, int m, int n_lv, int accumulate, int bin) {
    const int cap = measure ? c->resident_cap : n_samples;
    unsigned long long t_off = 0;
    for (int done = 0; done < n_samples;) {
        const int n = n_samples - done < cap ? n_samples - done : cap;
        enqueue_resident_launch(c, measure, n, m, n_lv, accumulate, bin, t_off);
        t_off += (unsigned long long)n * m;
        done += n;
    }
}

void enqueue_resident_launch(mcrg_ctx *c, bool measure, int n_samples, int m, int n_lv, int accumulate, int bin, unsigned long long t_off) {
    ResidentArgs a;
    a.planes = c->planes[c->cur];
    a.T4 = c->T4;
    a.T8 = c->T8;
    a.anti = c->anti;
    a.d_t = c->d_t;
    a.t_off = t_off;
    a.seed = c->seed;
    a.replica_base = c->replica_base;
    a.L = c->L;
    a.W = c->W;
    a.bits = c->bits;
    a.n_samples = n_samples;
    a.n_replicas = c->n_replicas;
    a.m = m;
    a.n_levels = n_lv;
    a.accumulate = accumulate;
    a.n_bins = c->n_bins;
    a.bin = bin;
    a.acc_lo = c->acc_lo;
    a.acc_hi = c->acc_hi;
    a.S_out = c->S_out;
    launch_resident(a, c->n_replicas, measure, c->resident_threads, c->stream);
}

int clamp_levels(const mcrg_ctx *c, int max_levels) {
    int n_lv = c->full_levels;
    if (max_levels >= 0 && max_levels < n_lv) n_lv = max_levels;
    return n_lv;
}

void destroy_graphs(mcrg_ctx *c) {
    for (auto &kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
}

}  // namespace

extern "C" {

const char *mcrg_last_error(void) { return g_err; }

int mcrg_version(void) { return 100; }

int mcrg_device_count(int *n) {
    if (!n) return fail(MCRG_ERR_ARG, "null pointer");
    *n = 0;
    CK(cudaGetDeviceCount(n));
    return 0;
}

int mcrg_levels_full(int L) {
    if (L < 2) return 0;
    return (int)std::floor(std::log((double)L) / std::log(2.0) + 1e-9) - 1;  // mcrg.cpp:43
}

int mcrg_ctx_create(int device, int L, int n_replicas, uint64_t seed, uint32_t replica_base, int n_bins,
                    mcrg_ctx **out) {
    if (!out) return fail(MCRG_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (!is_pow2(L) || L < 2 || L > 16384) return fail(MCRG_ERR_ARG, "L=%d must be a power of two in [2, 16384]", L);
    if (n_replicas < 1 || n_replicas > 65535) return fail(MCRG_ERR_ARG, "n_replicas=%d out of range [1, 65535]", n_replicas);
    if (n_bins < 1) return fail(MCRG_ERR_ARG, "n_bins=%d must be >= 1", n_bins);
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(MCRG_ERR_CUDA, "device %d not present (%d CUDA devices)", device, n_dev);
    CK(cudaSetDevice(device));
    mcrg_ctx *c = new mcrg_ctx();
    c->device = device;
    c->L = L;
    c->W = l0_words(L);
    c->bits = l0_bits(L);
    c->n_replicas = n_replicas;
    c->n_bins = n_bins;
    c->seed = seed;
    c->replica_base = replica_base;
    c->full_levels = ilog2h(L) - 1;  // mcrg.cpp:43; 0 for the 2x2 lattice
    if (const char *e = getenv("MCRG_STRIP_ROWS")) {  // same rules as mcrg_set_tuning; anything else is ignored
        const int v = atoi(e);
        if (v >= 2 && v <= L && (v & 1) == 0) c->strip_rows = v;
    }
    if (const char *e = getenv("MCRG_FUSE_SWEEPS")) {
        const int v = atoi(e);
        c->fuse_sweeps = v < 1 ? 1 : (v > 8 ? 8 : v);
    }
    if (const char *e = getenv("MCRG_USE_GRAPHS")) c->use_graphs = atoi(e);
    *out = c;  // so that a failure below can still be cleaned up by mcrg_ctx_destroy
    sweep0_max_smem();  // opt in to large dynamic shared memory once, outside any stream capture
    CK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
    size_t off = 0;
    for (int lv = 1; lv <= (c->full_levels > 1 ? c->full_levels : 1); ++lv) {  // level 1 always exists: k_sweep0 writes it
        const int Ln = L >> lv;
        c->level_off[lv] = off;
        size_t words = ((size_t)n_replicas * Ln * nat_words(Ln) + 3) & ~(size_t)3;  // 16-byte aligned for 128-bit loads
        off += words;
    }
    c->levels_words = off + 4;  // one set of blocked lattices; n_slots sets are allocated (see enqueue_sample)
    const size_t n_cnt = (size_t)n_replicas * (MAX_LEVELS + 1) * 4;
    c->cnt_cells = n_cnt;
    {
        // Slots (samples in flight): one per sample of the largest graph when that costs at most 4 GiB (C4: 64 x 28 MB), halved
        // until it does; lattices that run on the resident kernel never use more than one (their pyramids stay in shared memory),
        // so they get the minimum.  MCRG_SLOTS overrides (experiments, tests).  Results never depend on the count.
        const size_t per_slot = c->levels_words * 4 + n_cnt * 8;
        int n = L <= RESIDENT_MAX_L ? 4 : (PYR_SLOTS < GRAPH_CHUNK_LARGE ? PYR_SLOTS : GRAPH_CHUNK_LARGE);
        while (n > 4 && (size_t)n * per_slot > ((size_t)4 << 30)) n >>= 1;
        if (const char *e = getenv("MCRG_SLOTS")) {
            const int v = atoi(e);
            if (v >= 1 && v <= PYR_SLOTS) n = v;
        }
        c->n_slots = n;
    }
    {
        // The sweeps are the critical path of a run (sample s+1 cannot start before the sweep of sample s is done, a pyramid only
        // has to finish before its slot comes round again): the sweep stream gets the high priority, the pyramid streams the low
        // one, and the graphs are instantiated with per-node priorities (mcrg_run) — otherwise every node of a graph runs at the
        // priority of the stream the graph is launched into.  Measured (profiles/r2/priority_slots_ab.txt): pyramids above the
        // sweeps C4 m=1 -4 %, below +0.8 % (C5 +2 %).  MCRG_PYR_PRIO = hi | same | lo (default) for experiments.
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const char *pp = getenv("MCRG_PYR_PRIO");
        const bool pyr_hi = pp && !strcmp(pp, "hi"), pyr_same = pp && !strcmp(pp, "same");
        CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, (pyr_hi || pyr_same) ? prio_lo : prio_hi));
        for (int p = 0; p < c->n_slots; ++p)
            CK(cudaStreamCreateWithPriority(&c->stream2[p], cudaStreamNonBlocking, pyr_hi ? prio_hi : prio_lo));
    }
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    for (int p = 0; p < c->n_slots; ++p) {
        CK(cudaEventCreateWithFlags(&c->ev_meas[p], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_pyr[p], cudaEventDisableTiming));
    }
    if (const char *e = getenv("MCRG_OVERLAP")) c->overlap = atoi(e);
    if (const char *e = getenv("MCRG_PDL")) c->pdl = atoi(e);
    if (const char *e = getenv("MCRG_RESIDENT")) c->resident = atoi(e);
    if (const char *e = getenv("MCRG_RESIDENT_THREADS")) c->resident_threads = atoi(e);
    if (const char *e = getenv("MCRG_RESIDENT_MAX_SAMPLES")) {
        const int v = atoi(e);
        if (v >= 1 && v <= (1 << 22)) c->resident_cap = v;
    }
    const size_t plane_words = (size_t)n_replicas * 2 * L * c->W;
    CK(cudaMalloc(&c->planes[0], plane_words * 4));
    CK(cudaMalloc(&c->planes[1], plane_words * 4));
    CK(cudaMalloc(&c->levels, (size_t)c->n_slots * c->levels_words * 4));
    CK(cudaMemsetAsync(c->levels, 0, (size_t)c->n_slots * c->levels_words * 4, c->stream));
    CK(cudaMalloc(&c->d_level_off, sizeof(c->level_off)));
    CK(cudaMemcpyAsync(c->d_level_off, c->level_off, sizeof(c->level_off), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMalloc(&c->cnt, (size_t)c->n_slots * n_cnt * 8));
    CK(cudaMemsetAsync(c->cnt, 0, (size_t)c->n_slots * n_cnt * 8, c->stream));
    CK(cudaMalloc(&c->S_out, n_cnt * 8));
    CK(cudaMemsetAsync(c->S_out, 0, n_cnt * 8, c->stream));
    const size_t n_acc = (size_t)n_replicas * n_bins * N_SLOTS;
    CK(cudaMalloc(&c->acc_lo, n_acc * 8));
    CK(cudaMalloc(&c->acc_hi, n_acc * 8));
    CK(cudaMalloc(&c->T4, n_replicas * 4));
    CK(cudaMalloc(&c->T8, n_replicas * 4));
    CK(cudaMalloc(&c->anti, n_replicas * 4));
    CK(cudaMalloc(&c->TP, n_replicas * 4));
    CK(cudaMalloc(&c->rgnn_W, 4 * sizeof(double)));
    CK(cudaMalloc(&c->rgnn_acc, (size_t)n_replicas * 6 * sizeof(double)));
    CK(cudaMalloc(&c->rgnn_u, (size_t)n_replicas * sizeof(double)));
    CK(cudaMalloc(&c->rgnn_grad, (size_t)n_replicas * 4 * sizeof(double)));
    CK(cudaMemsetAsync(c->rgnn_W, 0, 4 * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->rgnn_acc, 0, (size_t)n_replicas * 6 * sizeof(double), c->stream));
    CK(cudaMalloc(&c->d_t, 8));
    CK(cudaMemsetAsync(c->d_t, 0, 8, c->stream));
    launch_init_cold(c->planes[0], L, n_replicas, c->stream);
    CK(cudaGetLastError());
    int rc = mcrg_accumulators_reset(c);
    if (rc) return rc;
    const double Kc = -0.5 * std::log(1.0 + std::sqrt(2.0));  // main.cpp:13
    rc = mcrg_set_couplings(c, &Kc, 1);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mcrg_ctx_destroy(mcrg_ctx *c) {
    if (!c) return 0;
    if (c->comm) {  // a context leaving its group: the group's other members keep their (now unusable) communicators
        mcrg_ctx *self = c;
        mcrg_comm_destroy_all(1, &self);
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int p = 0; p < PYR_SLOTS; ++p)
        if (c->stream2[p]) cudaStreamSynchronize(c->stream2[p]);
    destroy_graphs(c);
    cudaFree(c->planes[0]);
    cudaFree(c->planes[1]);
    cudaFree(c->levels);
    cudaFree(c->d_level_off);
    cudaFree(c->cnt);
    cudaFree(c->S_out);
    cudaFree(c->acc_lo);
    cudaFree(c->acc_hi);
    cudaFree(c->T4);
    cudaFree(c->T8);
    cudaFree(c->anti);
    cudaFree(c->TP);
    cudaFree(c->sw_parent);
    cudaFree(c->sw_coins);
    cudaFree(c->d_t);
    cudaFree(c->rgnn_W);
    cudaFree(c->rgnn_acc);
    cudaFree(c->rgnn_u);
    cudaFree(c->rgnn_grad);
    cudaFree(c->stage);
    for (auto &u : c->up) {
        if (u.stream) cudaStreamSynchronize(u.stream);
        cudaFree(u.buf);
        if (u.ev_copy) cudaEventDestroy(u.ev_copy);
        if (u.ev_consumed) cudaEventDestroy(u.ev_consumed);
        if (u.stream) cudaStreamDestroy(u.stream);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int p = 0; p < PYR_SLOTS; ++p) {
        if (c->ev_meas[p]) cudaEventDestroy(c->ev_meas[p]);
        if (c->ev_pyr[p]) cudaEventDestroy(c->ev_pyr[p]);
    }
    for (int p = 0; p < PYR_SLOTS; ++p)
        if (c->stream2[p]) cudaStreamDestroy(c->stream2[p]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int mcrg_sync(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

uint64_t mcrg_stream_handle(mcrg_ctx *c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }

int mcrg_timer_start(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return 0;
}

int mcrg_timer_stop(mcrg_ctx *c, float *ms) {
    if (!c || !ms) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

int mcrg_set_tuning(mcrg_ctx *c, int strip_rows, int fuse_sweeps, int use_graphs) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (strip_rows != 0 && ((strip_rows & 1) || strip_rows < 2 || strip_rows > c->L))
        return fail(MCRG_ERR_ARG, "strip_rows=%d must be 0 or an even number in [2, L]", strip_rows);
    if (fuse_sweeps < 1 || fuse_sweeps > 8) return fail(MCRG_ERR_ARG, "fuse_sweeps=%d must be in [1, 8]", fuse_sweeps);
    c->strip_rows = strip_rows;
    c->fuse_sweeps = fuse_sweeps;
    c->use_graphs = use_graphs;
    c->auto_R.clear();
    destroy_graphs(c);
    return 0;
}

int mcrg_strip_plan(mcrg_ctx *c, int fused_sweeps, int strip_rows, int *rows_out, double *cost_out) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (fused_sweeps < 1 || fused_sweeps > 8) return fail(MCRG_ERR_ARG, "fused_sweeps=%d must be in [1, 8]", fused_sweeps);
    CK(cudaSetDevice(c->device));
    const int H = 2 * fused_sweeps;
    int R = strip_rows;
    if (R == 0) {  // the automatic choice, whatever mcrg_set_tuning fixed
        const int keep = c->strip_rows;
        c->strip_rows = 0;
        R = choose_R(c, H);
        c->strip_rows = keep;
    } else if ((R & 1) || R < 2 || R > c->L) {
        return fail(MCRG_ERR_ARG, "strip_rows=%d must be 0 or an even number in [2, L]", R);
    }
    if (rows_out) *rows_out = R;
    if (cost_out) *cost_out = (long long)sweep0_smem_bytes(c->L, R, H) <= (long long)sweep0_max_smem() ? sweep0_cost(c, R, H, fused_sweeps) : -1.0;
    return 0;
}

int mcrg_set_update(mcrg_ctx *c, int mode) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (mode != MCRG_UPDATE_METROPOLIS && mode != MCRG_UPDATE_CLUSTER) return fail(MCRG_ERR_ARG, "unknown update mode %d", mode);
    CK(cudaSetDevice(c->device));
    if (mode == MCRG_UPDATE_CLUSTER && !c->sw_parent) {
        CK(cudaMalloc(&c->sw_parent, (size_t)c->n_replicas * c->L * c->L * sizeof(int)));
        CK(cudaMalloc(&c->sw_coins, (size_t)c->n_replicas * (((size_t)c->L * c->L + 127) / 128) * 16));
    }
    c->update_mode = mode;
    return 0;
}

int mcrg_set_couplings(mcrg_ctx *c, const double *K, int n) {
    if (!c || !K) return fail(MCRG_ERR_ARG, "null pointer");
    if (n != 1 && n != c->n_replicas) return fail(MCRG_ERR_ARG, "n=%d must be 1 or n_replicas=%d", n, c->n_replicas);
    CK(cudaSetDevice(c->device));
    std::vector<uint32_t> t4(c->n_replicas), t8(c->n_replicas), an(c->n_replicas), tp(c->n_replicas);
    for (int r = 0; r < c->n_replicas; ++r) {
        const double k = K[n == 1 ? 0 : r];
        if (!std::isfinite(k)) return fail(MCRG_ERR_ARG, "coupling %d is not finite", r);
        // min(1, exp(2 K s h)) with s h in {2, 4}: exp(-4|K|), exp(-8|K|); 32 fractional bits, floor, clamped
        double p4 = std::floor(std::exp(-4.0 * std::fabs(k)) * 4294967296.0);
        double p8 = std::floor(std::exp(-8.0 * std::fabs(k)) * 4294967296.0);
        if (p4 > 4294967295.0) p4 = 4294967295.0;
        if (p8 > 4294967295.0) p8 = 4294967295.0;
        t4[r] = (uint32_t)p4;
        t8[r] = (uint32_t)p8;
        an[r] = k > 0.0 ? 0xFFFFFFFFu : 0u;
        // cluster update: bond probability 1 - exp(-2|K|) (ising.cpp:9), same fixed-point convention
        double pb = std::floor((1.0 - std::exp(-2.0 * std::fabs(k))) * 4294967296.0);
        if (pb > 4294967295.0) pb = 4294967295.0;
        tp[r] = (uint32_t)pb;
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(c->T4, t4.data(), c->n_replicas * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->T8, t8.data(), c->n_replicas * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->anti, an.data(), c->n_replicas * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->TP, tp.data(), c->n_replicas * 4, cudaMemcpyHostToDevice));
    return 0;
}

int mcrg_init_hot(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    launch_init_hot(c->planes[c->cur], c->L, c->n_replicas, c->seed, c->replica_base, c->stream);
    CK(cudaGetLastError());
    return 0;
}

int mcrg_init_cold(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    launch_init_cold(c->planes[c->cur], c->L, c->n_replicas, c->stream);
    CK(cudaGetLastError());
    return 0;
}

int mcrg_set_spins_i32_colmajor(mcrg_ctx *c, int first, int count, const int32_t *host) {
    if (!c || !host) return fail(MCRG_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->n_replicas) return fail(MCRG_ERR_ARG, "replica range [%d, %d) out of [0, %d)", first, first + count, c->n_replicas);
    CK(cudaSetDevice(c->device));
    const size_t per = (size_t)c->L * c->L;
    size_t chunk = ((size_t)256 << 20) / (per * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)count) chunk = count;
    if (count == 0) return 0;
    int rc = ensure_stage(c, chunk * per);
    if (rc) return rc;
    for (size_t done = 0; done < (size_t)count; done += chunk) {
        const size_t k = ((size_t)count - done) < chunk ? ((size_t)count - done) : chunk;
        CK(cudaMemcpyAsync(c->stage, host + done * per, k * per * 4, cudaMemcpyHostToDevice, c->stream));
        launch_pack0(c->stage, c->planes[c->cur] + (size_t)(first + done) * 2 * c->L * c->W, c->L, (int)k, c->stream);
        CK(cudaGetLastError());
        if (done + k < (size_t)count) CK(cudaStreamSynchronize(c->stream));  // staging buffer is reused
    }
    return 0;
}

int mcrg_set_spins_packed(mcrg_ctx *c, int first, int count, const uint32_t *packed) {
    if (!c || !packed) return fail(MCRG_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->n_replicas) return fail(MCRG_ERR_ARG, "replica range [%d, %d) out of [0, %d)", first, first + count, c->n_replicas);
    if (count == 0) return 0;
    CK(cudaSetDevice(c->device));
    const size_t words = mcrg_packed_words(c->L, count);
    int rc = ensure_stage(c, words);  // 1 bit per spin: the staging buffer is small
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->stage, packed, words * 4, cudaMemcpyHostToDevice, c->stream));
    launch_pack_nat(reinterpret_cast<const uint32_t *>(c->stage), c->planes[c->cur] + (size_t)first * 2 * c->L * c->W, c->L, count, c->stream);
    CK(cudaGetLastError());
    return 0;
}

// start copying `ints` 32-bit words from pinned host memory into lane `k`'s device buffer on that lane's copy stream
static int upload_begin(mcrg_ctx *c, int k, int first, int count, const void *pinned_host, size_t ints) {
    mcrg_ctx::UploadLane &u = c->up[k];
    if (u.pending) return fail(MCRG_ERR_STATE, "an upload of this kind is already in flight: call mcrg_set_spins_commit first");
    CK(cudaSetDevice(c->device));
    if (!u.stream) {
        CK(cudaStreamCreateWithFlags(&u.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&u.ev_copy, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&u.ev_consumed, cudaEventDisableTiming));
    }
    if (u.ints < ints) {
        CK(cudaStreamSynchronize(u.stream));
        CK(cudaStreamSynchronize(c->stream));
        if (u.buf) cudaFree(u.buf);
        u.buf = nullptr;
        u.ints = 0;
        CK(cudaMalloc(&u.buf, ints * sizeof(int32_t)));
        u.ints = ints;
        u.consumed_recorded = false;
    }
    // the previous commit's unpack kernel must have consumed the buffer before it is overwritten
    if (u.consumed_recorded) CK(cudaStreamWaitEvent(u.stream, u.ev_consumed, 0));
    CK(cudaMemcpyAsync(u.buf, pinned_host, ints * sizeof(int32_t), cudaMemcpyHostToDevice, u.stream));
    CK(cudaEventRecord(u.ev_copy, u.stream));
    u.first = first;
    u.count = count;
    u.pending = true;
    return 0;
}

int mcrg_set_spins_i32_colmajor_begin(mcrg_ctx *c, int first, int count, const int32_t *pinned_host) {
    if (!c || !pinned_host) return fail(MCRG_ERR_ARG, "null pointer");
    if (first < 0 || count < 1 || first + count > c->n_replicas) return fail(MCRG_ERR_ARG, "replica range [%d, %d) out of [0, %d)", first, first + count, c->n_replicas);
    return upload_begin(c, 0, first, count, pinned_host, (size_t)count * c->L * c->L);
}

// The same for configurations already packed on the host (mcrg_host_pack_i32_colmajor): 1 bit per spin over PCIe.
int mcrg_set_spins_packed_begin(mcrg_ctx *c, int first, int count, const uint32_t *pinned_host_packed) {
    if (!c || !pinned_host_packed) return fail(MCRG_ERR_ARG, "null pointer");
    if (first < 0 || count < 1 || first + count > c->n_replicas) return fail(MCRG_ERR_ARG, "replica range [%d, %d) out of [0, %d)", first, first + count, c->n_replicas);
    return upload_begin(c, 1, first, count, pinned_host_packed, mcrg_packed_words(c->L, count));
}

int mcrg_set_spins_commit(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (!c->up[0].pending && !c->up[1].pending) return fail(MCRG_ERR_STATE, "no upload in flight");
    CK(cudaSetDevice(c->device));
    for (int k = 0; k < 2; ++k) {
        mcrg_ctx::UploadLane &u = c->up[k];
        if (!u.pending) continue;
        CK(cudaStreamWaitEvent(c->stream, u.ev_copy, 0));
        uint32_t *dst = c->planes[c->cur] + (size_t)u.first * 2 * c->L * c->W;
        if (k == 1) launch_pack_nat(reinterpret_cast<const uint32_t *>(u.buf), dst, c->L, u.count, c->stream);
        else launch_pack0(u.buf, dst, c->L, u.count, c->stream);
        CK(cudaGetLastError());
        CK(cudaEventRecord(u.ev_consumed, c->stream));
        u.consumed_recorded = true;
        u.pending = false;
    }
    return 0;
}

int mcrg_get_spins_i32_colmajor(mcrg_ctx *c, int first, int count, int32_t *host) {
    if (!c || !host) return fail(MCRG_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->n_replicas) return fail(MCRG_ERR_ARG, "replica range [%d, %d) out of [0, %d)", first, first + count, c->n_replicas);
    CK(cudaSetDevice(c->device));
    if (count == 0) return 0;
    const size_t per = (size_t)c->L * c->L;
    size_t chunk = ((size_t)256 << 20) / (per * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)count) chunk = count;
    int rc = ensure_stage(c, chunk * per);
    if (rc) return rc;
    for (size_t done = 0; done < (size_t)count; done += chunk) {
        const size_t k = ((size_t)count - done) < chunk ? ((size_t)count - done) : chunk;
        launch_unpack0(c->planes[c->cur] + (size_t)(first + done) * 2 * c->L * c->W, c->stage, c->L, (int)k, c->stream);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host + done * per, c->stage, k * per * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int mcrg_get_level_spins_i32_colmajor(mcrg_ctx *c, int replica, int level, int32_t *host) {
    if (!c || !host) return fail(MCRG_ERR_ARG, "null pointer");
    if (replica < 0 || replica >= c->n_replicas) return fail(MCRG_ERR_ARG, "replica %d out of range", replica);
    // level 1 is written by every measurement (k_sweep0<MEASURE>), deeper levels only up to the requested depth
    const int have = c->measured ? (c->last_levels > 1 ? c->last_levels : 1) : 0;
    if (level < 1 || level > have || (c->L >> level) < 1) return fail(MCRG_ERR_STATE, "level %d not produced by the last measurement (levels 1..%d)", level, have);
    CK(cudaSetDevice(c->device));
    const int Ln = c->L >> level;
    const size_t per = (size_t)Ln * Ln;
    int rc = ensure_stage(c, per);
    if (rc) return rc;
    launch_unpackN(level_ptr(c, level, c->last_parity) + (size_t)replica * Ln * nat_words(Ln), c->stage, Ln, 1, c->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, c->stage, per * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mcrg_get_sweep_counter(mcrg_ctx *c, uint64_t *t) {
    if (!c || !t) return fail(MCRG_ERR_ARG, "null pointer");
    *t = c->t_host;
    return 0;
}

int mcrg_set_sweep_counter(mcrg_ctx *c, uint64_t t) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    c->t_host = t;
    unsigned long long v = t;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(c->d_t, &v, 8, cudaMemcpyHostToDevice));
    return 0;
}

int mcrg_sweep(mcrg_ctx *c, int n_sweeps) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (n_sweeps < 0) return fail(MCRG_ERR_ARG, "n_sweeps=%d < 0", n_sweeps);
    if (n_sweeps == 0) return 0;
    CK(cudaSetDevice(c->device));
    if (use_resident(c)) enqueue_resident(c, false, 1, n_sweeps, 0, 0, 0);
    else enqueue_updates(c, n_sweeps, 0);
    launch_advance_t(c->d_t, (unsigned long long)n_sweeps, c->stream);
    c->t_host += (unsigned long long)n_sweeps;
    CK(cudaGetLastError());
    return 0;
}

int mcrg_measure(mcrg_ctx *c, int max_levels, int64_t *S, int *n_lv_out) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const int n_lv = clamp_levels(c, max_levels);
    enqueue_sample(c, n_lv, 0, 0, 0, 0, 0);
    join_pyramids(c);
    CK(cudaGetLastError());
    if (n_lv_out) *n_lv_out = n_lv;
    if (S) {
        std::vector<long long> tmp((size_t)c->n_replicas * (MAX_LEVELS + 1) * 4);
        CK(cudaMemcpyAsync(tmp.data(), c->S_out, tmp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < c->n_replicas; ++r)
            for (int lv = 0; lv <= n_lv; ++lv)
                for (int k = 0; k < 4; ++k)
                    S[((size_t)r * (n_lv + 1) + lv) * 4 + k] = tmp[((size_t)r * (MAX_LEVELS + 1) + lv) * 4 + k];
    }
    return 0;
}

size_t mcrg_tie_words(int L, int max_levels) {
    if (L < 2 || (L & (L - 1))) return 0;
    int n_lv = ilog2h(L) - 1;
    if (max_levels >= 0 && max_levels < n_lv) n_lv = max_levels;
    return tie_level_off(L, n_lv + 1);
}

int mcrg_measure_supplied(mcrg_ctx *c, int max_levels, const uint32_t *tie_bits, int64_t *S, int *n_lv_out) {
    if (!c || !tie_bits) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    const int n_lv = clamp_levels(c, max_levels);
    const size_t per = mcrg_tie_words(c->L, n_lv), words = per * (size_t)c->n_replicas;
    uint32_t *dev = nullptr;
    if (words > 0) {
        CK(cudaMalloc(&dev, words * 4));
        CK(cudaMemcpyAsync(dev, tie_bits, words * 4, cudaMemcpyHostToDevice, c->stream));
    }
    c->ties = dev;
    c->tie_stride = per;
    const int rc = mcrg_measure(c, max_levels, S, n_lv_out);
    c->ties = nullptr;
    c->tie_stride = 0;
    cudaStreamSynchronize(c->stream);
    cudaFree(dev);
    return rc;
}

int mcrg_observables(mcrg_ctx *c, int64_t *Snn, int64_t *Snnn, int64_t *Splaq, int64_t *M) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    std::vector<int64_t> S((size_t)c->n_replicas * 4);
    int rc = mcrg_measure(c, 0, S.data(), nullptr);
    if (rc) return rc;
    for (int r = 0; r < c->n_replicas; ++r) {
        if (Snn) Snn[r] = S[(size_t)r * 4 + 0];
        if (Snnn) Snnn[r] = S[(size_t)r * 4 + 1];
        if (Splaq) Splaq[r] = S[(size_t)r * 4 + 2];
        if (M) M[r] = S[(size_t)r * 4 + 3];
    }
    return 0;
}

int mcrg_run(mcrg_ctx *c, int n_samples, int sweeps_per_sample, int max_levels, int bin) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (n_samples < 0 || sweeps_per_sample < 0) return fail(MCRG_ERR_ARG, "negative count");
    if (bin < 0 || bin >= c->n_bins) return fail(MCRG_ERR_ARG, "bin %d out of [0, %d)", bin, c->n_bins);
    if (n_samples == 0) return 0;
    CK(cudaSetDevice(c->device));
    const int n_lv = clamp_levels(c, max_levels);
    const int m = sweeps_per_sample;
    if (use_resident(c)) {  // small lattices: the whole block of samples is one launch
        enqueue_resident(c, true, n_samples, m, n_lv, 1, bin);
        if (m > 0) launch_advance_t(c->d_t, (unsigned long long)n_samples * m, c->stream);
        c->t_host += (unsigned long long)n_samples * m;
        c->last_levels = n_lv;
        c->measured = false;  // the blocked lattices of a resident run never leave shared memory
        CK(cudaGetLastError());
        return 0;
    }
    int done = 0;
    choose_R(c, 2);  // strip heights are chosen (occupancy queries) before any stream capture starts
    for (int k = 1; k <= c->fuse_sweeps; ++k) choose_R(c, 2 * k);
    if (c->use_graphs) {
        static const int forced_chunk = [] { const char *e = getenv("MCRG_GRAPH_CHUNK"); const int v = e ? atoi(e) : 0; return (v >= 2 && v <= 256) ? v : 0; }();
        const int tiers[2] = {forced_chunk ? forced_chunk : GRAPH_CHUNK_LARGE, forced_chunk ? forced_chunk : GRAPH_CHUNK_SMALL};
        for (int tier = 0; tier < 2; ++tier) {
            const int chunk = tiers[tier];
            while (n_samples - done >= chunk) {
                GraphKey key{m, n_lv, bin, chunk, c->cur, c->strip_rows, c->fuse_sweeps, c->update_mode};
                auto it = c->graphs.find(key);
                const int cur_before = c->cur;
                if (it == c->graphs.end()) {
                    cudaGraph_t g = nullptr;
                    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                    for (int s = 0; s < chunk; ++s) enqueue_sample(c, n_lv, m, 1, bin, (unsigned long long)s * m, s % c->n_slots);
                    join_pyramids(c);  // every graph is self-contained: the side streams join back before the capture ends
                    launch_advance_t(c->d_t, (unsigned long long)chunk * m, c->stream);
                    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
                    if (e != cudaSuccess) {
                        c->cur = cur_before;
                        return fail(MCRG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
                    }
                    cudaGraphExec_t ge = nullptr;
                    static const bool node_prio = [] { const char *e = getenv("MCRG_GRAPH_PRIO"); return !(e && atoi(e) == 0); }();  // see mcrg_ctx_create
                    e = cudaGraphInstantiate(&ge, g, node_prio ? cudaGraphInstantiateFlagUseNodePriority : 0);
                    cudaGraphDestroy(g);
                    if (e != cudaSuccess) {
                        c->cur = cur_before;
                        return fail(MCRG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
                    }
                    it = c->graphs.emplace(key, ge).first;
                    // the capture advanced c->cur exactly as a replay will
                } else {
                    // replay flips the ping-pong index as often as the capture did
                    int flips_per_sample = m > 0 ? 1 : 0;
                    if (m > 1) flips_per_sample += (m - 1 + c->fuse_sweeps - 1) / c->fuse_sweeps;
                    if (c->update_mode == MCRG_UPDATE_CLUSTER) flips_per_sample = 0;  // cluster updates work in place
                    if ((chunk * flips_per_sample) & 1) c->cur ^= 1;
                    c->last_levels = n_lv;
                    c->last_parity = (chunk - 1) % c->n_slots;
                    c->measured = true;
                }
                CK(cudaGraphLaunch(it->second, c->stream));
                c->t_host += (unsigned long long)chunk * m;
                done += chunk;
            }
        }
    }
    const int rest = n_samples - done;
    if (rest > 0) {
        for (int s = 0; s < rest; ++s) enqueue_sample(c, n_lv, m, 1, bin, (unsigned long long)s * m, s % c->n_slots);
        join_pyramids(c);
        if (m > 0) launch_advance_t(c->d_t, (unsigned long long)rest * m, c->stream);
        c->t_host += (unsigned long long)rest * m;
    }
    CK(cudaGetLastError());
    return 0;
}

int mcrg_profile_kernels(mcrg_ctx *c, int n_samples, int sweeps_per_sample, int max_levels, float *out_ms) {
    if (!c || !out_ms) return fail(MCRG_ERR_ARG, "null pointer");
    if (n_samples < 1 || n_samples > 4096 || sweeps_per_sample < 0) return fail(MCRG_ERR_ARG, "bad counts");
    CK(cudaSetDevice(c->device));
    const int n_lv = clamp_levels(c, max_levels);
    const int m = sweeps_per_sample;
    std::vector<cudaEvent_t> ev((size_t)n_samples * 5);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    for (int s = 0; s < n_samples; ++s) enqueue_sample(c, n_lv, m, 1, 0, (unsigned long long)s * m, s % c->n_slots, &ev[(size_t)s * 5]);
    if (m > 0) launch_advance_t(c->d_t, (unsigned long long)n_samples * m, c->stream);
    c->t_host += (unsigned long long)n_samples * m;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    double acc[4] = {0, 0, 0, 0};
    for (int s = 0; s < n_samples; ++s)
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev[(size_t)s * 5 + k], ev[(size_t)s * 5 + k + 1]));
            acc[k] += ms;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    out_ms[0] = (float)(acc[0] / n_samples);  // k_sweep0<true>
    out_ms[1] = (float)(acc[3] / n_samples);  // k_sweep0<false> launches after the measurement
    out_ms[2] = (float)(acc[1] / n_samples);  // k_level launches
    out_ms[3] = (float)(acc[2] / n_samples);  // k_tail
    return 0;
}

int mcrg_probe_philox_rate(mcrg_ctx *c, double *calls_per_s) {
    if (!c || !calls_per_s) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (probe_philox_rate(c->stream, calls_per_s) != 0) return fail(MCRG_ERR_CUDA, "Philox throughput probe failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int mcrg_rgnn_set_weights(mcrg_ctx *c, const double *W) {
    if (!c || !W) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->rgnn_W, W, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // W may be a temporary on the caller's side
    return 0;
}

int mcrg_rgnn_eval(mcrg_ctx *c, double h, double *u, double *grad) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (!(h > 0.0)) return fail(MCRG_ERR_ARG, "h must be positive");
    CK(cudaSetDevice(c->device));
    launch_rgnn(c->planes[c->cur], c->L, c->n_replicas, c->rgnn_W, h, c->rgnn_u, c->rgnn_grad, c->rgnn_acc, 0, c->stream);
    CK(cudaGetLastError());
    if (u) CK(cudaMemcpyAsync(u, c->rgnn_u, (size_t)c->n_replicas * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (grad) CK(cudaMemcpyAsync(grad, c->rgnn_grad, (size_t)c->n_replicas * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mcrg_rgnn_run(mcrg_ctx *c, int n_samples, int sweeps_per_sample, double h) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    if (n_samples < 0 || sweeps_per_sample < 0) return fail(MCRG_ERR_ARG, "negative count");
    if (!(h > 0.0)) return fail(MCRG_ERR_ARG, "h must be positive");
    CK(cudaSetDevice(c->device));
    for (int s = 0; s < n_samples; ++s) {
        if (sweeps_per_sample > 0) enqueue_updates(c, sweeps_per_sample, (unsigned long long)s * sweeps_per_sample);
        launch_rgnn(c->planes[c->cur], c->L, c->n_replicas, c->rgnn_W, h, nullptr, nullptr, c->rgnn_acc, 1, c->stream);
    }
    if (n_samples > 0 && sweeps_per_sample > 0) {
        launch_advance_t(c->d_t, (unsigned long long)n_samples * sweeps_per_sample, c->stream);
        c->t_host += (unsigned long long)n_samples * sweeps_per_sample;
    }
    CK(cudaGetLastError());
    return 0;
}

int mcrg_rgnn_accumulators_reset(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->rgnn_acc, 0, (size_t)c->n_replicas * 6 * sizeof(double), c->stream));
    return 0;
}

int mcrg_rgnn_accumulators_get(mcrg_ctx *c, double *out) {
    if (!c || !out) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->rgnn_acc, (size_t)c->n_replicas * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mcrg_accumulators_layout(mcrg_acc_layout *o) {
    if (!o) return fail(MCRG_ERR_ARG, "null pointer");
    o->n_slots = N_SLOTS;
    o->slot_n = SLOT_N;
    o->slot_absm = SLOT_ABSM;
    o->slot_m2 = SLOT_M2;
    o->slot_s = SLOT_S;
    o->slot_ss = SLOT_SS;
    o->slot_sbs = SLOT_SBS;
    o->slot_sb0 = SLOT_SB0;
    o->slot_m4 = SLOT_M4;
    return 0;
}

int mcrg_accumulators_reset(mcrg_ctx *c) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const size_t n_acc = (size_t)c->n_replicas * c->n_bins * N_SLOTS;
    CK(cudaMemsetAsync(c->acc_lo, 0, n_acc * 8, c->stream));
    CK(cudaMemsetAsync(c->acc_hi, 0, n_acc * 8, c->stream));
    return 0;
}

int mcrg_accumulators_get(mcrg_ctx *c, int64_t *hi, uint64_t *lo) {
    if (!c) return fail(MCRG_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const size_t n_acc = (size_t)c->n_replicas * c->n_bins * N_SLOTS;
    if (hi) CK(cudaMemcpyAsync(hi, c->acc_hi, n_acc * 8, cudaMemcpyDeviceToHost, c->stream));
    if (lo) CK(cudaMemcpyAsync(lo, c->acc_lo, n_acc * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mcrg_accumulators_total_limbs_device(mcrg_ctx *c, void *dev_out) {
    if (!c || !dev_out) return fail(MCRG_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    launch_total_limbs(c->acc_lo, c->acc_hi, c->n_replicas * c->n_bins, (long long *)dev_out, c->stream);
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
