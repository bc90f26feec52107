/* mcrg_b200 — C ABI of the B200-native MCRG hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA or torch types.  Each entry point names the
 * reference interface it replaces (file:line under the reference's src/).  The reference has no device code and
 * no FFI of its own; the binding a maintainer would add is the set of C++ classes in mcrg_b200/host/ (same class
 * names and signatures as lattice.hpp / ising.hpp / mcrg.hpp), shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative mcrg_status on failure; mcrg_last_error() describes the
 *     last failure of the calling thread.  The reference has no error handling at all.
 *   - one context = one device = a batch of `n_replicas` independent L x L periodic Ising lattices (the
 *     reference runs one chain per MPI rank: mcrg.cpp:42-50); a context is not thread-safe, distinct contexts
 *     may be used from distinct threads.
 *   - spins cross the boundary in the reference's layout: int32 +-1, column-major, element (i,j) at j*L+i
 *     (definitions.hpp:16), replicas concatenated.
 *   - L must be a power of two, 2 <= L <= 16384.
 *   - all work is enqueued on the context's stream; calls that return host data synchronise it.
 *   - the library is CUDA-only: if no device is present every call fails with MCRG_ERR_CUDA.  There is no CPU
 *     path in the product.
 */
#ifndef MCRG_B200_H
#define MCRG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcrg_ctx mcrg_ctx;

enum mcrg_status {
    MCRG_OK = 0,
    MCRG_ERR_ARG = -1,   /* bad argument (size not a power of two, index out of range, ...) */
    MCRG_ERR_CUDA = -2,  /* a CUDA runtime call failed (including "no device") */
    MCRG_ERR_STATE = -3  /* call not valid in the current state */
};

/* operators measured at every blocking level, in this order */
enum { MCRG_OP_NN = 0, MCRG_OP_NNN = 1, MCRG_OP_PLAQ = 2, MCRG_OP_SUM = 3, MCRG_NOBS = 4 };
#define MCRG_MAX_LEVELS 15
#define MCRG_NOP 3 /* even operators entering the RG matrix: NN, NNN, PLAQ (the reference uses the first two) */

const char *mcrg_last_error(void);
int mcrg_device_count(int *n);
int mcrg_version(void);

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* Replaces `new Lattice(N)` per rank (lattice.cpp:3-15, mcrg.cpp:49).  replica_base = global id of replica 0
 * (enters every Philox counter, so results do not depend on how replicas are spread over devices).
 * n_bins >= 1 accumulator bins per replica.  The lattices start all-up; call mcrg_init_hot for the reference's
 * hot start. */
int mcrg_ctx_create(int device, int L, int n_replicas, uint64_t seed, uint32_t replica_base, int n_bins,
                    mcrg_ctx **out);
int mcrg_ctx_destroy(mcrg_ctx *ctx);
int mcrg_sync(mcrg_ctx *ctx);
/* the cudaStream_t the context enqueues on, as an integer handle (for event timing by the caller) */
uint64_t mcrg_stream_handle(mcrg_ctx *ctx);
/* device-side timing on the context's stream */
int mcrg_timer_start(mcrg_ctx *ctx);
int mcrg_timer_stop(mcrg_ctx *ctx, float *ms);
int mcrg_levels_full(int L); /* floor(log L / log 2) - 1, mcrg.cpp:43 */
/* tuning knobs: rows per strip of the sweep kernel (0 = heuristic; lattices up to 512^2 then use the resident kernel,
 * a non-zero value forces the strip kernel), sweeps fused per launch, CUDA-graph use */
int mcrg_set_tuning(mcrg_ctx *ctx, int strip_rows, int fuse_sweeps, int use_graphs);
/* Introspection of the strip-height choice (profiles/strip_scan.py, tests): with strip_rows = 0 returns the height the library
 * picks for launches that fuse `fused_sweeps` sweeps, else evaluates the given height; *cost = the launch-time estimate the
 * choice minimises (row steps of one SM; -1 if the strip does not fit in shared memory).  No device work. */
int mcrg_strip_plan(mcrg_ctx *ctx, int fused_sweeps, int strip_rows, int *rows_out, double *cost_out);

/* Which Markov-chain update mcrg_sweep / mcrg_run / mcrg_rgnn_run apply (one sweep-counter tick each):
 *   MCRG_UPDATE_METROPOLIS  one full checkerboard Metropolis sweep (default; the north-star hot path);
 *   MCRG_UPDATE_CLUSTER     one Swendsen-Wang cluster update: bonds between equal neighbours are activated with
 *                           probability 1 - exp(-2|K|) — the add probability of IsingModel::IsingModel, ising.cpp:9 —
 *                           and every cluster flips with probability 1/2.  Same family and same stationary distribution
 *                           as the reference's Wolff update (grow_cluster, ising.cpp:96-149); no critical slowing down.
 * Selecting the cluster update allocates 4 bytes per site for the union-find forest. */
enum { MCRG_UPDATE_METROPOLIS = 0, MCRG_UPDATE_CLUSTER = 1 };
int mcrg_set_update(mcrg_ctx *ctx, int mode);

/* ---- state ---------------------------------------------------------------------------------------------- */
/* IsingModel(K) (ising.cpp:3-11): n = 1 (all replicas) or n = n_replicas.  K < 0 is ferromagnetic. */
int mcrg_set_couplings(mcrg_ctx *ctx, const double *K, int n);
/* Lattice::initialize_random_spins (lattice.cpp:33-41): i.i.d. fair spins, here from Philox(seed, replica) */
int mcrg_init_hot(mcrg_ctx *ctx);
int mcrg_init_cold(mcrg_ctx *ctx);
/* Lattice(int a, imat spins) (lattice.cpp:18-30) / reading Lattice::spins_ (lattice.hpp:16) */
int mcrg_set_spins_i32_colmajor(mcrg_ctx *ctx, int first, int count, const int32_t *host_spins);
int mcrg_get_spins_i32_colmajor(mcrg_ctx *ctx, int first, int count, int32_t *host_spins);
/* Pipelined form of mcrg_set_spins_i32_colmajor for drivers that stream configurations through the device:
 * _begin starts copying `count` replicas from PINNED host memory to a device buffer on a separate copy stream and
 * returns at once — the copy overlaps whatever the context's stream is running; _commit makes the context's stream
 * wait for that copy and packs it into the lattices (replacing replicas [first, first+count)).  One upload of each kind
 * (int32, host-packed) may be in flight per context — for disjoint replica ranges — and one _commit takes in both; the
 * host buffer may be reused after _commit and mcrg_sync. */
int mcrg_set_spins_i32_colmajor_begin(mcrg_ctx *ctx, int first, int count, const int32_t *pinned_host_spins);
/* the same pipeline for configurations packed on the host first (mcrg_host_pack_i32_colmajor): 32x fewer PCIe bytes */
int mcrg_set_spins_packed_begin(mcrg_ctx *ctx, int first, int count, const uint32_t *pinned_host_packed);
int mcrg_set_spins_commit(mcrg_ctx *ctx);
/* block spins produced by the last measurement; level in 1..levels of that measurement; (L>>level)^2 ints */
/* The same upload with the 32-fold smaller PCIe transfer: pack on the host (plain CPU code on `n_threads` threads, no
 * device work; 1 bit per spin, replica-major, internal row y = reference column j, max(1, L/32) words per row, bit k
 * of word w = spin (i = 32w+k, j = y) is +1), then hand the packed words to the device.  mcrg_set_spins_packed is
 * stream-ordered (asynchronous for pinned `packed`); the buffer must stay valid until the next synchronising call. */
size_t mcrg_packed_words(int L, int count);
int mcrg_host_pack_i32_colmajor(const int32_t *host_spins, int L, int count, uint32_t *packed, int n_threads);
int mcrg_set_spins_packed(mcrg_ctx *ctx, int first, int count, const uint32_t *packed);
/* Measurement aid for the packing above (bench.py `e2e.host_roofline`): streams `n_ints` int32 through `n_threads` host
 * threads with no conversion and returns their sum; the caller times it.  No device work. */
int mcrg_host_read_probe(const int32_t *buf, size_t n_ints, int n_threads, int64_t *sum_out);
int mcrg_get_level_spins_i32_colmajor(mcrg_ctx *ctx, int replica, int level, int32_t *host_spins);
int mcrg_get_sweep_counter(mcrg_ctx *ctx, uint64_t *t);
int mcrg_set_sweep_counter(mcrg_ctx *ctx, uint64_t t);

/* ---- the hot path --------------------------------------------------------------------------------------- */
/* IsingModel::sample_new_configuration x n (ising.cpp:87-93) and IsingModel::equilibrate(.., n, false)
 * (ising.cpp:14-84): n full checkerboard Metropolis sweeps of every replica. */
int mcrg_sweep(mcrg_ctx *ctx, int n_sweeps);
/* Lattice::calc_interactions at every blocking level of the current configurations (lattice.cpp:102-120 after
 * repeated block_spin_transformation, mcrg.cpp:314-348; ties by Philox keyed with the sweep counter).
 * max_levels < 0: full pyramid.  S (optional) receives [replica][n_lv+1][MCRG_NOBS]; *n_lv_out the level count. */
int mcrg_measure(mcrg_ctx *ctx, int max_levels, int64_t *S, int *n_lv_out);
/* level-0 sums per replica: calc_interactions, calc_nearest_neighbor_interaction (lattice.cpp:84-120), the
 * spin sum behind calc_magnetization (ising.cpp:176-179) and the plaquette sum.  Any pointer may be NULL. */
/* tie_mode = supplied ("teacher forcing" through the ABI): the same measurement with the CALLER's tie coins
 * instead of the Philox ones, so that the device's block spins can be compared bit for bit with lattices whose ties were
 * drawn elsewhere — e.g. by the reference's own rng (mcrg.cpp:333-335).  tie_bits: per replica mcrg_tie_words(L, max_levels)
 * words = the packed coin words of levels 1, 2, .. laid end to end, each in the natural layout of that output lattice
 * (row = reference column jb, bit = reference row ib, max(1, Ln/32) words per row; bit 1 = the tied block becomes +1;
 * bits of blocks that do not tie are ignored). */
size_t mcrg_tie_words(int L, int max_levels);
int mcrg_measure_supplied(mcrg_ctx *ctx, int max_levels, const uint32_t *tie_bits, int64_t *S, int *n_levels_out);
int mcrg_observables(mcrg_ctx *ctx, int64_t *Snn, int64_t *Snnn, int64_t *Splaq, int64_t *M);
/* the sample loop of calc_critical_exponent (mcrg.cpp:72-98), n_samples times for every replica:
 *   measure the current configuration at all levels, add S, S(n) x S(n-1), S(n) x S(n) into bin `bin`,
 *   then sweeps_per_sample Metropolis sweeps. */
int mcrg_run(mcrg_ctx *ctx, int n_samples, int sweeps_per_sample, int max_levels, int bin);

/* Measurement aid: runs n_samples samples of mcrg_run (bin 0) with an event between the kernels and returns the
 * average device time in ms per sample of {measure+first-sweep kernel, further sweep kernels, level kernels,
 * tail kernel}.  The chain and the accumulators advance exactly as in mcrg_run. */
int mcrg_profile_kernels(mcrg_ctx *ctx, int n_samples, int sweeps_per_sample, int max_levels, float out_ms[4]);

/* Measurement aid for the roofline note: the measured rate (calls per second, whole device) of the arithmetic core of the
 * Metropolis update — Philox4x32-10 + 4-plane lazy threshold compare, two independent calls in flight per thread — i.e. the
 * instruction-issue ceiling a sweep (about 2.1 calls per 32 sites) can approach on this GPU at its current clocks. */
int mcrg_probe_philox_rate(mcrg_ctx *ctx, double *calls_per_s);

/* ---- RGNN: consumer of the sampler ---------------------------------------------------------- */
/* RenormalizationGroupNeuralNetwork::set_weights (rgnn.cpp:38-42): W is the 2x2 filter, column-major (W[k*2+r] = W(r,k)) */
int mcrg_rgnn_set_weights(mcrg_ctx *ctx, const double *W);
/* scalar_output (rgnn.cpp:281-307) and calc_gradient_scalar_output (rgnn.cpp:310-339, central differences with step
 * h) of every replica's current configuration: u[replica], grad[replica][4] (column-major 2x2).  Either may be NULL. */
int mcrg_rgnn_eval(mcrg_ctx *ctx, double h, double *u, double *grad);
/* the sample loop of train_scalar_output for one lattice size (rgnn.cpp:106-123): n_samples x { sweeps_per_sample
 * sweeps, then u and grad of the new configuration added to this replica's sums }.
 * sums per replica: { sum u, sum u^2, sum grad(0,0), grad(1,0), grad(0,1), grad(1,1) }. */
int mcrg_rgnn_run(mcrg_ctx *ctx, int n_samples, int sweeps_per_sample, double h);
int mcrg_rgnn_accumulators_reset(mcrg_ctx *ctx);
int mcrg_rgnn_accumulators_get(mcrg_ctx *ctx, double *out /* [replica][6] */);

/* ---- accumulators (mcrg.cpp:53-70 containers), exact 128-bit integers ------------------------------------ */
typedef struct {
    int n_slots;
    int slot_n, slot_absm, slot_m2; /* samples, sum |M|, sum M^2 */
    int slot_s;   /* + lv*MCRG_NOP + op                sum S^(lv)_op                         */
    int slot_ss;  /* + lv*9 + b*MCRG_NOP + a           sum S^(lv)_a S^(lv)_b   (Sb_Sb of level lv, mcrg.cpp:89) */
    int slot_sbs; /* + (n-1)*9 + b*MCRG_NOP + a        sum S^(n)_a S^(n-1)_b   (Sb_S, mcrg.cpp:88), flatten order
                                                       of definitions.cpp:9-19 */
    int slot_sb0; /* + (n-1)*9 + b*MCRG_NOP + a        sum S^(n)_a S^(0)_b     (blocked level against level 0: the two-
                                                       lattice matching of approx_critical_point, mcrg.cpp:262-263) */
    int slot_m4;  /* + {0,1,2}: sum h*h, sum h*l, sum l*l with M^2 = h*2^20 + l (l < 2^20): the EXACT sum of M^4 is
                     HH*2^40 + 2*HL*2^20 + LL (M^4 reaches 2^112 at L = 16384; the Binder cumulant's fourth moment rides
                     the same int64 limb all-reduce as every other slot; ising.cpp:50-53 reduces E^2, M^2 likewise) */
} mcrg_acc_layout;
int mcrg_accumulators_layout(mcrg_acc_layout *out);
int mcrg_accumulators_reset(mcrg_ctx *ctx);
/* hi/lo: [replica][bin][n_slots]; either may be NULL */
int mcrg_accumulators_get(mcrg_ctx *ctx, int64_t *hi, uint64_t *lo);
/* totals over this context's replicas and bins as 32-bit limbs held in int64 (4 limbs per slot, little endian,
 * top limb signed): summing such vectors over ranks with an int64 all-reduce (NCCL ncclInt64/ncclSum) is exact
 * and order independent for up to 2^31 ranks — the collective of mcrg.cpp:101-103.  The vector is written to
 * DEVICE memory `dev_out` (4*n_slots int64) on the context's stream; pass it straight to the all-reduce. */
int mcrg_accumulators_total_limbs_device(mcrg_ctx *ctx, void *dev_out);

/* ---- several GPUs in one process ------------------------------------------------------------------------------
 * Replaces `mpirun -n P` + the three MPI_Allreduce calls of mcrg.cpp:101-103 for C/C++ hosts.  Create one context per
 * device (give each its own replica_base so that the chains differ), then mcrg_comm_init_all once; all the usual calls
 * are asynchronous per context, so the devices work concurrently.  mcrg_allreduce_accumulators forms every device's
 * limb totals, sums them with one ncclAllReduce(ncclInt64, ncclSum) per device (NVLink / NVSwitch) and returns the
 * grand totals over all replicas and bins of all contexts as exact 128-bit integers hi/lo[n_slots] (either may be NULL).
 * NCCL is dlopen'ed at mcrg_comm_init_all; the library does not depend on it otherwise.  (One process per GPU with
 * torch.distributed, as bench.py does, needs none of this: use mcrg_accumulators_total_limbs_device.) */
int mcrg_comm_init_all(int n, mcrg_ctx **ctxs);
int mcrg_allreduce_accumulators(int n, mcrg_ctx **ctxs, int64_t *hi, uint64_t *lo);
int mcrg_comm_destroy_all(int n, mcrg_ctx **ctxs);

#ifdef __cplusplus
}
#endif
#endif
