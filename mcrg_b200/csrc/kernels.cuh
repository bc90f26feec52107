// Kernel-side declarations shared by kernels.cu and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tile.cuh"

namespace mcrg {

#if defined(__CUDACC__)
__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }  // v a power of two
#endif

constexpr int MAX_LEVELS = 15;          // L <= 2^15; levels 0..14 at most
constexpr int NOP = 3;                  // even operators: nn, nnn, plaquette
constexpr int TAIL_MAX_L = 512;         // blocked lattices up to this size CAN finish inside one CTA's shared memory (k_tail's buffers);
                                        // the host starts the tail at 512^2 when there are enough replicas to fill the SMs with
                                        // one CTA each, else at 256^2 (capi.cu: tail_start_size)
constexpr int SWEEP_THREADS = 256;
// capacity of a warp's "still undecided after 8 bit planes" queue for a strip with `words` words per colour
// (expected fill: ~10 % of the words)
// per warp: a fifth of the words a warp handles in one half-sweep, at least 16 entries (16 bytes each)
MCRG_HD int sweep0_queue_cap(int words, int warps) { const int c = (words / warps + 4) / 5; return c > 16 ? c : 16; }

// accumulator slots (per replica, per bin), all exact 128-bit integers (lo: uint64, hi: int64)
constexpr int SLOT_N = 0;                                    // samples
constexpr int SLOT_ABSM = 1;                                 // sum |M|,  M = sum of spins at level 0
constexpr int SLOT_M2 = 2;                                   // sum M^2
constexpr int SLOT_S = 3;                                    // + lv*NOP + op          sum S^(lv)_op
constexpr int SLOT_SS = SLOT_S + NOP * (MAX_LEVELS + 1);     // + lv*9 + b*3 + a        sum S^(lv)_a S^(lv)_b
constexpr int SLOT_SBS = SLOT_SS + NOP * NOP * (MAX_LEVELS + 1);  // + (n-1)*9 + b*3 + a  sum S^(n)_a S^(n-1)_b
constexpr int SLOT_SB0 = SLOT_SBS + NOP * NOP * MAX_LEVELS;       // + (n-1)*9 + b*3 + a  sum S^(n)_a S^(0)_b  (two-lattice
                                                                  //   matching needs S(blocked) x S(level 0), mcrg.cpp:262-263)
// sum M^4, exact: M^2 = h * 2^20 + l (l < 2^20), three slots sum h*h, sum h*l, sum l*l;  sum M^4 = HH * 2^40 + 2 * HL * 2^20 + LL.
// (M^4 reaches 2^112 at L = 16384 — one 128-bit slot would overflow after 2^15 ordered samples — and k_resident sums a launch
// in 64 bits: with L <= 512, h <= 2^16 and every product stays below 2^40 like the correlator products.)
constexpr int SLOT_M4 = SLOT_SB0 + NOP * NOP * MAX_LEVELS;        // + {0: h*h, 1: h*l, 2: l*l}
constexpr int N_SLOTS = SLOT_M4 + 3;
constexpr int M4_SPLIT_BITS = 20;

struct SweepArgs {
    const uint32_t *src;       // planes [replica][colour][y][w]
    uint32_t *dst;
    uint32_t *level1;          // natural layout [replica][y][w] of the L/2 lattice (MEASURE only)
    const uint32_t *ties;      // nullptr: Philox tie coins; else caller-supplied coins [replica][tie_stride] (mcrg_measure_supplied)
    size_t tie_stride;
    unsigned long long *cnt;   // [replica][MAX_LEVELS+1][4] raw popcounts (MEASURE only)
    const uint32_t *T4, *T8, *anti;  // per replica
    const unsigned long long *d_t;   // device-resident sweep counter
    unsigned long long t_off;
    uint64_t seed;
    uint32_t replica_base;
    int L, W, bits;
    int R, H, nsw;
    int strips;                // ceil(L / R): the last strip may be shorter
};

struct LevelArgs {
    const uint32_t *in;        // natural layout [replica][y][w], lattice size Ln
    uint32_t *out;             // natural layout of Ln/2 (nullptr: measure only)
    const uint32_t *ties;      // as in SweepArgs
    size_t tie_stride;
    int L;                     // level-0 size (locates a level inside the supplied coins)
    unsigned long long *cnt;
    const unsigned long long *d_t;
    unsigned long long t_off;
    uint64_t seed;
    uint32_t replica_base;
    int Ln, level;             // level index of `in`
    int R, strips;
};

struct TailArgs {
    const uint32_t *in;        // level `start` lattice, natural layout [replica][y][w] (unused if start > n_levels)
    uint32_t *levels_out;      // base of the level buffers
    const uint32_t *ties;      // as in SweepArgs
    size_t tie_stride;
    const size_t *level_off;   // [MAX_LEVELS+1] word offset of each level's [replica][y][w] block
    unsigned long long *cnt;
    long long *S_out;          // [replica][MAX_LEVELS+1][4] converted sums of the last measurement
    unsigned long long *acc_lo;
    long long *acc_hi;
    const unsigned long long *d_t;
    unsigned long long t_off;
    uint64_t seed;
    uint32_t replica_base;
    int L;                     // level-0 size
    int start;                 // first level handled here (>= 1)
    int n_levels;              // blocking levels of this measurement (levels 0..n_levels exist)
    int n_bins, bin;
    int accumulate;
};

constexpr int RESIDENT_MAX_L = 512;  // a replica up to this size (plus its pyramid) lives in one CTA's shared memory

struct ResidentArgs {
    uint32_t *planes;          // [replica][colour][y][w], updated in place
    const uint32_t *T4, *T8, *anti;
    const unsigned long long *d_t;
    unsigned long long t_off;
    uint64_t seed;
    uint32_t replica_base;
    int L, W, bits;
    int n_samples, m;          // n_samples x { [measure], m sweeps }
    int n_replicas;            // (k_resident_multi: several replicas per CTA)
    int n_levels, accumulate, n_bins, bin;
    unsigned long long *acc_lo;
    long long *acc_hi;
    long long *S_out;
};

// shared-memory carve-up of k_resident, in 32-bit words: planes | queue | bufA | bufB | 128-bit accumulators
struct ResidentLayout {
    int queue_off, bufA_off, bufB_off, acc_off, total_words, cap;
};
MCRG_HD ResidentLayout resident_layout(int L, int threads, int n_levels) {
    ResidentLayout o;
    const int W = l0_words(L), words = (L + 2) * W, warps = threads / 32;
    o.cap = sweep0_queue_cap(words, warps);
    o.queue_off = (2 * words + 3) & ~3;
    o.bufA_off = o.queue_off + 4 * o.cap * warps;
    const int L1 = L / 2 > 0 ? L / 2 : 1, L2 = L / 4 > 0 ? L / 4 : 1;
    o.bufB_off = (o.bufA_off + L1 * nat_words(L1) + 3) & ~3;
    o.acc_off = (o.bufB_off + L2 * nat_words(L2) + 3) & ~3;
    const int n_live = 6 + (NOP + NOP * NOP) * (n_levels + 1) + 2 * NOP * NOP * n_levels;  // = acc_live_slots (kernels.cu)
    o.total_words = o.acc_off + 4 * n_live;
    return o;
}

// Caller-supplied tie coins (mcrg_measure_supplied): per replica the packed coin words of levels 1, 2, .. laid end to end,
// each level in the natural layout of that (output) lattice.  Word offset of level lv >= 1:
MCRG_HD size_t tie_level_off(int L, int lv) {
    size_t off = 0;
    for (int k = 1; k < lv; ++k) off += (size_t)(L >> k) * nat_words(L >> k);
    return off;
}

// Swendsen-Wang cluster update (stands in for the reference's Wolff update, ising.cpp:87-155)
struct SwArgs {
    uint32_t *planes;          // [replica][colour][y][w], flipped in place
    int *parent;               // [replica][L*L] union-find forest, root = smallest site index of the cluster
    uint32_t *coins;           // [replica][ceil(L*L/128)][4] cluster coins, bit i = coin of the cluster rooted at site i
    const uint32_t *TP;        // per replica: floor((1 - exp(-2|K|)) 2^32), the bond probability of ising.cpp:9
    const uint32_t *anti;
    const unsigned long long *d_t;
    unsigned long long t_off;
    uint64_t seed;
    uint32_t replica_base;
    int L, W, bits;
};
void launch_sw_update(const SwArgs &a, int n_replicas, cudaStream_t st);

void launch_resident(const ResidentArgs &a, int n_replicas, bool measure, int forced_threads, cudaStream_t st);
void launch_sweep0(const SweepArgs &a, int n_replicas, bool measure, cudaStream_t st, bool pdl);  // pdl: see launch_pdl (kernels.cu)
void launch_level(const LevelArgs &a, int n_replicas, cudaStream_t st, bool pdl);
void launch_tail(const TailArgs &a, int n_replicas, cudaStream_t st, bool pdl);
void launch_total_limbs(const unsigned long long *lo, const long long *hi, int n_rb, long long *out, cudaStream_t st);
void launch_rgnn(const uint32_t *planes, int L, int n_replicas, const double *W, double h, double *u_out, double *grad_out,
                 double *acc, int accumulate, cudaStream_t st);
void launch_advance_t(unsigned long long *d_t, unsigned long long by, cudaStream_t st);
void launch_init_hot(uint32_t *planes, int L, int n_replicas, uint64_t seed, uint32_t replica_base, cudaStream_t st);
void launch_init_cold(uint32_t *planes, int L, int n_replicas, cudaStream_t st);
void launch_pack0(const int32_t *spins, uint32_t *planes, int L, int n_replicas, cudaStream_t st);
void launch_pack_nat(const uint32_t *nat, uint32_t *planes, int L, int n_replicas, cudaStream_t st);
void launch_unpack0(const uint32_t *planes, int32_t *spins, int L, int n_replicas, cudaStream_t st);
void launch_unpackN(const uint32_t *lev, int32_t *spins, int Ln, int n_replicas, cudaStream_t st);
int probe_philox_rate(cudaStream_t st, double *calls_per_s);
size_t sweep0_smem_bytes(int L, int R, int H);
int sweep0_threads(int L, int R, int H);
int sweep0_max_smem();
int sweep0_occupancy(int L, int R, int H);

}  // namespace mcrg
