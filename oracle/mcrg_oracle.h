/* CPU oracle for the MCRG hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (mcrg_b200/) never does.  Two kinds of function live here:
 *
 *  (R) restatements of the reference algorithm, each citing the reference file:line it follows.  They are
 *      pinned against the UNMODIFIED reference (oracle/_ref/libmcrg_ref.so, built from /root/reference/src)
 *      by tests/test_oracle_vs_ref.py and against the committed fixtures in tests/golden/ (generated from
 *      that same library by tests/golden/make_golden.py).
 *  (S) the scalar specification of what the reference does NOT contain and the north_star adds: the
 *      checkerboard Metropolis sweep, the Philox4x32-10 keying of every random decision, the plaquette
 *      operator.  The reference's sampler is Wolff (ising.cpp:87-155) with a process-global mt19937_64
 *      (definitions.cpp:3-4), so there is nothing to pin these to bit-for-bit: parity for (S) is
 *      "CUDA == this scalar code on the same keys" plus 3-sigma statistics against the reference's sampler.
 *
 * Spin arrays are in the reference's layout: int32 +-1, column-major, element (i,j) at j*N+i
 * (definitions.hpp:16, ising.cpp:117-118).
 */
#ifndef MCRG_ORACLE_H
#define MCRG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (S) Philox4x32-10 and the keying conventions shared with the CUDA kernels ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
enum { ORC_PURPOSE_MC = 1, ORC_PURPOSE_TIE = 2, ORC_PURPOSE_INIT = 3, ORC_PURPOSE_SW_BOND = 4, ORC_PURPOSE_SW_FLIP = 5 };
/* counter = (word, replica, t_lo, purpose<<28 | j<<20 | t_hi(20 bits)); key = (seed_lo, seed_hi) */
void orc_philox_keyed(uint64_t seed, uint32_t word, uint32_t replica, uint64_t t, int purpose, int j,
                      uint32_t out[4]);
/* 2^32-scaled acceptance thresholds for |K|: T4 = floor(exp(-4|K|) 2^32), T8 = floor(exp(-8|K|) 2^32), clamped */
void orc_thresholds(double K, uint32_t *T4, uint32_t *T8);

/* ---- (R) deterministic observables ---- */
/* lattice.cpp:102-120 — out = {S_nn, S_nnn}, double-counted neighbour sums, exact integers */
void orc_calc_interactions(int N, const int32_t *spins, int64_t out[2]);
/* lattice.cpp:84-99 */
int64_t orc_calc_nn(int N, const int32_t *spins);
/* ising.cpp:158-173 — term-by-term double accumulation of K*s*s', divided by N*N */
double orc_calc_energy(int N, const int32_t *spins, double K);
/* ising.cpp:176-179 — INTEGER division of the spin sum by N*N */
double orc_calc_magnetization(int N, const int32_t *spins);
int64_t orc_sum_spins(int N, const int32_t *spins);
/* (S) four-spin plaquette sum over all N*N unit cells (north_star extension; no reference counterpart) */
int64_t orc_plaquette(int N, const int32_t *spins);

/* ---- (R) block-spin decimation, mcrg.cpp:314-348 ---- */
/* ties (block sum == 0) take tie_spins[jb*Nb+ib] (+-1, same layout as out).  tie_mask (optional) gets 1 at ties. */
void orc_block_spin_supplied(int N, int b, const int32_t *spins, const int32_t *tie_spins, int32_t *out,
                             int32_t *tie_mask);
/* (S) the Philox tie convention: spin chosen for block (ib,jb) of the OUTPUT lattice of size Nb at level `level` */
int32_t orc_tie_spin(uint64_t seed, uint32_t replica, uint64_t t, int level, int Nb, int ib, int jb);
/* b = 2 blocking with Philox ties */
void orc_block_spin_philox(int N, const int32_t *spins, uint64_t seed, uint32_t replica, uint64_t t, int level,
                           int32_t *out);
/* full pyramid: S[(lv)*4 + {0:nn,1:nnn,2:plaq,3:sum}] for lv = 0..n_levels; returns n_levels actually built
 * (min(max_levels, log2(N)-1), mcrg.cpp:43).  level_spins (optional) receives the concatenated blocked lattices. */
int orc_pyramid(int N, const int32_t *spins, uint64_t seed, uint32_t replica, uint64_t t, int max_levels,
                int64_t *S, int32_t *level_spins);

/* ---- (R) accumulation and RG matrix, mcrg.cpp:72-131 ---- */
/* One sample: S is [(n_lv+1)][nop] (doubles, integer valued).  Adds into S_sum [(n_lv+1)*nop],
 * SbS [n_lv*nop*nop], SbSb [n_lv*nop*nop] with the reference's column-major flatten (definitions.cpp:9-19):
 * index beta*nop+alpha holds Sb_alpha * S_beta. */
void orc_accumulate(int n_lv, int nop, const double *S, double *S_sum, double *SbS, double *SbSb);
/* exact variant on integers: products summed in 128 bits; hi/lo limbs returned */
void orc_accumulate_i128(int n_lv, int nop, const int64_t *S, int64_t *S_sum, int64_t *SbS_hi, uint64_t *SbS_lo,
                         int64_t *SbSb_hi, uint64_t *SbSb_lo);
/* mcrg.cpp:106-131: averages -> A = <SbSb>-<Sb><Sb>^T, B = <SbS>-<Sb><S>^T, T = A^-1 B, lambda = largest real
 * eigenvalue part, nu = ln b / ln lambda.  Inputs are SUMS over n_samples. */
void orc_rg_eigenvalues(int n_lv, int nop, double n_samples, int b, const double *S_sum, const double *SbS,
                        const double *SbSb, double *lambdas, double *nus);
/* definitions.cpp:79-87 */
int orc_split_samples(int rank, int n_processes, int n_samples);
/* mcrg.cpp:43 */
int orc_n_transformations(int N, int b);

/* ---- (S) the sampler specification ---- */
void orc_hot_start(int L, uint64_t seed, uint32_t replica, int32_t *spins);
/* n_sweeps full checkerboard sweeps (black = (i+j) even first, then white), starting at sweep counter t0 */
void orc_metropolis(int L, int32_t *spins, double K, uint64_t seed, uint32_t replica, uint64_t t0, int n_sweeps);
/* (S) n Swendsen-Wang cluster updates.  Bond between site (x,y) and its +x / +y neighbour: active iff the bond is
 * satisfied (equal spins for K<0) and U < floor((1-exp(-2|K|)) 2^32), the add probability of ising.cpp:9, with
 * U = element 0 (+x) / 1 (+y) of Philox(word = y*L+x, purpose SW_BOND); every cluster flips iff its coin is set, the
 * coin of a cluster being bit (root & 127) of the 128-bit output of Philox(word = root >> 7, purpose SW_FLIP) with
 * root = smallest y*L+x in the cluster (element (root >> 5) & 3, bit root & 31): one call serves 128 site indices.  Same stationary distribution as the
 * reference's Wolff update (ising.cpp:87-155). */
void orc_swendsen_wang(int L, int32_t *spins, double K, uint64_t seed, uint32_t replica, uint64_t t0, int n_updates);
/* plain scalar Metropolis with a xorshift generator, for CPU timing only (attempts/s baseline) */
double orc_metropolis_timing(int L, double K, int n_sweeps, uint64_t seed);

/* ---- (R) RGNN forward, rgnn.cpp:281-307, and finite-difference gradient, rgnn.cpp:310-339 ---- */
double orc_rgnn_scalar_output(int N, const int32_t *spins, int b, const double *W);
void orc_rgnn_gradient(int N, const int32_t *spins, int b, double *W, double h, double *grad);

#ifdef __cplusplus
}
#endif
#endif
