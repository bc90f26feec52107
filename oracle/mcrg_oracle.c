/* CPU oracle for the MCRG hot path.  TEST INFRASTRUCTURE ONLY — see mcrg_oracle.h for the rules.
 *
 * Written from the behaviour of the reference (file:line cited per function), not from its text: plain C,
 * integer arithmetic wherever the quantity is an integer, no Eigen.
 */
#include "mcrg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------ */
/* (S) Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)          */
/* ------------------------------------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox_keyed(uint64_t seed, uint32_t word, uint32_t replica, uint64_t t, int purpose, int j,
                      uint32_t out[4]) {
    uint32_t ctr[4], key[2];
    ctr[0] = word;
    ctr[1] = replica;
    ctr[2] = (uint32_t)t;
    ctr[3] = ((uint32_t)purpose << 28) | (((uint32_t)j & 0xFFu) << 20) | (uint32_t)((t >> 32) & 0xFFFFFu);
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    orc_philox4x32_10(ctr, key, out);
}

void orc_thresholds(double K, uint32_t *T4, uint32_t *T8) {
    /* Metropolis for weight exp(-K sum s s') (ising.cpp:8-9 sign convention, K<0 ferromagnetic): a flip of s
     * with neighbour sum h is accepted with min(1, exp(2 K s h)); the two non-trivial values are
     * exp(-4|K|) and exp(-8|K|).  Fixed point with 32 fractional bits, floor, clamped to 2^32-1. */
    double a = fabs(K);
    double p4 = floor(exp(-4.0 * a) * 4294967296.0);
    double p8 = floor(exp(-8.0 * a) * 4294967296.0);
    if (p4 > 4294967295.0) p4 = 4294967295.0;
    if (p8 > 4294967295.0) p8 = 4294967295.0;
    *T4 = (uint32_t)p4;
    *T8 = (uint32_t)p8;
}

/* ------------------------------------------------------------------------------------------------ */
/* (R) observables                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

static inline int wrapi(int x, int N) { return ((x % N) + N) % N; } /* lattice.cpp:126 idiom */
static inline int32_t SP(const int32_t *s, int N, int i, int j) { return s[(size_t)j * N + i]; }

void orc_calc_interactions(int N, const int32_t *spins, int64_t out[2]) {
    /* lattice.cpp:102-120 with the neighbour tables of lattice.cpp:124-151 */
    int64_t snn = 0, snnn = 0;
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) {
            int ip = wrapi(i + 1, N), im = wrapi(i - 1, N), jp = wrapi(j + 1, N), jm = wrapi(j - 1, N);
            int32_t c = SP(spins, N, i, j);
            snn += c * SP(spins, N, ip, j) + c * SP(spins, N, im, j) + c * SP(spins, N, i, jp) +
                   c * SP(spins, N, i, jm);
            snnn += c * SP(spins, N, ip, jp) + c * SP(spins, N, im, jp) + c * SP(spins, N, ip, jm) +
                    c * SP(spins, N, im, jm);
        }
    }
    out[0] = snn;
    out[1] = snnn;
}

int64_t orc_calc_nn(int N, const int32_t *spins) {
    /* lattice.cpp:84-99 */
    int64_t o[2];
    orc_calc_interactions(N, spins, o);
    return o[0];
}

double orc_calc_energy(int N, const int32_t *spins, double K) {
    /* ising.cpp:158-173: same visiting order (i outer, j inner, neighbour k = +i, -i, +j, -j) and the same
     * left-to-right product K*s*s' accumulated in a double, so the rounding sequence is the reference's. */
    double E = 0.0;
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) {
            int ni[4] = {wrapi(i + 1, N), wrapi(i - 1, N), i, i};
            int nj[4] = {j, j, wrapi(j + 1, N), wrapi(j - 1, N)};
            for (int k = 0; k < 4; ++k) E += K * SP(spins, N, i, j) * SP(spins, N, ni[k], nj[k]);
        }
    }
    return E / (N * N);
}

int64_t orc_sum_spins(int N, const int32_t *spins) {
    int64_t m = 0;
    for (size_t k = 0; k < (size_t)N * N; ++k) m += spins[k];
    return m;
}

double orc_calc_magnetization(int N, const int32_t *spins) {
    /* ising.cpp:176-179: int sum / int N*N, truncating toward zero, then widened to double */
    int sum = (int)orc_sum_spins(N, spins);
    return (double)(sum / (N * N));
}

int64_t orc_plaquette(int N, const int32_t *spins) {
    int64_t p = 0;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            int ip = wrapi(i + 1, N), jp = wrapi(j + 1, N);
            p += SP(spins, N, i, j) * SP(spins, N, ip, j) * SP(spins, N, i, jp) * SP(spins, N, ip, jp);
        }
    return p;
}

/* ------------------------------------------------------------------------------------------------ */
/* (R) block-spin decimation                                                                          */
/* ------------------------------------------------------------------------------------------------ */

void orc_block_spin_supplied(int N, int b, const int32_t *spins, const int32_t *tie_spins, int32_t *out,
                             int32_t *tie_mask) {
    /* mcrg.cpp:314-348: Nb = N/b (truncating), block (ib,jb) = rows ib*b.., cols jb*b..; sign of the block
     * sum; a zero sum takes a coin — here the caller's coin for that block. */
    int Nb = N / b;
    for (int ib = 0; ib < Nb; ++ib) {
        for (int jb = 0; jb < Nb; ++jb) {
            int tot = 0;
            for (int i = ib * b; i < (ib + 1) * b; ++i)
                for (int j = jb * b; j < (jb + 1) * b; ++j) tot += SP(spins, N, i, j);
            size_t o = (size_t)jb * Nb + ib;
            if (tie_mask) tie_mask[o] = (tot == 0);
            if (tot == 0) out[o] = tie_spins[o];
            else out[o] = tot > 0 ? 1 : -1;
        }
    }
}

int32_t orc_tie_spin(uint64_t seed, uint32_t replica, uint64_t t, int level, int Nb, int ib, int jb) {
    /* natural packed layout of the output lattice: row = jb, bit = ib, Wb words of 32 bits per row; the 32 coins of
     * word q = jb*Wb + ib/32 are element (q>>8)&3 of the Philox call whose counter word is q with bits 8-9 removed */
    int Wb = Nb >= 32 ? Nb / 32 : 1;
    uint32_t q = (uint32_t)jb * (uint32_t)Wb + ((uint32_t)ib >> 5);
    uint32_t group = ((q >> 10) << 8) | (q & 255u);
    uint32_t r[4];
    orc_philox_keyed(seed, group, replica, t, ORC_PURPOSE_TIE, level, r);
    return ((r[(q >> 8) & 3u] >> (ib & 31)) & 1u) ? 1 : -1;
}

void orc_block_spin_philox(int N, const int32_t *spins, uint64_t seed, uint32_t replica, uint64_t t, int level,
                           int32_t *out) {
    int Nb = N / 2;
    int32_t *ties = (int32_t *)malloc(sizeof(int32_t) * (size_t)Nb * Nb);
    for (int jb = 0; jb < Nb; ++jb)
        for (int ib = 0; ib < Nb; ++ib) ties[(size_t)jb * Nb + ib] = orc_tie_spin(seed, replica, t, level, Nb, ib, jb);
    orc_block_spin_supplied(N, 2, spins, ties, out, NULL);
    free(ties);
}

int orc_n_transformations(int N, int b) {
    /* mcrg.cpp:43 */
    return (int)floor(log((double)N) / log((double)b)) - 1;
}

int orc_pyramid(int N, const int32_t *spins, uint64_t seed, uint32_t replica, uint64_t t, int max_levels,
                int64_t *S, int32_t *level_spins) {
    int n_lv = orc_n_transformations(N, 2);
    if (max_levels >= 0 && max_levels < n_lv) n_lv = max_levels;
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * N);
    int32_t *nxt = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * N);
    memcpy(cur, spins, sizeof(int32_t) * (size_t)N * N);
    int n = N;
    size_t off = 0;
    for (int lv = 0; lv <= n_lv; ++lv) {
        int64_t o[2];
        orc_calc_interactions(n, cur, o);
        S[lv * 4 + 0] = o[0];
        S[lv * 4 + 1] = o[1];
        S[lv * 4 + 2] = orc_plaquette(n, cur);
        S[lv * 4 + 3] = orc_sum_spins(n, cur);
        if (lv == n_lv) break;
        orc_block_spin_philox(n, cur, seed, replica, t, lv + 1, nxt);
        n /= 2;
        if (level_spins) {
            memcpy(level_spins + off, nxt, sizeof(int32_t) * (size_t)n * n);
            off += (size_t)n * n;
        }
        int32_t *tmp = cur; cur = nxt; nxt = tmp;
    }
    free(cur);
    free(nxt);
    return n_lv;
}

/* ------------------------------------------------------------------------------------------------ */
/* (R) accumulation and RG matrix                                                                     */
/* ------------------------------------------------------------------------------------------------ */

void orc_accumulate(int n_lv, int nop, const double *S, double *S_sum, double *SbS, double *SbSb) {
    /* mcrg.cpp:80-97.  Sb = S.row(n), S = S.row(n-1); Sb_S = Sb * S^T; flatten is column-major
     * (definitions.cpp:9-19): v[beta*nop + alpha] = Sb[alpha]*S[beta]. */
    for (int n = 1; n <= n_lv; ++n) {
        const double *Sb = S + (size_t)n * nop;
        const double *Sp = S + (size_t)(n - 1) * nop;
        for (int beta = 0; beta < nop; ++beta)
            for (int alpha = 0; alpha < nop; ++alpha) {
                SbS[(size_t)(n - 1) * nop * nop + beta * nop + alpha] += Sb[alpha] * Sp[beta];
                SbSb[(size_t)(n - 1) * nop * nop + beta * nop + alpha] += Sb[alpha] * Sb[beta];
            }
    }
    for (int k = 0; k < (n_lv + 1) * nop; ++k) S_sum[k] += S[k];
}

static void add128(int64_t *hi, uint64_t *lo, __int128 v) {
    __int128 cur = ((__int128)*hi << 64) | (__int128)*lo;
    cur += v;
    *hi = (int64_t)(cur >> 64);
    *lo = (uint64_t)cur;
}

void orc_accumulate_i128(int n_lv, int nop, const int64_t *S, int64_t *S_sum, int64_t *SbS_hi, uint64_t *SbS_lo,
                         int64_t *SbSb_hi, uint64_t *SbSb_lo) {
    for (int n = 1; n <= n_lv; ++n) {
        const int64_t *Sb = S + (size_t)n * nop;
        const int64_t *Sp = S + (size_t)(n - 1) * nop;
        for (int beta = 0; beta < nop; ++beta)
            for (int alpha = 0; alpha < nop; ++alpha) {
                size_t k = (size_t)(n - 1) * nop * nop + beta * nop + alpha;
                add128(&SbS_hi[k], &SbS_lo[k], (__int128)Sb[alpha] * Sp[beta]);
                add128(&SbSb_hi[k], &SbSb_lo[k], (__int128)Sb[alpha] * Sb[beta]);
            }
    }
    for (int k = 0; k < (n_lv + 1) * nop; ++k) S_sum[k] += S[k];
}

/* small dense helpers (nop <= 3), column-major like the reference's unflatten (definitions.cpp:21-31) */
static int invert_small(int n, const double *A, double *Ai) {
    double m[3][6];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            m[r][c] = A[c * n + r];
            m[r][n + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (fabs(m[r][col]) > fabs(m[piv][col])) piv = r;
        if (m[piv][col] == 0.0) return -1;
        if (piv != col)
            for (int c = 0; c < 2 * n; ++c) { double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
        double d = m[col][col];
        for (int c = 0; c < 2 * n; ++c) m[col][c] /= d;
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            double f = m[r][col];
            for (int c = 0; c < 2 * n; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) Ai[c * n + r] = m[r][n + c];
    return 0;
}

/* real parts of the eigenvalues of a small real matrix (column-major); returns the largest */
static double largest_real_eigenvalue(int n, const double *T) {
    if (n == 1) return T[0];
    if (n == 2) {
        double a = T[0], c = T[1], b = T[2], d = T[3]; /* [a b; c d] */
        double tr = a + d, det = a * d - b * c;
        double disc = tr * tr / 4.0 - det;
        if (disc < 0.0) return tr / 2.0; /* complex pair: both real parts tr/2 (EigenSolver .real()) */
        return tr / 2.0 + sqrt(disc);
    }
    /* n == 3: lambda^3 - c2 lambda^2 + c1 lambda - c0 = 0 */
    double a11 = T[0], a21 = T[1], a31 = T[2], a12 = T[3], a22 = T[4], a32 = T[5], a13 = T[6], a23 = T[7], a33 = T[8];
    double c2 = a11 + a22 + a33;
    double c1 = a11 * a22 - a12 * a21 + a11 * a33 - a13 * a31 + a22 * a33 - a23 * a32;
    double c0 = a11 * (a22 * a33 - a23 * a32) - a12 * (a21 * a33 - a23 * a31) + a13 * (a21 * a32 - a22 * a31);
    /* depressed cubic x = lambda - c2/3: x^3 + p x + q = 0 */
    double sh = c2 / 3.0;
    double p = c1 - c2 * c2 / 3.0;
    double q = -2.0 * c2 * c2 * c2 / 27.0 + c2 * c1 / 3.0 - c0;
    double disc = q * q / 4.0 + p * p * p / 27.0;
    if (disc > 0.0) {
        double sq = sqrt(disc);
        double u = cbrt(-q / 2.0 + sq), v = cbrt(-q / 2.0 - sq);
        double real_root = u + v + sh;
        double pair_re = -(u + v) / 2.0 + sh;
        return real_root > pair_re ? real_root : pair_re;
    }
    double rr = 2.0 * sqrt(-p / 3.0);
    double arg = (p == 0.0) ? 0.0 : (3.0 * q / (p * rr));
    if (arg > 1.0) arg = 1.0;
    if (arg < -1.0) arg = -1.0;
    double phi = acos(arg) / 3.0;
    double best = -INFINITY;
    for (int k = 0; k < 3; ++k) {
        double x = rr * cos(phi - 2.0 * M_PI * k / 3.0) + sh;
        if (x > best) best = x;
    }
    return best;
}

void orc_rg_eigenvalues(int n_lv, int nop, double n_samples, int b, const double *S_sum, const double *SbS,
                        const double *SbSb, double *lambdas, double *nus) {
    /* mcrg.cpp:106-131 */
    for (int n = 0; n < n_lv; ++n) {
        double A[9], B[9], Ai[9], T[9];
        const double *Sb = S_sum + (size_t)(n + 1) * nop;
        const double *Sp = S_sum + (size_t)n * nop;
        for (int beta = 0; beta < nop; ++beta)
            for (int alpha = 0; alpha < nop; ++alpha) {
                size_t k = (size_t)n * nop * nop + beta * nop + alpha;
                double sb_a = Sb[alpha] / n_samples, sb_b = Sb[beta] / n_samples, s_b = Sp[beta] / n_samples;
                B[beta * nop + alpha] = SbS[k] / n_samples - sb_a * s_b;   /* dSb_dK,  mcrg.cpp:121 */
                A[beta * nop + alpha] = SbSb[k] / n_samples - sb_a * sb_b; /* dSb_dKb, mcrg.cpp:122 */
            }
        if (invert_small(nop, A, Ai) != 0) {
            lambdas[n] = NAN;
            if (nus) nus[n] = NAN;
            continue;
        }
        for (int c = 0; c < nop; ++c)
            for (int r = 0; r < nop; ++r) {
                double acc = 0.0;
                for (int k = 0; k < nop; ++k) acc += Ai[k * nop + r] * B[c * nop + k];
                T[c * nop + r] = acc; /* T = A^-1 B, mcrg.cpp:123 */
            }
        double lam = largest_real_eigenvalue(nop, T);
        lambdas[n] = lam;
        if (nus) nus[n] = log((double)b) / log(lam); /* mcrg.cpp:131 */
    }
}

int orc_split_samples(int rank, int n_processes, int n_samples) {
    /* definitions.cpp:79-87 */
    int n_loc = (int)ceil((double)n_samples / (double)n_processes);
    if (rank == 0) n_loc = n_samples - (n_processes - 1) * n_loc;
    return n_loc;
}

/* ------------------------------------------------------------------------------------------------ */
/* (S) sampler specification                                                                          */
/* ------------------------------------------------------------------------------------------------ */
/* Internal coordinates: y = reference column j, x = reference row i, so an internal row is contiguous in the
 * reference's column-major array.  Colour of a site = (x+y)&1 (0 = black).  Level-0 state is held as two
 * colour planes; within plane c, row y, the sites are x = 2x'+((y+c)&1), x' = x>>1, packed 32 per word:
 * word = x'>>5, lane = x'&31, W = max(1, L/64) words per row and colour; the Philox "word" coordinate is
 * (c*L + y)*W + word.  */

static inline uint32_t mc_word_id(int L, int c, int y, int xh) {
    int W = L >= 64 ? L / 64 : 1;
    return (uint32_t)(((size_t)c * L + y) * W + (xh >> 5));
}

void orc_hot_start(int L, uint64_t seed, uint32_t replica, int32_t *spins) {
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x) {
            int c = (x + y) & 1, xh = x >> 1;
            uint32_t r[4];
            orc_philox_keyed(seed, mc_word_id(L, c, y, xh), replica, 0, ORC_PURPOSE_INIT, 0, r);
            spins[(size_t)y * L + x] = ((r[0] >> (xh & 31)) & 1u) ? 1 : -1;
        }
}

/* uniform U in [0,2^32) for (site, sweep): bit k (MSB first) is bit `lane` of the k-th Philox output word of
 * that site's packed word, output word k = call (k>>2), element (k&3).  mc_four_bits returns the four bits call j supplies. */
static uint32_t mc_four_bits(uint64_t seed, uint32_t word, int lane, uint32_t replica, uint64_t t, int j) {
    uint32_t r[4], v = 0;
    orc_philox_keyed(seed, word, replica, t, ORC_PURPOSE_MC, j, r);
    for (int e = 0; e < 4; ++e) v = (v << 1) | ((r[e] >> lane) & 1u);
    return v;
}

void orc_metropolis(int L, int32_t *spins, double K, uint64_t seed, uint32_t replica, uint64_t t0, int n_sweeps) {
    uint32_t T4, T8;
    orc_thresholds(K, &T4, &T8);
    int ferro = (K <= 0.0);
    for (int sw = 0; sw < n_sweeps; ++sw) {
        uint64_t t = t0 + (uint64_t)sw;
        for (int c = 0; c < 2; ++c) {
            for (int y = 0; y < L; ++y)
                for (int x = 0; x < L; ++x) {
                    if (((x + y) & 1) != c) continue;
                    int32_t s = spins[(size_t)y * L + x];
                    int yp = (y + 1) % L, ym = (y + L - 1) % L, xp = (x + 1) % L, xm = (x + L - 1) % L;
                    int32_t nb[4] = {spins[(size_t)ym * L + x], spins[(size_t)yp * L + x], spins[(size_t)y * L + xm],
                                     spins[(size_t)y * L + xp]};
                    int A = 0; /* unfavourable bonds that the flip would repair: anti-aligned if ferro */
                    for (int k = 0; k < 4; ++k) A += ferro ? (nb[k] != s) : (nb[k] == s);
                    int flip;
                    if (A >= 2) flip = 1;
                    else {
                        /* flip = U < T with the 32-bit uniform U of this (site, sweep), compared four bits at a time, MSB
                         * first: the first differing group decides, so most sites need one Philox call, not eight */
                        const uint32_t T = (A == 1) ? T4 : T8;
                        const int xh = x >> 1;
                        const uint32_t word = mc_word_id(L, c, y, xh);
                        flip = 0; /* U == T is not an acceptance */
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t u4 = mc_four_bits(seed, word, xh & 31, replica, t, j), t4 = (T >> (28 - 4 * j)) & 15u;
                            if (u4 != t4) {
                                flip = u4 < t4;
                                break;
                            }
                        }
                    }
                    if (flip) spins[(size_t)y * L + x] = -s;
                }
        }
    }
}

static int uf_find(int *parent, int x) {
    while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}

void orc_swendsen_wang(int L, int32_t *spins, double K, uint64_t seed, uint32_t replica, uint64_t t0, int n_updates) {
    const int n = L * L;
    int *parent = (int *)malloc(sizeof(int) * (size_t)n);
    double pd = floor((1.0 - exp(-2.0 * fabs(K))) * 4294967296.0);
    if (pd > 4294967295.0) pd = 4294967295.0;
    const uint32_t TP = (uint32_t)pd;
    const int ferro = (K <= 0.0);
    for (int u = 0; u < n_updates; ++u) {
        const uint64_t t = t0 + (uint64_t)u;
        for (int i = 0; i < n; ++i) parent[i] = i;
        for (int y = 0; y < L; ++y)
            for (int x = 0; x < L; ++x) {
                const int i = y * L + x;
                const int nb[2] = {y * L + (x + 1) % L, ((y + 1) % L) * L + x};
                uint32_t r[4];
                orc_philox_keyed(seed, (uint32_t)i, replica, t, ORC_PURPOSE_SW_BOND, 0, r);
                for (int d = 0; d < 2; ++d) {
                    const int satisfied = ferro ? (spins[i] == spins[nb[d]]) : (spins[i] != spins[nb[d]]);
                    if (satisfied && r[d] < TP) {
                        int a = uf_find(parent, i), b = uf_find(parent, nb[d]);
                        if (a != b) {
                            if (a < b) parent[b] = a;  /* the root is always the smallest index of the cluster */
                            else parent[a] = b;
                        }
                    }
                }
            }
        for (int i = 0; i < n; ++i) {
            const int root = uf_find(parent, i);
            uint32_t r[4];
            orc_philox_keyed(seed, (uint32_t)root >> 7, replica, t, ORC_PURPOSE_SW_FLIP, 0, r);
            if ((r[((uint32_t)root >> 5) & 3u] >> ((uint32_t)root & 31u)) & 1u) spins[i] = -spins[i];
        }
    }
    free(parent);
}

double orc_metropolis_timing(int L, double K, int n_sweeps, uint64_t seed) {
    /* A plain one-spin-per-byte checkerboard Metropolis with a table of acceptance thresholds and xorshift64*;
     * used only to quote a scalar CPU attempts/s figure next to the GPU number.  Returns seconds. */
    int8_t *s = (int8_t *)malloc((size_t)L * L);
    uint64_t st = seed ? seed : 88172645463325252ull;
    for (size_t k = 0; k < (size_t)L * L; ++k) {
        st ^= st >> 12; st ^= st << 25; st ^= st >> 27;
        s[k] = ((st * 2685821657736338717ull) >> 63) ? 1 : -1;
    }
    uint32_t thr[9];
    for (int e = -4; e <= 4; ++e) {
        double p = exp(2.0 * K * e);
        thr[e + 4] = p >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(p * 4294967296.0);
    }
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int sw = 0; sw < n_sweeps; ++sw)
        for (int c = 0; c < 2; ++c)
            for (int y = 0; y < L; ++y) {
                int yp = (y + 1) % L, ym = (y + L - 1) % L;
                for (int x = (y + c) & 1; x < L; x += 2) {
                    int xp = x + 1 == L ? 0 : x + 1, xm = x == 0 ? L - 1 : x - 1;
                    int sh = s[(size_t)y * L + x] *
                             (s[(size_t)ym * L + x] + s[(size_t)yp * L + x] + s[(size_t)y * L + xm] + s[(size_t)y * L + xp]);
                    st ^= st >> 12; st ^= st << 25; st ^= st >> 27;
                    uint32_t u = (uint32_t)((st * 2685821657736338717ull) >> 32);
                    if (u <= thr[sh + 4]) s[(size_t)y * L + x] = -s[(size_t)y * L + x];
                }
            }
    clock_gettime(CLOCK_MONOTONIC, &b);
    volatile int8_t sink = s[0];
    (void)sink;
    free(s);
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------ */
/* (R) RGNN forward and gradient                                                                      */
/* ------------------------------------------------------------------------------------------------ */

double orc_rgnn_scalar_output(int N, const int32_t *spins, int b, const double *W) {
    /* rgnn.cpp:281-307: until 1x1, replace each bxb block B by the entrywise-L1 norm of the MATRIX PRODUCT
     * W*B.  All matrices column-major. */
    int n = N;
    double *cur = (double *)malloc(sizeof(double) * (size_t)N * N);
    double *nxt = (double *)malloc(sizeof(double) * (size_t)N * N);
    for (size_t k = 0; k < (size_t)N * N; ++k) cur[k] = (double)spins[k];
    while (n > 1) {
        int m = n / b;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                double l1 = 0.0;
                for (int c = 0; c < b; ++c)
                    for (int r = 0; r < b; ++r) {
                        double acc = 0.0;
                        for (int k = 0; k < b; ++k) acc += W[k * b + r] * cur[(size_t)(b * j + c) * n + (b * i + k)];
                        l1 += fabs(acc);
                    }
                nxt[(size_t)j * m + i] = l1;
            }
        double *t = cur; cur = nxt; nxt = t;
        n = m;
    }
    double out = cur[0];
    free(cur);
    free(nxt);
    return out;
}

void orc_rgnn_gradient(int N, const int32_t *spins, int b, double *W, double h, double *grad) {
    /* rgnn.cpp:310-339: central differences, one weight at a time, weight restored afterwards */
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) {
            double w = W[j * b + i];
            W[j * b + i] = w + h;
            double o1 = orc_rgnn_scalar_output(N, spins, b, W);
            W[j * b + i] = w - h;
            double o2 = orc_rgnn_scalar_output(N, spins, b, W);
            grad[j * b + i] = (o1 - o2) / (2 * h);
            W[j * b + i] = w;
        }
}
