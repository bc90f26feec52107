// Tile phases of the MCRG hot path: what one thread does for one word of a strip that sits in fast memory
// (shared memory on the GPU, a plain array in the CPU emulation used by the tests).  The kernels in
// kernels.cu add the parts that only exist on the device: staging, warp/block reductions, atomics, stores.
#pragma once
#include "bitops.cuh"

namespace mcrg {

// raw popcounts; converted to the reference's double-counted sums by counts_to_S()
struct Counts {
    uint32_t anti_nn;   // anti-aligned nearest-neighbour bonds (each bond once)
    uint32_t anti_nnn;  // anti-aligned next-nearest (diagonal) bonds (each bond once)
    uint32_t odd_plaq;  // unit cells whose four-spin product is -1
    uint32_t up;        // up spins
};

// lattice.cpp:102-120 double-counts every bond: S_nn = 2*sum_bonds s s' = 2*(2 n^2 - 2 anti) and likewise S_nnn;
// plaquette sum = n^2 - 2 odd; magnetisation sum = 2 up - n^2.   (n = lattice size at this level)
MCRG_HD void counts_to_S(long long n, unsigned long long anti_nn, unsigned long long anti_nnn,
                         unsigned long long odd_plaq, unsigned long long up, long long S[4]) {
    const long long n2 = n * n;
    S[0] = 4 * n2 - 4 * (long long)anti_nn;
    S[1] = 4 * n2 - 4 * (long long)anti_nnn;
    S[2] = n2 - 2 * (long long)odd_plaq;
    S[3] = 2 * (long long)up - n2;
}

// ---- level 0: colour-separated strip -------------------------------------------------------------------------
struct Strip0 {
    uint32_t *base;   // plane c at base + c*rows*W; word (lr, w) at [lr*W + w], lr = 0 .. rows-1
    int rows;         // R + 2H
    int W, bits;
    uint32_t mask;
    int L;
    int y_first;      // global row of local row 0, in [0, L)
};

MCRG_HD uint32_t *s0_plane(const Strip0 &s, int c) { return s.base + c * s.rows * s.W; }
MCRG_HD uint32_t s0_get(const Strip0 &s, int c, int lr, int w) { return s.base[(c * s.rows + lr) * s.W + w]; }
MCRG_HD uint32_t s0_up(const Strip0 &s, int c, int lr, int w) {  // index x'+1
    const int wn = (w + 1 == s.W) ? 0 : w + 1;
    return shift_up_index(s0_get(s, c, lr, w), s0_get(s, c, lr, wn), s.bits, s.mask);
}
MCRG_HD uint32_t s0_dn(const Strip0 &s, int c, int lr, int w) {  // index x'-1
    const int wp = (w == 0) ? s.W - 1 : w - 1;
    return shift_down_index(s0_get(s, c, lr, w), s0_get(s, c, lr, wp), s.bits, s.mask);
}

// Correlator popcounts of rows lr0 (global y even) and lr0+1, word w, plus the majority/tie words of block row
// y/2 (the four spins of block xb are bit xb of black[y], white[y], black[y+1], white[y+1] — no compaction).
// Needs row lr0+2 in the strip.  Bonds are enumerated as "right and down from every site", diagonals as
// "down-right and down-left from every site", plaquettes by their top-left site.
MCRG_HD void measure_pair0(const Strip0 &s, int lr0, int w, Counts &cnt, uint32_t &maj, uint32_t &tie) {
    const int lr1 = lr0 + 1, lr2 = lr0 + 2;
    const uint32_t b0 = s0_get(s, 0, lr0, w), w0 = s0_get(s, 1, lr0, w);
    const uint32_t b1 = s0_get(s, 0, lr1, w), w1 = s0_get(s, 1, lr1, w);
    const uint32_t b2 = s0_get(s, 0, lr2, w), w2 = s0_get(s, 1, lr2, w);
    const uint32_t b0u = s0_up(s, 0, lr0, w);  // black row y, index x'+1
    const uint32_t w1u = s0_up(s, 1, lr1, w);
    const uint32_t b2u = s0_up(s, 0, lr2, w);
    const uint32_t b1d = s0_dn(s, 0, lr1, w);  // black row y+1, index x'-1
    const uint32_t w2d = s0_dn(s, 1, lr2, w);
    // even row y: black sites x = 2x', white sites x = 2x'+1
    uint32_t nn = popc32(b0 ^ w0) + popc32(w0 ^ b0u) + popc32(b0 ^ w1) + popc32(w0 ^ b1);
    uint32_t nnn = popc32(b0 ^ b1) + popc32(b0 ^ b1d) + popc32(w0 ^ w1u) + popc32(w0 ^ w1);
    uint32_t pq = popc32(b0 ^ w0 ^ w1 ^ b1) + popc32(w0 ^ b0u ^ b1 ^ w1u);
    // odd row y+1: black sites x = 2x'+1, white sites x = 2x'
    nn += popc32(b1 ^ w1u) + popc32(w1 ^ b1) + popc32(b1 ^ w2) + popc32(w1 ^ b2);
    nnn += popc32(b1 ^ b2u) + popc32(b1 ^ b2) + popc32(w1 ^ w2) + popc32(w1 ^ w2d);
    pq += popc32(b1 ^ w1u ^ w2 ^ b2u) + popc32(w1 ^ b1 ^ b2 ^ w2);
    cnt.anti_nn += nn;
    cnt.anti_nnn += nnn;
    cnt.odd_plaq += pq;
    cnt.up += popc32(b0) + popc32(w0) + popc32(b1) + popc32(w1);
    majority4(b0, w0, b1, w1, maj, tie);
}

// One Metropolis word update, in place: colour c, local row lr (needs rows lr-1 and lr+1 of the other colour).
MCRG_HD void update_word0(const Strip0 &s, int c, int lr, int w, const McParams &p, uint32_t replica,
                          uint64_t sweep) {
    const int o = 1 - c;
    int y = s.y_first + lr;
    if (y >= s.L) y -= s.L;
    if (y >= s.L) y %= s.L;
    const uint32_t t = s0_get(s, c, lr, w);
    const uint32_t u = s0_get(s, o, lr - 1, w);
    const uint32_t d = s0_get(s, o, lr + 1, w);
    const uint32_t n0 = s0_get(s, o, lr, w);
    const uint32_t n1 = ((y + c) & 1) ? s0_up(s, o, lr, w) : s0_dn(s, o, lr, w);
    const uint32_t word_id = (uint32_t)(((size_t)c * s.L + y) * s.W + w);
    const uint32_t flip = metropolis_flip_mask(t, u, d, n0, n1, s.mask, p, word_id, replica, sweep);
    s.base[(c * s.rows + lr) * s.W + w] = t ^ flip;
}

// ---- blocked levels: natural layout ----------------------------------------------------------------------------
struct StripN {
    const uint32_t *x;  // x[lr*W + w]
    int W, bits;
    uint32_t mask;
};

MCRG_HD uint32_t sn_get(const StripN &s, int lr, int w) { return s.x[lr * s.W + w]; }
MCRG_HD uint32_t sn_up(const StripN &s, int lr, int w) {
    const int wn = (w + 1 == s.W) ? 0 : w + 1;
    return shift_up_index(sn_get(s, lr, w), sn_get(s, lr, wn), s.bits, s.mask);
}
MCRG_HD uint32_t sn_dn(const StripN &s, int lr, int w) {
    const int wp = (w == 0) ? s.W - 1 : w - 1;
    return shift_down_index(sn_get(s, lr, w), sn_get(s, lr, wp), s.bits, s.mask);
}

// correlator popcounts of local row lr (needs row lr_below = the next row, periodic handled by the caller)
MCRG_HD void measure_rowN(const StripN &s, int lr, int lr_below, int w, Counts &cnt) {
    const uint32_t r0 = sn_get(s, lr, w), r1 = sn_get(s, lr_below, w);
    const uint32_t r0u = sn_up(s, lr, w), r1u = sn_up(s, lr_below, w), r1d = sn_dn(s, lr_below, w);
    cnt.anti_nn += popc32(r0 ^ r0u) + popc32(r0 ^ r1);
    cnt.anti_nnn += popc32(r0 ^ r1u) + popc32(r0 ^ r1d);
    cnt.odd_plaq += popc32(r0 ^ r0u ^ r1 ^ r1u);
    cnt.up += popc32(r0);
}

// majority/tie words of output word wb of block row (lr, lr+1): input words 2wb, 2wb+1 (or word 0 if W == 1)
MCRG_HD void block_pairN(const StripN &s, int lr, int wb, uint32_t &maj, uint32_t &tie) {
    uint32_t m, t;
    const int wi = (s.W == 1) ? 0 : 2 * wb;
    uint32_t r0 = sn_get(s, lr, wi), r1 = sn_get(s, lr + 1, wi);
    majority4(r0, r0 >> 1, r1, r1 >> 1, m, t);
    maj = compress_even(m);
    tie = compress_even(t);
    if (s.W > 1) {
        r0 = sn_get(s, lr, wi + 1);
        r1 = sn_get(s, lr + 1, wi + 1);
        majority4(r0, r0 >> 1, r1, r1 >> 1, m, t);
        maj |= compress_even(m) << 16;
        tie |= compress_even(t) << 16;
    }
}

}  // namespace mcrg
