// Fast forms of the per-word Metropolis arithmetic used by the sweep kernels (kernels.cu: mc_row, mc_finish).
//
// Each function here computes exactly what the scalar specification in bitops.cuh (philox4x32_10, metropolis_flip_mask)
// computes, with less work.  They are host/device code so that tests/cpu_emul/emul.cpp can check them against the
// specification on a machine without a GPU (tests/test_tile_emulation.py::test_fast_paths_equal_the_specification);
// on the device the sweeps are additionally compared bit for bit with the oracle's scalar sampler.
#pragma once
#include "bitops.cuh"

namespace mcrg {

// 32 x 32 -> 64 bit product as (hi, lo); on the device one IMAD.WIDE
MCRG_HD void mulwide(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1,%0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

// A = a1 + a2 + a3 + a4 per lane (a_i = 1: the bond to neighbour i is broken and a flip would repair it): full adder of
// three, then the fourth.  ge2 = lanes with A >= 2, sel = lanes with A == 1; the lanes with A <= 1 are ~ge2.
MCRG_HD void mc_neighbour_count(uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t &ge2, uint32_t &sel) {
    const uint32_t s3 = a1 ^ a2 ^ a3, c3 = (a1 & a2) | (a3 & (a1 | a2));
    ge2 = c3 | (s3 & a4);
    sel = (s3 ^ a4) & ~c3;
}

// Planes 0-3 when T4 < 1/4 (|K| > 0.3466, which includes the whole critical region) and therefore T8 <= T4^2 < 1/16:
// the threshold bit of planes 0 and 1 is 0 for every lane and that of planes 2 and 3 is 0 for the A == 0 lanes, so
//   planes 0, 1: a lane survives only if its uniform has both bits clear, nobody is accepted — one LOP3 for both;
//   plane 2 / 3: T4's bit (template parameter XY = 2 * bit2 + bit3, the same for every lane) decides between
//                "threshold bit = sel" (two LOP3) and "threshold bit = 0" (one).
// Same decisions as mc_compare4 for such thresholds; mc_half_sweep_t checks the table before choosing this path.
template <int XY>
MCRG_HD void mc_compare4_nz(const U4 &r, uint32_t sel, uint32_t &eq, uint32_t &lt) {
    eq &= ~(r.x | r.y);
    if (XY & 2) {
        lt |= eq & ~r.z & sel;
        eq &= ~(r.z ^ sel);
    } else {
        eq &= ~r.z;
    }
    if (XY & 1) {
        lt |= eq & ~r.w & sel;
        eq &= ~(r.w ^ sel);
    } else {
        eq &= ~r.w;
    }
}

// Pass 1 draws calls j = 0 and 1 of the same word.  Of the counter (word, replica, t_lo, c3_j) only `word` changes from
// row to row and only c3 differs between the two calls, so part of rounds 0 and 1 is constant over a half-sweep:
//   round 0:  M1 * t_lo  (hence the new c0 = hi ^ replica ^ k0 and the new c1 = lo)
//   round 1:  M0 * c0    (its hi/lo halves)
// McPhiloxHead holds these three words; mc_philox_pair then spends 2 + 2 x 17 instead of 2 x 20 multiplications per row.
// Same function as philox4x32_10 (bitops.cuh), bit for bit.
struct McPhiloxHead {
    uint32_t b0;      // c1 after round 0
    uint32_t h1, l1;  // hi / lo of M0 * (c0 after round 0)
};

MCRG_HD McPhiloxHead mc_philox_head(uint64_t seed, uint32_t replica, uint32_t t_lo) {
    uint32_t hi, lo;
    mulwide(0xCD9E8D57u, t_lo, hi, lo);
    McPhiloxHead h;
    h.b0 = lo;
    mulwide(0xD2511F53u, hi ^ replica ^ (uint32_t)seed, h.h1, h.l1);
    return h;
}

MCRG_HD U4 mc_philox_rounds2to9(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    k0 += 2u * 0x9E3779B9u;
    k1 += 2u * 0xBB67AE85u;
#pragma unroll
    for (int r = 2; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulwide(0xD2511F53u, c0, h0, l0);
        mulwide(0xCD9E8D57u, c2, h1, l1);
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    U4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// calls j = 0 and j = 1 of word `word_id` (c3 = c3_base | j << 20)
MCRG_HD void mc_philox_pair(const McPhiloxHead &h, uint64_t seed, uint32_t word_id, uint32_t c3_base, U4 &r0,
                                               U4 &r1) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t ph, pl;
    mulwide(0xD2511F53u, word_id, ph, pl);            // round 0, the product that depends on the word
    const uint32_t c2a = ph ^ c3_base ^ k1;           // c2 after round 0, call 0
    const uint32_t c2b = c2a ^ (1u << 20);            //                   call 1 (c3 differs in bit 20 only)
    const uint32_t c2n = h.h1 ^ pl ^ (k1 + 0xBB67AE85u);  // c2 after round 1 (both calls); c3 after round 1 = h.l1
    uint32_t qh, ql;
    mulwide(0xCD9E8D57u, c2a, qh, ql);                // round 1, call 0
    r0 = mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, c2n, h.l1, k0, k1);
    mulwide(0xCD9E8D57u, c2b, qh, ql);                // round 1, call 1
    r1 = mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, c2n, h.l1, k0, k1);
}

// call j of word `word_id` with the shared head (pass 2 and the inline overflow path)
MCRG_HD U4 mc_philox_j(const McPhiloxHead &h, uint64_t seed, uint32_t word_id, uint32_t c3_base, int j) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t ph, pl, qh, ql;
    mulwide(0xD2511F53u, word_id, ph, pl);
    mulwide(0xCD9E8D57u, ph ^ (c3_base | ((uint32_t)j << 20)) ^ k1, qh, ql);
    return mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, h.h1 ^ pl ^ (k1 + 0xBB67AE85u), h.l1, k0, k1);
}

// The per-plane threshold masks (bit k of T4 / T8 replicated over a word); the kernels keep the table in shared memory:
// broadcast loads on the otherwise idle LSU pipe instead of shifts on the ALU pipe, which is the binding pipe.
struct McTable {
    uint32_t tm[32][2];  // [plane][0: T4 bit, 1: T8 bit] as 0 / 0xFFFFFFFF
};

MCRG_HD void mc_table_fill(McTable &tab, uint32_t T4, uint32_t T8) {
    for (int k = 0; k < 32; ++k) {
        tab.tm[k][0] = ((T4 >> (31 - k)) & 1u) ? 0xFFFFFFFFu : 0u;
        tab.tm[k][1] = ((T8 >> (31 - k)) & 1u) ? 0xFFFFFFFFu : 0u;
    }
}

// planes [plane0, plane0 + 4) for thresholds of any shape
MCRG_HD void mc_compare4(const U4 &r, const McTable *tab, int plane0, uint32_t sel, uint32_t &eq, uint32_t &lt) {
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t t4 = tab->tm[plane0 + e][0], t8 = tab->tm[plane0 + e][1];
        const uint32_t tm = (sel & t4) | (~sel & t8);  // this lane's threshold bit (A==1 lanes: T4, A==0: T8)
        lt |= eq & ~rr[e] & tm;                        // U bit 0 where T bit 1, prefix equal: U < T
        eq &= ~(rr[e] ^ tm);
    }
}

// ---- one row pair, one word column: what the sweep kernels execute (specification: metropolis_flip_pair, bitops.cuh) ----
// Pass 1 = three Philox calls for the two words — call 0 of each and call 1 of the A row's word, which the B row shares on the
// lanes the A row no longer needs — and eight lazily compared planes per word, straight-line.  What is left goes to pass 2.
struct McPairOut {
    uint32_t flip_a, eq_a, sel_a;         // A row: flips decided so far; lanes still undecided after planes 0-7 (next: its call 2)
    uint32_t flip_b, eq_b, sel_b, own_b;  // B row likewise; own_b = lanes that could not share and start at their OWN call 1
};

// NZ < 0: any thresholds;  NZ = 0..3: thresholds with T4 < 1/4 whose planes 2 and 3 are NZ (mc_compare4_nz)
// a1..a4 = t ^ neighbour ^ anti per row; `mask` = valid lanes of a word
template <int NZ>
MCRG_HD void mc_pair_pass1(const uint32_t aa[4], const uint32_t ab[4], uint32_t mask, const McPhiloxHead &h, uint64_t seed,
                           uint32_t word_a, uint32_t word_b, uint32_t c3_base, const McTable *tab, McPairOut &o) {
    uint32_t ge2a, ge2b;
    mc_neighbour_count(aa[0], aa[1], aa[2], aa[3], ge2a, o.sel_a);
    mc_neighbour_count(ab[0], ab[1], ab[2], ab[3], ge2b, o.sel_b);
    uint32_t eqa = ~ge2a & mask, eqb = ~ge2b & mask;  // A == 1 or A == 0: lanes that need a random number
    uint32_t lta = 0u, ltb = 0u;                      // subsets of the initial eq, hence disjoint from the A >= 2 lanes
    U4 r0a, r1, r0b;
    mc_philox_pair(h, seed, word_a, c3_base, r0a, r1);
    r0b = mc_philox_j(h, seed, word_b, c3_base, 0);
    if (NZ >= 0) {
        mc_compare4_nz<NZ>(r0a, o.sel_a, eqa, lta);
        mc_compare4_nz<NZ>(r0b, o.sel_b, eqb, ltb);
    } else {
        mc_compare4(r0a, tab, 0, o.sel_a, eqa, lta);
        mc_compare4(r0b, tab, 0, o.sel_b, eqb, ltb);
    }
    const uint32_t und_a = eqa;  // the A row's lanes that use the shared call themselves
    mc_compare4(r1, tab, 4, o.sel_a, eqa, lta);
    o.own_b = eqb & und_a;
    eqb &= ~und_a;
    mc_compare4(r1, tab, 4, o.sel_b, eqb, ltb);
    o.flip_a = (ge2a & mask) | lta;
    o.eq_a = eqa;
    o.flip_b = (ge2b & mask) | ltb;
    o.eq_b = eqb;
}

// pass 2, one word: lanes `own` start at the word's own call 1 (planes 4-7), lanes `eq` at call 2; returns the flips
MCRG_HD uint32_t mc_finish(uint32_t eq, uint32_t own, uint32_t sel, const McTable *tab, const McPhiloxHead &h, uint64_t seed,
                           uint32_t word_id, uint32_t c3_base) {
    uint32_t lt = 0u;
    if (own != 0u) mc_compare4(mc_philox_j(h, seed, word_id, c3_base, 1), tab, 4, sel, own, lt);
    eq |= own;
    for (int j = 2; j < 8 && eq != 0u; ++j) mc_compare4(mc_philox_j(h, seed, word_id, c3_base, j), tab, 4 * j, sel, eq, lt);
    return lt;
}

}  // namespace mcrg
