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

// The same two functions with c3_base ^ k1 formed once per half-sweep (`ck`; the call index occupies bits of c3_base that are zero,
// so OR-ing it in equals XOR-ing it in): the row loop then has no use for the key word k1 any more — it had been re-read from
// the constant bank every row pair.
MCRG_HD void mc_philox_pair_ck(const McPhiloxHead &h, uint64_t seed, uint32_t word_id, uint32_t ck, U4 &r0, U4 &r1) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t ph, pl;
    mulwide(0xD2511F53u, word_id, ph, pl);
    const uint32_t c2a = ph ^ ck;
    const uint32_t c2b = c2a ^ (1u << 20);
    const uint32_t c2n = h.h1 ^ pl ^ (k1 + 0xBB67AE85u);
    uint32_t qh, ql;
    mulwide(0xCD9E8D57u, c2a, qh, ql);
    r0 = mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, c2n, h.l1, k0, k1);
    mulwide(0xCD9E8D57u, c2b, qh, ql);
    r1 = mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, c2n, h.l1, k0, k1);
}

MCRG_HD U4 mc_philox_j_ck(const McPhiloxHead &h, uint64_t seed, uint32_t word_id, uint32_t ck, int j) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t ph, pl, qh, ql;
    mulwide(0xD2511F53u, word_id, ph, pl);
    mulwide(0xCD9E8D57u, ph ^ ck ^ ((uint32_t)j << 20), qh, ql);
    return mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, h.h1 ^ pl ^ (k1 + 0xBB67AE85u), h.l1, k0, k1);
}

// call j of word `word_id` with the shared head (pass 2 and the inline overflow path)
MCRG_HD U4 mc_philox_j(const McPhiloxHead &h, uint64_t seed, uint32_t word_id, uint32_t c3_base, int j) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t ph, pl, qh, ql;
    mulwide(0xD2511F53u, word_id, ph, pl);
    mulwide(0xCD9E8D57u, ph ^ (c3_base | ((uint32_t)j << 20)) ^ k1, qh, ql);
    return mc_philox_rounds2to9(qh ^ h.b0 ^ (k0 + 0x9E3779B9u), ql, h.h1 ^ pl ^ (k1 + 0xBB67AE85u), h.l1, k0, k1);
}

}  // namespace mcrg
