// Word-level primitives of the bit-packed MCRG hot path.
//
// Everything here is a pure function of 32-bit words (one bit per spin, up = 1), usable from device code and
// — for the CPU emulation used by tests/test_tile_emulation.py — from plain C++ (MCRG_HD expands to nothing).
//
// Conventions (specified in scalar form by oracle/mcrg_oracle.c, section "(S) sampler specification"):
//   * internal coordinates y = reference column j, x = reference row i  (definitions.hpp:16 is column-major,
//     so an internal row is contiguous in the reference's array);
//   * level-0 state = two colour planes, colour = (x+y)&1; plane c, row y holds the sites
//     x = 2x' + ((y+c)&1), x' packed 32 per word;  W = max(1, L/64) words per row and colour,
//     `bits` = min(32, L/2) valid bits per word;
//   * blocked levels (n >= 1) = natural layout: row y, bit x, Wn = max(1, Ln/32), bits = min(32, Ln);
//   * every random decision is a Philox4x32-10 output selected by (seed; word, replica, sweep, purpose, j).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MCRG_HD __host__ __device__ __forceinline__
#else
#define MCRG_HD inline
#endif

namespace mcrg {

enum : int { PURPOSE_MC = 1, PURPOSE_TIE = 2, PURPOSE_INIT = 3, PURPOSE_SW_BOND = 4, PURPOSE_SW_FLIP = 5 };

struct U4 {
    uint32_t x, y, z, w;
};

MCRG_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  mul.wide -> one IMAD.WIDE per 32x32->64 product.
MCRG_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__) && !defined(MCRG_PHILOX_NO_PTX)  // explicit mul.wide + unpack: fewer register-pair moves
        uint32_t h0, l0, h1, l1;
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1,%0}, p;\n\t}" : "=r"(h0), "=r"(l0) : "r"(0xD2511F53u), "r"(c0));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1,%0}, p;\n\t}" : "=r"(h1), "=r"(l1) : "r"(0xCD9E8D57u), "r"(c2));
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
#endif
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    U4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// counter = (word, replica, t_lo, purpose<<28 | j<<20 | t_hi[20]); key = (seed_lo, seed_hi)
MCRG_HD U4 philox_keyed(uint64_t seed, uint32_t word, uint32_t replica, uint64_t t, int purpose, int j) {
    const uint32_t c3 = ((uint32_t)purpose << 28) | (((uint32_t)j & 0xFFu) << 20) | (uint32_t)((t >> 32) & 0xFFFFFu);
    return philox4x32_10(word, replica, (uint32_t)t, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ---- horizontal neighbours inside a packed row (periodic) ------------------------------------------------
// index+1: bit k of the result is bit k+1 of the row; `next` is the following word (the word itself if W==1)
MCRG_HD uint32_t shift_up_index(uint32_t cur, uint32_t next, int bits, uint32_t mask) {
    return ((cur >> 1) | (next << (bits - 1))) & mask;
}
// index-1: bit k of the result is bit k-1 of the row; `prev` is the preceding word
MCRG_HD uint32_t shift_down_index(uint32_t cur, uint32_t prev, int bits, uint32_t mask) {
    return ((cur << 1) | (prev >> (bits - 1))) & mask;
}

MCRG_HD uint32_t valid_mask(int bits) { return bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u); }

// ---- checkerboard Metropolis on one word of 32 same-colour sites --------------------------------------------
// t: the sites; u, d, s0, s1: their four neighbours (other colour).  `anti` = 0 for K <= 0 (ferromagnetic,
// ising.cpp:8-9 sign convention), ~0 for K > 0.  A = number of bonds the flip would repair; flip always if
// A >= 2, with probability exp(-4|K|) if A == 1 and exp(-8|K|) if A == 0, decided as U < T4 / U < T8 where
// the 32-bit uniform U of lane l has, as its k-th most significant bit, bit l of the k-th Philox output word
// (call j = k>>2, element k&3).  The comparison is evaluated lazily, MSB first, for all 32 lanes at once,
// and stops as soon as every lane is decided — identical to the full 32-bit comparison in the oracle.
struct McParams {
    uint64_t seed;
    uint32_t T4, T8;  // floor(exp(-4|K|) 2^32), floor(exp(-8|K|) 2^32)
    uint32_t anti;    // 0 or 0xFFFFFFFF
};

MCRG_HD uint32_t metropolis_flip_mask(uint32_t t, uint32_t u, uint32_t d, uint32_t s0, uint32_t s1, uint32_t mask,
                                      const McParams &p, uint32_t word_id, uint32_t replica, uint64_t sweep) {
    const uint32_t a1 = t ^ u ^ p.anti, a2 = t ^ d ^ p.anti, a3 = t ^ s0 ^ p.anti, a4 = t ^ s1 ^ p.anti;
    const uint32_t x12 = a1 ^ a2, c12 = a1 & a2, x34 = a3 ^ a4, c34 = a3 & a4;
    const uint32_t ge2 = c12 | c34 | (x12 & x34);          // A >= 2
    const uint32_t m1 = (x12 ^ x34) & ~(c12 | c34) & mask; // A == 1
    const uint32_t m0 = ~(a1 | a2 | a3 | a4) & mask;        // A == 0
    uint32_t eq4 = m1, eq8 = m0, lt = 0;
    uint32_t T4s = p.T4, T8s = p.T8;
    for (int j = 0; j < 8 && (eq4 | eq8) != 0u; ++j) {
        const U4 r = philox_keyed(p.seed, word_id, replica, sweep, PURPOSE_MC, j);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t tm4 = (uint32_t)(-(int32_t)(T4s >> 31));
            const uint32_t tm8 = (uint32_t)(-(int32_t)(T8s >> 31));
            lt |= (eq4 & ~rr[e] & tm4) | (eq8 & ~rr[e] & tm8);
            eq4 &= ~(rr[e] ^ tm4);
            eq8 &= ~(rr[e] ^ tm8);
            T4s <<= 1;
            T8s <<= 1;
        }
    }
    return (ge2 | lt) & mask;
}

// ---- b = 2 majority rule (mcrg.cpp:314-348) on four bit-planes a,b,c,d of the same blocks ----------------------
MCRG_HD void majority4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t &maj, uint32_t &tie) {
    const uint32_t x = a ^ b, cab = a & b, y = c ^ d, ccd = c & d;
    maj = (cab & (y | ccd)) | (ccd & x);          // three or four up
    tie = (x & y) | ((cab ^ ccd) & ~(x | y));      // exactly two up: block sum == 0
}

// gather the even-position bits of v into the low 16 bits
MCRG_HD uint32_t compress_even(uint32_t v) {
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0F0F0F0Fu;
    v = (v | (v >> 4)) & 0x00FF00FFu;
    v = (v | (v >> 8)) & 0x0000FFFFu;
    return v;
}

// Tie coins.  The 32 coins of output word q (= yb*Wb + wb) of the level-`level` lattice are ONE of the four output
// words of a Philox call: call index = q with bits 8-9 removed, element = bits 8-9 of q.  A thread that walks its words
// with stride 256 (every kernel here does, with 256 threads) therefore needs one call per four words; TieCache keeps
// the last call.  The mapping is a fixed function of q, so results do not depend on who computes it.
MCRG_HD uint32_t tie_group(uint32_t q) { return ((q >> 10) << 8) | (q & 255u); }
MCRG_HD uint32_t tie_pick(const U4 &r, uint32_t q) {
    const uint32_t e = (q >> 8) & 3u;
    return e == 0u ? r.x : (e == 1u ? r.y : (e == 2u ? r.z : r.w));
}
MCRG_HD uint32_t tie_word(uint64_t seed, uint32_t q, uint32_t replica, uint64_t t, int level) {
    return tie_pick(philox_keyed(seed, tie_group(q), replica, t, PURPOSE_TIE, level), q);
}
struct TieCache {
    uint32_t group;
    U4 r;
    MCRG_HD void init() { group = 0xFFFFFFFFu; r.x = r.y = r.z = r.w = 0u; }
    MCRG_HD uint32_t get(uint64_t seed, uint32_t q, uint32_t replica, uint64_t t, int level) {
        const uint32_t g = tie_group(q);
        if (g != group) {
            r = philox_keyed(seed, g, replica, t, PURPOSE_TIE, level);
            group = g;
        }
        return tie_pick(r, q);
    }
};

// geometry helpers
MCRG_HD int l0_words(int L) { return L >= 64 ? L / 64 : 1; }
MCRG_HD int l0_bits(int L) { return L >= 64 ? 32 : L / 2; }
MCRG_HD int nat_words(int Ln) { return Ln >= 32 ? Ln / 32 : 1; }
MCRG_HD int nat_bits(int Ln) { return Ln >= 32 ? 32 : Ln; }

}  // namespace mcrg
