// CPU emulation of the CUDA kernels' strip logic.  TEST INFRASTRUCTURE ONLY (never part of the product).
//
// mcrg_b200/csrc/{bitops,tile}.cuh hold the per-word device functions as host/device code.  This file walks
// them with the same strip / halo / phase structure as kernels.cu (k_sweep0, k_level, k_tail), "shared memory"
// being a std::vector, so that the bit tricks, the halo recomputation and the Philox keying can be checked
// against the oracle on a machine without a GPU.  What it cannot cover — staging, reductions, atomics, launch
// geometry — is covered by the -m gpu tests on the real kernels.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mcrg_b200/csrc/mcfast.cuh"
#include "../../mcrg_b200/csrc/tile.cuh"

using namespace mcrg;

namespace {

constexpr int MAX_LEVELS = 15;
constexpr int TAIL_MAX_L = 256;

struct Emu {
    int L, W, bits;
    uint64_t seed;
    uint32_t replica;
    std::vector<uint32_t> planes[2];
    int cur = 0;
    std::vector<std::vector<uint32_t>> levels;  // [lv] natural layout
    unsigned long long cnt[MAX_LEVELS + 1][4];
    McParams mc;
    uint64_t t = 0;
};

void sweep0(Emu &e, int R, int nsw, bool measure) {
    const int L = e.L, W = e.W, H = nsw > 0 ? 2 * nsw : 2, rows = R + 2 * H;
    const std::vector<uint32_t> &src = e.planes[e.cur];
    std::vector<uint32_t> &dst = e.planes[1 - e.cur];
    std::vector<uint32_t> smem((size_t)2 * rows * W);
    for (int strip = 0; strip < L / R; ++strip) {
        const int y0 = strip * R;
        Strip0 s;
        s.base = smem.data();
        s.rows = rows;
        s.W = W;
        s.bits = e.bits;
        s.mask = valid_mask(e.bits);
        s.L = L;
        s.y_first = (y0 - H) & (L - 1);
        for (int c = 0; c < 2; ++c)
            for (int lr = 0; lr < rows; ++lr) {
                const int y = (s.y_first + lr) & (L - 1);
                for (int w = 0; w < W; ++w) smem[((size_t)c * rows + lr) * W + w] = src[((size_t)c * L + y) * W + w];
            }
        if (measure) {
            Counts c = {0, 0, 0, 0};
            for (int i = 0; i < R / 2; ++i)
                for (int w = 0; w < W; ++w) {
                    uint32_t maj, tie;
                    measure_pair0(s, H + 2 * i, w, c, maj, tie);
                    const uint32_t q = (uint32_t)((y0 / 2 + i) * W + w);
                    uint32_t out = maj;
                    if (tie) out |= tie & tie_word(e.seed, q, e.replica, e.t, 1);
                    e.levels[1][q] = out;
                }
            e.cnt[0][0] += c.anti_nn;
            e.cnt[0][1] += c.anti_nnn;
            e.cnt[0][2] += c.odd_plaq;
            e.cnt[0][3] += c.up;
        }
        if (nsw > 0) {
            for (int h = 0; h < 2 * nsw; ++h) {
                const int c = h & 1;
                for (int lr = 1 + h; lr < rows - 1 - h; ++lr)
                    for (int w = 0; w < W; ++w) update_word0(s, c, lr, w, e.mc, e.replica, e.t + (uint64_t)(h >> 1));
            }
            for (int c = 0; c < 2; ++c)
                for (int lr = 0; lr < R; ++lr)
                    for (int w = 0; w < W; ++w)
                        dst[((size_t)c * L + y0 + lr) * W + w] = smem[((size_t)c * rows + H + lr) * W + w];
        }
    }
    if (nsw > 0) {
        e.cur ^= 1;
    }
}

void level_strips(Emu &e, int lv, int R, bool do_block) {
    const int Ln = e.L >> lv, Wn = nat_words(Ln);
    const std::vector<uint32_t> &in = e.levels[lv];
    std::vector<uint32_t> smem((size_t)(R + 1) * Wn);
    for (int strip = 0; strip < Ln / R; ++strip) {
        const int y0 = strip * R;
        for (int lr = 0; lr <= R; ++lr)
            for (int w = 0; w < Wn; ++w) smem[(size_t)lr * Wn + w] = in[(size_t)((y0 + lr) & (Ln - 1)) * Wn + w];
        StripN s;
        s.x = smem.data();
        s.W = Wn;
        s.bits = nat_bits(Ln);
        s.mask = valid_mask(s.bits);
        Counts c = {0, 0, 0, 0};
        for (int lr = 0; lr < R; ++lr)
            for (int w = 0; w < Wn; ++w) measure_rowN(s, lr, lr + 1, w, c);
        if (do_block) {
            const int Lb = Ln / 2, Wb = nat_words(Lb);
            for (int i = 0; i < R / 2; ++i)
                for (int wb = 0; wb < Wb; ++wb) {
                    uint32_t maj, tie;
                    block_pairN(s, 2 * i, wb, maj, tie);
                    const uint32_t q = (uint32_t)((y0 / 2 + i) * Wb + wb);
                    uint32_t o = maj;
                    if (tie) o |= tie & tie_word(e.seed, q, e.replica, e.t, lv + 1);
                    e.levels[lv + 1][q] = o;
                }
        }
        e.cnt[lv][0] += c.anti_nn;
        e.cnt[lv][1] += c.anti_nnn;
        e.cnt[lv][2] += c.odd_plaq;
        e.cnt[lv][3] += c.up;
    }
}

void tail(Emu &e, int start, int n_levels) {
    for (int lv = start; lv <= n_levels; ++lv) {
        const int Ln = e.L >> lv, Wn = nat_words(Ln);
        StripN s;
        s.x = e.levels[lv].data();
        s.W = Wn;
        s.bits = nat_bits(Ln);
        s.mask = valid_mask(s.bits);
        Counts c = {0, 0, 0, 0};
        for (int lr = 0; lr < Ln; ++lr)
            for (int w = 0; w < Wn; ++w) measure_rowN(s, lr, lr + 1 == Ln ? 0 : lr + 1, w, c);
        e.cnt[lv][0] += c.anti_nn;
        e.cnt[lv][1] += c.anti_nnn;
        e.cnt[lv][2] += c.odd_plaq;
        e.cnt[lv][3] += c.up;
        if (lv < n_levels) {
            const int Lb = Ln / 2, Wb = nat_words(Lb);
            for (int yb = 0; yb < Lb; ++yb)
                for (int wb = 0; wb < Wb; ++wb) {
                    uint32_t maj, tie;
                    block_pairN(s, 2 * yb, wb, maj, tie);
                    const uint32_t q = (uint32_t)(yb * Wb + wb);
                    uint32_t o = maj;
                    if (tie) o |= tie & tie_word(e.seed, q, e.replica, e.t, lv + 1);
                    e.levels[lv + 1][q] = o;
                }
        }
    }
}

void pack0(Emu &e, const int32_t *spins) {
    const int L = e.L, W = e.W;
    std::vector<uint32_t> &pl = e.planes[e.cur];
    std::fill(pl.begin(), pl.end(), 0u);
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x)
            if (spins[(size_t)y * L + x] > 0) {
                const int c = (x + y) & 1, xh = x >> 1;
                pl[((size_t)c * L + y) * W + (xh >> 5)] |= 1u << (xh & 31);
            }
}

void unpack0(const Emu &e, int32_t *spins) {
    const int L = e.L, W = e.W;
    const std::vector<uint32_t> &pl = e.planes[e.cur];
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x) {
            const int c = (x + y) & 1, xh = x >> 1;
            spins[(size_t)y * L + x] = ((pl[((size_t)c * L + y) * W + (xh >> 5)] >> (xh & 31)) & 1u) ? 1 : -1;
        }
}

Emu make(int L, uint64_t seed, uint32_t replica, uint32_t T4, uint32_t T8, uint32_t anti, uint64_t t) {
    Emu e;
    e.L = L;
    e.W = l0_words(L);
    e.bits = l0_bits(L);
    e.seed = seed;
    e.replica = replica;
    e.planes[0].assign((size_t)2 * L * e.W, 0u);
    e.planes[1].assign((size_t)2 * L * e.W, 0u);
    e.levels.resize(MAX_LEVELS + 1);
    for (int lv = 1; (L >> lv) >= 2 || lv == 1; ++lv) e.levels[lv].assign((size_t)(L >> lv) * nat_words(L >> lv), 0u);  // level 1 always: sweep0 writes it
    std::memset(e.cnt, 0, sizeof e.cnt);
    e.mc.seed = seed;
    e.mc.T4 = T4;
    e.mc.T8 = T8;
    e.mc.anti = anti;
    e.t = t;
    return e;
}

}  // namespace

extern "C" {

// n_sweeps Metropolis sweeps with strips of R rows and `fuse` sweeps per pass, exactly as mcrg_sweep() sequences
// k_sweep0<false>; spins in/out in the reference layout.
void emul_sweep(int L, int32_t *spins, int R, int fuse, int n_sweeps, uint64_t seed, uint32_t replica, uint32_t T4,
                uint32_t T8, uint32_t anti, uint64_t t0) {
    Emu e = make(L, seed, replica, T4, T8, anti, t0);
    pack0(e, spins);
    int done = 0;
    while (done < n_sweeps) {
        const int k = (n_sweeps - done) < fuse ? (n_sweeps - done) : fuse;
        e.t = t0 + done;
        sweep0(e, R, k, false);
        done += k;
    }
    unpack0(e, spins);
}

// One measurement as enqueue_sample() sequences it (k_sweep0<true> with nsw sweeps fused, k_level while the
// blocked lattice is larger than TAIL_MAX_L, k_tail): S[(lv)*4+k], packed levels unpacked to level_spins
// (concatenated, reference layout), spins updated in place if nsw > 0.  Returns the number of levels.
int emul_measure(int L, int32_t *spins, int R, int Rn, int nsw, int max_levels, uint64_t seed, uint32_t replica,
                 uint32_t T4, uint32_t T8, uint32_t anti, uint64_t t, int64_t *S, int32_t *level_spins) {
    Emu e = make(L, seed, replica, T4, T8, anti, t);
    pack0(e, spins);
    int n_lv = 0;
    while ((L >> (n_lv + 1)) >= 2) ++n_lv;  // log2(L) - 1
    if (max_levels >= 0 && max_levels < n_lv) n_lv = max_levels;
    sweep0(e, R, nsw, true);
    int lv = 1;
    while (lv <= n_lv && (L >> lv) > TAIL_MAX_L) {
        const int Ln = L >> lv;
        level_strips(e, lv, Rn < Ln ? Rn : Ln, lv < n_lv);
        ++lv;
    }
    tail(e, lv, n_lv);
    for (int k = 0; k <= n_lv; ++k) {
        long long s4[4];
        counts_to_S((long long)(L >> k), e.cnt[k][0], e.cnt[k][1], e.cnt[k][2], e.cnt[k][3], s4);
        for (int j = 0; j < 4; ++j) S[k * 4 + j] = s4[j];
    }
    if (level_spins) {
        size_t off = 0;
        for (int k = 1; k <= n_lv; ++k) {
            const int Ln = L >> k, Wn = nat_words(Ln);
            for (int y = 0; y < Ln; ++y)
                for (int x = 0; x < Ln; ++x)
                    level_spins[off + (size_t)y * Ln + x] = ((e.levels[k][(size_t)y * Wn + (x >> 5)] >> (x & 31)) & 1u) ? 1 : -1;
            off += (size_t)Ln * Ln;
        }
    }
    if (nsw > 0) unpack0(e, spins);
    return n_lv;
}

void emul_hot_start(int L, uint64_t seed, uint32_t replica, int32_t *spins) {
    Emu e = make(L, seed, replica, 0, 0, 0, 0);
    const int W = e.W;
    for (size_t idx = 0; idx < (size_t)2 * L * W; ++idx)
        e.planes[0][idx] = philox_keyed(seed, (uint32_t)idx, replica, 0ull, PURPOSE_INIT, 0).x & valid_mask(e.bits);
    unpack0(e, spins);
}

// ---- fast forms of the per-word arithmetic (mcfast.cuh) against the specification (bitops.cuh) -----------------------
// The sweep kernels do not call philox4x32_10 / metropolis_flip_mask: they share the word-independent part of Philox
// rounds 0-1 between calls, count broken bonds with a full adder and compare the first call's planes with code
// specialised on the leading threshold bits.  Random words, keys and thresholds of every pattern: both ways must agree
// on every bit.  Returns the number of disagreements.
static uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void cmp4_generic(const U4 &r, uint32_t T4, uint32_t T8, int plane0, uint32_t sel, uint32_t &eq, uint32_t &lt) {
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    for (int e = 0; e < 4; ++e) {
        const int k = plane0 + e;
        const uint32_t t4 = ((T4 >> (31 - k)) & 1u) ? 0xFFFFFFFFu : 0u, t8 = ((T8 >> (31 - k)) & 1u) ? 0xFFFFFFFFu : 0u;
        const uint32_t tm = (sel & t4) | (~sel & t8);
        lt |= eq & ~rr[e] & tm;
        eq &= ~(rr[e] ^ tm);
    }
}

static bool same(const U4 &a, const U4 &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

int emul_fast_paths(int n_trials, uint64_t seed0) {
    uint64_t x = seed0;
    int bad = 0;
    for (int it = 0; it < n_trials; ++it) {
        const uint64_t seed = splitmix(x), t = (it & 7) ? splitmix(x) : (splitmix(x) | 0xFFFFFFFF00000000ull);
        const uint32_t replica = (uint32_t)splitmix(x), word = (uint32_t)splitmix(x) & 0x3FFFFFFu;
        // (1) Philox with the shared head == Philox4x32-10, every call index
        const uint32_t c3_base = ((uint32_t)PURPOSE_MC << 28) | (uint32_t)((t >> 32) & 0xFFFFFu);
        const McPhiloxHead h = mc_philox_head(seed, replica, (uint32_t)t);
        U4 r0, r1;
        mc_philox_pair(h, seed, word, c3_base, r0, r1);
        if (!same(r0, philox_keyed(seed, word, replica, t, PURPOSE_MC, 0))) ++bad;
        if (!same(r1, philox_keyed(seed, word, replica, t, PURPOSE_MC, 1))) ++bad;
        for (int j = 0; j < 8; ++j)
            if (!same(mc_philox_j(h, seed, word, c3_base, j), philox_keyed(seed, word, replica, t, PURPOSE_MC, j))) ++bad;
        // the forms the strip / resident kernels call: c3_base ^ key word 1 formed once (mc_philox_pair_ck, mc_philox_j_ck)
        const uint32_t ck = c3_base ^ (uint32_t)(seed >> 32);
        U4 s0, s1;
        mc_philox_pair_ck(h, seed, word, ck, s0, s1);
        if (!same(s0, r0) || !same(s1, r1)) ++bad;
        for (int j = 0; j < 8; ++j)
            if (!same(mc_philox_j_ck(h, seed, word, ck, j), philox_keyed(seed, word, replica, t, PURPOSE_MC, j))) ++bad;
        // (2) one word update: thresholds of every leading-bit pattern (and general ones), spins near and far from order
        McParams p;
        p.seed = seed;
        p.anti = (splitmix(x) & 1) ? 0xFFFFFFFFu : 0u;
        const int pat = it % 6;  // 0..3: T4 < 1/4 with planes (2,3) = pat; 4: general; 5: tiny thresholds
        if (pat < 4) {
            p.T4 = ((uint32_t)pat << 28) | ((uint32_t)splitmix(x) & 0x0FFFFFFFu);
            p.T8 = (uint32_t)splitmix(x) & 0x0FFFFFFFu;
        } else if (pat == 4) {
            p.T4 = (uint32_t)splitmix(x) | 0x40000000u;
            p.T8 = (uint32_t)splitmix(x);
        } else {
            p.T4 = (uint32_t)splitmix(x) & 0xFFFFu;
            p.T8 = (uint32_t)splitmix(x) & 0xFFu;
        }
        const uint32_t tw = (uint32_t)splitmix(x);
        uint32_t nb[4];
        for (int k = 0; k < 4; ++k) {  // neighbours = the word itself with a sparse, dense or random set of differences
            const uint32_t m1 = (uint32_t)splitmix(x), m2 = (uint32_t)splitmix(x), m3 = (uint32_t)splitmix(x);
            const int kind = (int)(splitmix(x) % 3);
            nb[k] = (tw ^ p.anti) ^ (kind == 0 ? (m1 & m2 & m3) : kind == 1 ? (m1 | m2) : m1);
        }
        const uint32_t want = metropolis_flip_mask(tw, nb[0], nb[1], nb[2], nb[3], 0xFFFFFFFFu, p, word, replica, t);
        const uint32_t a1 = tw ^ nb[0] ^ p.anti, a2 = tw ^ nb[1] ^ p.anti, a3 = tw ^ nb[2] ^ p.anti, a4 = tw ^ nb[3] ^ p.anti;
        uint32_t ge2, sel;
        mc_neighbour_count(a1, a2, a3, a4, ge2, sel);
        uint32_t eq = ~ge2, lt = 0u;
        const bool nz = (p.T4 >> 30) == 0u && (p.T8 >> 28) == 0u;
        if (nz) {
            switch ((p.T4 >> 28) & 3u) {
                case 0: mc_compare4_nz<0>(r0, sel, eq, lt); break;
                case 1: mc_compare4_nz<1>(r0, sel, eq, lt); break;
                case 2: mc_compare4_nz<2>(r0, sel, eq, lt); break;
                default: mc_compare4_nz<3>(r0, sel, eq, lt); break;
            }
        } else {
            cmp4_generic(r0, p.T4, p.T8, 0, sel, eq, lt);
        }
        cmp4_generic(r1, p.T4, p.T8, 4, sel, eq, lt);
        for (int j = 2; j < 8 && eq != 0u; ++j) cmp4_generic(mc_philox_j(h, seed, word, c3_base, j), p.T4, p.T8, 4 * j, sel, eq, lt);
        if ((ge2 | lt) != want) ++bad;
    }
    return bad;
}

// ---- pass 2 of a half-sweep as k_sweep0 schedules it (mc_half_sweep_t<.., REQUEUE = true>) -----------------------------------
// A warp's queue of words with undecided lanes is consumed one entry per lane, 32 at a time.  Every batch but the last runs ONE
// further Philox call per entry and appends the entries that still have undecided lanes to the tail (with the next call index);
// the last batch finishes its entries completely; entries that do not fit the segment any more are finished on the spot.  Here:
// the same schedule, lane by lane, against "finish every entry on its own" — the flip mask of every entry must be the same
// whatever the queue length, the capacity and the luck of the draws.  Returns the number of entries that differ.
int emul_requeue_pass2(int n_trials, uint64_t seed0) {
    uint64_t x = seed0;
    int bad = 0;
    for (int it = 0; it < n_trials; ++it) {
        const uint64_t seed = splitmix(x), t = splitmix(x);
        const uint32_t replica = (uint32_t)splitmix(x);
        const uint32_t c3_base = ((uint32_t)PURPOSE_MC << 28) | (uint32_t)((t >> 32) & 0xFFFFFu);
        const McPhiloxHead h = mc_philox_head(seed, replica, (uint32_t)t);
        // thresholds that leave many lanes undecided for several calls: long runs of equal leading bits are what re-queues
        const uint32_t T4 = (it & 1) ? (uint32_t)splitmix(x) : ((uint32_t)splitmix(x) & 0x000FFFFFu), T8 = (it & 2) ? (uint32_t)splitmix(x) : 0u;
        const int n0 = (int)(splitmix(x) % 200), cap = 8 + (int)(splitmix(x) % 220);
        struct Ent { uint32_t word, eq, sel, j; };
        std::vector<Ent> q;
        std::vector<uint32_t> want, got;
        for (int e = 0; e < n0; ++e) {
            Ent en;
            en.word = (uint32_t)splitmix(x) & 0x3FFFFFFu;
            en.sel = (uint32_t)splitmix(x);
            en.eq = (it & 4) ? (uint32_t)splitmix(x) : ((uint32_t)splitmix(x) & (uint32_t)splitmix(x) & (uint32_t)splitmix(x));
            if (en.eq == 0u) en.eq = 1u << (e & 31);
            en.j = 2;
            q.push_back(en);
        }
        // reference: every entry finished on its own (mc_finish); entries beyond the capacity never reach the queue (mc_push
        // finishes them inline in pass 1), so the schedule below only sees min(n0, cap)
        const int total0 = n0 < cap ? n0 : cap;
        auto finish = [&](uint32_t word, uint32_t eq, uint32_t sel, int j0) {
            uint32_t lt = 0u;
            for (int j = j0; j < 8 && eq != 0u; ++j) cmp4_generic(mc_philox_j(h, seed, word, c3_base, j), T4, T8, 4 * j, sel, eq, lt);
            return lt;
        };
        // an entry keeps its word through the re-queues: make the words unique and key the result by them
        for (int e = 0; e < total0; ++e) q[e].word = (q[e].word & 0x03FFFF00u) | (uint32_t)(e & 0xFF) | ((uint32_t)(e >> 8) << 26);
        for (int e = 0; e < total0; ++e) want.push_back(finish(q[e].word, q[e].eq, q[e].sel, 2));
        got.assign(total0, 0u);
        auto index_of = [&](uint32_t word) { return (int)((word & 0xFFu) | ((word >> 26) << 8)); };
        q.resize(cap > total0 ? cap : total0);
        int total = total0;
        for (int base = 0; base < total; base += 32) {
            const bool last = total <= base + 32;
            std::vector<Ent> again;
            for (int lane = 0; lane < 32; ++lane) {
                const int e = base + lane;
                if (e >= total) continue;
                Ent en = q[e];
                if (last) {
                    got[index_of(en.word)] ^= finish(en.word, en.eq, en.sel, (int)en.j);
                } else {
                    uint32_t lt = 0u;
                    cmp4_generic(mc_philox_j(h, seed, en.word, c3_base, (int)en.j), T4, T8, 4 * (int)en.j, en.sel, en.eq, lt);
                    got[index_of(en.word)] ^= lt;
                    if (en.j >= 7) en.eq = 0u;
                    if (en.eq != 0u) {
                        en.j += 1;
                        again.push_back(en);
                    }
                }
            }
            for (size_t r = 0; r < again.size(); ++r) {  // rank among the lanes that go again
                const int slot = total + (int)r;
                if (slot < cap) q[slot] = again[r];
                else got[index_of(again[r].word)] ^= finish(again[r].word, again[r].eq, again[r].sel, (int)again[r].j);
            }
            total = total + (int)again.size() < cap ? total + (int)again.size() : cap;  // total0 <= cap
        }
        for (int e = 0; e < total0; ++e)
            if (got[e] != want[e]) ++bad;
    }
    return bad;
}

}  // extern "C"
