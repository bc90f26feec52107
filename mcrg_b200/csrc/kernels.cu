// sm_100a kernels of the MCRG hot path (no tensor cores: nothing here is a contraction; the work is bitwise
// integer + Philox, staged through shared memory, reduced with warp primitives and one atomic pass per CTA).
//
//   k_sweep0<MEASURE>  one strip of R rows of one replica: stage strip + halo in shared memory (TMA bulk copies +
//                      mbarrier; plain loads for rows shorter than 16 bytes),
//                      [MEASURE: level-0 correlator popcounts + b=2 majority block to level 1 (Philox ties)],
//                      nsw full checkerboard Metropolis sweeps with halo recomputation (counter-based RNG makes
//                      the redundant halo updates bit-identical to the owning strip's), store the strip (TMA).
//                      Replaces IsingModel::sample_new_configuration (ising.cpp:87-155, Wolff there, Metropolis
//                      here per the north_star), Lattice::calc_interactions (lattice.cpp:102-120) and the first
//                      block_spin_transformation (mcrg.cpp:314-348) of the sample loop mcrg.cpp:72-98.
//   k_level            natural-layout level n: correlator popcounts + block to level n+1, strips of rows.
//   k_tail             one CTA per replica: remaining (small) levels entirely in shared memory, then the
//                      accumulation of mcrg.cpp:86-97 (S, S(n) x S(n-1), S(n) x S(n)) into exact 128-bit sums.
//   k_resident         lattices up to 512^2: the whole replica, its pyramid and the accumulators of a block of samples
//                      live in one CTA's shared memory; one launch per block of samples.
// The cluster update, the RGNN kernels and the boundary/bookkeeping kernels are in cluster.cu, rgnn.cu, util_kernels.cu.
#include "kernels.cuh"
#include "mcfast.cuh"

// 3 CTAs of 256 threads per SM = 80 registers per thread: the row body of mc_row keeps its constants in registers
// (measured with per-thread staging loads: 0.274 ms per launch of the headline configuration against 0.286 ms with
// 4 CTAs / 64 registers; 0.216 ms today, see DESIGN.md section 3.1 for what was removed since)
#ifndef MCRG_SWEEP_MIN_BLOCKS
#define MCRG_SWEEP_MIN_BLOCKS 3
#endif
// k_level: 128 threads x 32 registers = 4096 registers per CTA, exactly what three resident sweep CTAs leave free on an
// SM (65536 - 3 x 256 x 80), so the pyramid of sample s (side stream) shares SMs with the sweep of sample s+1 instead of
// waiting for one of its CTA slots (measured: 36.8 -> 36.1 ms per 128-sample step)
#ifndef MCRG_LEVEL_THREADS
#define MCRG_LEVEL_THREADS 128
#endif
#ifndef MCRG_LEVEL_MIN_BLOCKS
#define MCRG_LEVEL_MIN_BLOCKS 16
#endif
// one-warp k_resident CTAs per SM (lattices up to 64^2): 28 -> at most 72 registers per thread
#ifndef RESIDENT_SMALL_BLOCKS
#define RESIDENT_SMALL_BLOCKS 28
#endif

namespace mcrg {

namespace {

// Warp-reduce the four counters and add lane 0's totals into shared (or global) 32-bit cells.
__device__ __forceinline__ void warp_reduce_to(const Counts &c, unsigned int *cells) {
    const unsigned int a = __reduce_add_sync(0xFFFFFFFFu, c.anti_nn);
    const unsigned int b = __reduce_add_sync(0xFFFFFFFFu, c.anti_nnn);
    const unsigned int p = __reduce_add_sync(0xFFFFFFFFu, c.odd_plaq);
    const unsigned int u = __reduce_add_sync(0xFFFFFFFFu, c.up);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&cells[0], a);
        atomicAdd(&cells[1], b);
        atomicAdd(&cells[2], p);
        atomicAdd(&cells[3], u);
    }
}

// ---- TMA bulk copies (cp.async.bulk, one instruction per contiguous run of rows) + mbarrier -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
// ---- programmatic dependent launch: a kernel launched with the programmatic-serialization attribute (launch_pdl) may have its
// CTAs scheduled while the previous kernel of the stream is still draining; it must not touch global memory before pdl_wait(),
// which returns once every prerequisite grid has completed and its writes are visible.  pdl_trigger() lets the NEXT kernel's
// CTAs be scheduled as soon as every CTA of this grid has passed it.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

// rows [y_first, y_first + nrows) (periodic) of one plane -> shared memory, issued by ONE thread: one bulk copy per
// contiguous run of rows (two when the strip wraps around the lattice edge).  Returns the bytes put in flight.
__device__ __forceinline__ uint32_t bulk_stage_rows(uint32_t *dst, const uint32_t *src_plane, int y_first, int nrows, int W, int L,
                                                    unsigned long long *bar) {
    int y = y_first, left = nrows;
    uint32_t total = 0;
    while (left > 0) {
        const int n = left < L - y ? left : L - y;
        const uint32_t bytes = (uint32_t)n * (uint32_t)W * 4u;
        bulk_g2s(dst, src_plane + (size_t)y * W, bytes, bar);
        dst += n * W;
        total += bytes;
        left -= n;
        y = 0;
    }
    return total;
}

// Staging entry points used by every kernel: TMA when the rows are 16-byte multiples (L >= 256), else plain loads.
// Protocol: tile_stage_begin (all threads; contains a barrier when TMA is used), tile_stage_plane per plane with the SAME
// total byte count announced up front, tile_stage_wait (all threads), then the caller's __syncthreads().
__device__ __forceinline__ void stage_rows(uint32_t *dst, const uint32_t *src_plane, int y_first, int nrows, int W, int L);

__device__ __forceinline__ bool tile_stage_begin(unsigned long long *bar, int W, uint32_t total_bytes) {
    const bool tma = (W & 3) == 0;
    if (tma) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (threadIdx.x == 0) mbar_expect_tx(bar, total_bytes);
    }
    return tma;
}
__device__ __forceinline__ void tile_stage_plane(bool tma, unsigned long long *bar, uint32_t *dst, const uint32_t *src_plane,
                                                 int y_first, int nrows, int W, int L) {
    if (tma) {
        if (threadIdx.x == 0) bulk_stage_rows(dst, src_plane, y_first, nrows, W, L, bar);
    } else {
        stage_rows(dst, src_plane, y_first, nrows, W, L);
    }
}
__device__ __forceinline__ void tile_stage_wait(bool tma, unsigned long long *bar) {
    if (tma) mbar_wait(bar, 0);  // plain loads need nothing beyond the caller's barrier
}

// rows shorter than 16 bytes (W < 4, i.e. L < 256): plain 4-byte loads, periodic in y (L and W are powers of two)
__device__ __forceinline__ void stage_rows(uint32_t *dst, const uint32_t *src_plane, int y_first, int nrows, int W,
                                           int L) {
    const int lw = ilog2(W), n = nrows << lw;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int lr = idx >> lw, w = idx & (W - 1);
        const int y = (y_first + lr) & (L - 1);
        dst[idx] = __ldg(src_plane + (size_t)y * W + w);
    }
}

__device__ __forceinline__ void unstage_rows(uint32_t *dst_plane, const uint32_t *src, int y_first, int nrows, int W,
                                             int L) {
    if ((W & 3) == 0) {
        const int W4 = W >> 2, l4 = ilog2(W4), n4 = nrows << l4;
        for (int idx = threadIdx.x; idx < n4; idx += blockDim.x) {
            const int lr = idx >> l4, w4 = idx & (W4 - 1);
            const int y = (y_first + lr) & (L - 1);
            reinterpret_cast<uint4 *>(dst_plane + (size_t)y * W)[w4] = reinterpret_cast<const uint4 *>(src)[idx];
        }
    } else {
        const int lw = ilog2(W), n = nrows << lw;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int lr = idx >> lw, w = idx & (W - 1);
            const int y = (y_first + lr) & (L - 1);
            dst_plane[(size_t)y * W + w] = src[idx];
        }
    }
}

// ---- Metropolis update of a strip, device form -----------------------------------------------------------------
// Same decisions as update_word0()/metropolis_flip_mask() in tile.cuh (the scalar specification is the oracle's
// orc_metropolis), organised for the SM:
//   pass 1  every word: neighbour count (full adder), then Philox calls j = 0 and 1 back to back (two independent chains ->
//           ILP; the word-independent part of rounds 0-1 is shared, mc_philox_pair) and 8 lazily-compared bit planes,
//           straight-line, no divergence; the first four planes by code specialised on the leading threshold bits
//           (mc_compare4_nz).  After 8 planes a lane is still undecided with probability 2^-8, i.e. ~10 % of the words
//           keep a few undecided lanes: those words are appended to the warp's shared-memory queue (ballot rank), decided
//           lanes are written back at once.
//   pass 2  the queue is consumed densely, one entry per thread, calls j = 2.. until every lane is decided.
// The per-plane threshold masks (bit k of T4 / T8 replicated over a word) are a 64-entry table in shared memory:
// broadcast LDS on the otherwise idle LSU pipe instead of shifts on the ALU pipe, which is the binding pipe.
struct McTable {
    uint32_t tm[32][2];  // [plane][0: T4 bit, 1: T8 bit] as 0 / 0xFFFFFFFF
};

// The table of the strip / resident kernels is ONE static shared variable referenced by name (mc_table()): its address is an
// immediate of the LDS.  Reached through a generic pointer in a struct, the compiler rebuilt the shared-window address inside
// the row loop (S2R + MOV + LEA per row pair) rather than keep it in a register.  k_resident_multi, whose lanes have different
// tables, passes pointers (mc_compare4 below).
__device__ __forceinline__ McTable &mc_table() {
    __shared__ __align__(16) McTable tab;
    return tab;
}

__device__ __forceinline__ void mc_compare4_s(const U4 &r, uint32_t tab_s, int plane0, uint32_t sel, uint32_t &eq, uint32_t &lt) {
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint32_t t4, t8;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t4), "=r"(t8) : "r"(tab_s + 8u * (uint32_t)(plane0 + e)));
        const uint32_t tm = (sel & t4) | (~sel & t8);
        lt |= eq & ~rr[e] & tm;
        eq &= ~(rr[e] ^ tm);
    }
}

__device__ __forceinline__ void mc_compare4(const U4 &r, const McTable *tab, int plane0, uint32_t sel, uint32_t &eq,
                                            uint32_t &lt) {
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint2 t48 = *reinterpret_cast<const uint2 *>(tab->tm[plane0 + e]);
        const uint32_t tm = (sel & t48.x) | (~sel & t48.y);  // this lane's threshold bit (A==1 lanes: T4, A==0: T8)
        lt |= eq & ~rr[e] & tm;                              // U bit 0 where T bit 1, prefix equal: U < T
        eq &= ~(rr[e] ^ tm);
    }
}

struct McQueue {
    uint4 *ent;      // [warp][cap]: {tile word offset, undecided lanes, selector (A==1 lanes), -}
    int cap;         // entries per warp
};

// finish one word whose lanes `eq` are still undecided after planes [0, 4*j0): calls j0, j0+1, ... (rare path)
__device__ __forceinline__ uint32_t mc_finish(uint32_t eq, uint32_t sel, int j0, const McTable *tab, const McPhiloxHead &h,
                                              uint64_t seed, uint32_t word_id, uint32_t c3_base) {
    uint32_t lt = 0;
    for (int j = j0; j < 8 && eq != 0u; ++j) mc_compare4(mc_philox_j(h, seed, word_id, c3_base, j), tab, 4 * j, sel, eq, lt);
    return lt;
}

__device__ __forceinline__ uint32_t mc_finish_s(uint32_t eq, uint32_t sel, int j0, uint32_t tab_s, const McPhiloxHead &h, uint64_t seed,
                                                uint32_t word_id, uint32_t ck) {
    uint32_t lt = 0;
    for (int j = j0; j < 8 && eq != 0u; ++j) mc_compare4_s(mc_philox_j_ck(h, seed, word_id, ck, j), tab_s, 4 * j, sel, eq, lt);
    return lt;
}

// One half-sweep (colour c) over local rows [lr_lo, lr_lo + nrows).
// Thread layout: column w = tid & (W-1), row group g = tid >> lw; a group owns a CONTIGUOUS block of rows and every
// thread walks down its column.  The other-colour words above / at / below the current row then form a sliding window
// in registers (one new load per row instead of three), all shared-memory addresses are "pointer + constant", and the
// direction of the in-row neighbour shift — which alternates with the row parity — is a template parameter of the
// row body (rows are processed in pairs).  For 32-bit words the shift is one funnel shift.
// Words that still have undecided lanes after the 8 planes of pass 1 are appended to a queue PRIVATE TO THE WARP
// (slot = warp-uniform running count + rank in the ballot: no atomics, no shuffles) and finished densely by the same
// warp — only __syncwarp() between the passes.  Every warp executes the same number of row steps (inactive steps are
// predicated off), so the ballots are full-warp even when a warp spans several row groups (W < 32).
struct McWalk {
    uint32_t *pc;        // this thread's word of the row being updated
    const uint32_t *po;  // the same position in the other colour's plane
    uint32_t u, n0;      // other-colour words of rows lr-1 and lr
    uint32_t yw;         // (y_first + lr) << lw, NOT wrapped: the word id masks it
    uint32_t off;        // word offset of pc inside its plane (queue entries carry it)
    int n_queued;        // warp-uniform
};

struct McConst {
    uint4 *my_q;         // this warp's queue segment: {offset, undecided lanes, selector, -}
    uint32_t tab_s;      // shared-space address of the threshold table, pinned in a register (see mc_half_sweep_t)
    McPhiloxHead head;   // the part of Philox rounds 0 and 1 that is constant over the half-sweep
    uint64_t seed;
    uint32_t replica, t_lo, ck, anti, mask, wid_c, yw_mask, lanes_below;  // ck = c3_base ^ key word 1 (mc_philox_pair_ck)
    int W, bits, d_up, d_dn, qcap, n_act;
    bool w_first, w_last;  // this thread's column is the first / last word of a row
};

// word id of the Philox counter: colour*L*W + y*W + w.  wid_c = colour*L*W + w and (yw & yw_mask) = (y mod L)*W occupy
// disjoint bits, so one LOP3 builds it from the running row counter.
__device__ __forceinline__ uint32_t mc_word_id(const McWalk &k, const McConst &g) { return (k.yw & g.yw_mask) | g.wid_c; }

// NZ < 0: any thresholds;  NZ = 0..3: thresholds with T4 < 1/4 whose planes 2 and 3 are NZ (see mc_compare4_nz)
// CHK: the thread may have fewer rows than the loop runs steps (it >= n_act: step predicated off)
// Updates one word and steps to the next row; eq = its lanes still undecided after pass 1 (0: none), sel = its A == 1 lanes.
template <int WT, int P, bool B32, int NZ, bool CHK>
__device__ __forceinline__ void mc_row(McWalk &k, const McConst &g, int it, uint32_t &eq, uint32_t &sel) {
    eq = 0u;
    sel = 0u;
    if (!CHK || it < g.n_act) {
        const uint32_t t = *k.pc;
        const uint32_t d = k.po[g.W];
        uint32_t n1;
        // in-row neighbour word w +- 1 (periodic): with W known at compile time the two possible offsets are immediates and
        // the choice is a predicate (the compiler otherwise recomputes the offset every row to save a register)
        if (P) {
            const uint32_t nb = WT > 1 ? (g.w_last ? k.po[1 - WT] : k.po[1]) : k.po[g.d_up];
            n1 = B32 ? __funnelshift_r(k.n0, nb, 1) : shift_up_index(k.n0, nb, g.bits, g.mask);
        } else {
            const uint32_t nb = WT > 1 ? (g.w_first ? k.po[WT - 1] : k.po[-1]) : k.po[g.d_dn];
            n1 = B32 ? __funnelshift_l(nb, k.n0, 1) : shift_down_index(k.n0, nb, g.bits, g.mask);
        }
        const uint32_t a1 = t ^ k.u ^ g.anti, a2 = t ^ d ^ g.anti, a3 = t ^ k.n0 ^ g.anti, a4 = t ^ n1 ^ g.anti;
        uint32_t ge2;                       // A >= 2: these lanes flip unconditionally; sel: A == 1
        mc_neighbour_count(a1, a2, a3, a4, ge2, sel);
        eq = ~ge2;                          // A == 1 or A == 0: lanes that need a random number
        if (!B32) {
            ge2 &= g.mask;
            sel &= g.mask;
            eq &= g.mask;
        }
        uint32_t lt = 0;                    // subset of the initial eq, hence disjoint from the A >= 2 lanes
        const uint32_t word_id = mc_word_id(k, g);
        U4 r0, r1;
        mc_philox_pair_ck(g.head, g.seed, word_id, g.ck, r0, r1);
        if (NZ >= 0) mc_compare4_nz<NZ>(r0, sel, eq, lt);
        else mc_compare4_s(r0, g.tab_s, 0, sel, eq, lt);
        mc_compare4_s(r1, g.tab_s, 4, sel, eq, lt);
        *k.pc = t ^ (ge2 | lt);
        k.u = k.n0;
        k.n0 = d;
    }
    k.pc += g.W;
    k.po += g.W;
    k.off += (uint32_t)g.W;
    k.yw += (uint32_t)g.W;
}

// Append the words of this step that keep undecided lanes to the warp's queue; `back` = rows between the word and the walker's
// current position (the walker has already stepped past it).  Every lane of the warp calls this.
__device__ __forceinline__ void mc_push(McWalk &k, const McConst &g, unsigned pend, bool need, int back, uint32_t eq, uint32_t sel) {
    if (need) {
        const int slot = k.n_queued + __popc(pend & g.lanes_below);
        const uint32_t off = k.off - (uint32_t)(back * g.W);
        if (slot < g.qcap) {
            g.my_q[slot] = make_uint4(off, eq, sel, 0u);
        } else {  // segment full (does not happen for equilibrium-like data; kept for exactness): finish inline
            const uint32_t yw = k.yw - (uint32_t)(back * g.W);
            k.pc[-back * g.W] ^= mc_finish_s(eq, sel, 2, g.tab_s, g.head, g.seed, (yw & g.yw_mask) | g.wid_c, g.ck);
        }
    }
    k.n_queued += __popc(pend);
}

// Two rows per step.  ~10 % of the words keep undecided lanes, i.e. almost every warp-row has a few, so the queue push is on the
// critical path of every row: the two rows of a step share ONE push — a lane with one pending word (20 % of the lanes) appends
// it, and only when some lane of the warp has both words pending (a third of the steps) a second push takes those.
template <int WT, int P0, bool B32, int NZ, bool CHK>
__device__ __forceinline__ void mc_walk(McWalk &k, const McConst &g, int n_steps) {
    for (int it = 0; it < n_steps; it += 2) {  // n_steps is even
        uint32_t eq_a, sel_a, eq_b, sel_b;
        mc_row<WT, P0, B32, NZ, CHK>(k, g, it, eq_a, sel_a);
        mc_row<WT, 1 - P0, B32, NZ, CHK>(k, g, it + 1, eq_b, sel_b);
        const bool a = eq_a != 0u, b = eq_b != 0u;
        const unsigned pend = __ballot_sync(0xFFFFFFFFu, a | b);
        mc_push(k, g, pend, a | b, a ? 2 : 1, a ? eq_a : eq_b, a ? sel_a : sel_b);
        const unsigned both = __ballot_sync(0xFFFFFFFFu, a & b);
        if (both != 0u) mc_push(k, g, both, a & b, 1, eq_b, sel_b);  // warp-uniform branch
    }
}

template <int WT, bool B32, int NZ, bool CHK>
__device__ __forceinline__ void mc_walk_par(int par0, McWalk &k, const McConst &g, int n_steps) {
    if (par0) mc_walk<WT, 1, B32, NZ, CHK>(k, g, n_steps);
    else mc_walk<WT, 0, B32, NZ, CHK>(k, g, n_steps);
}

template <int WT, bool CHK>
__device__ __forceinline__ void mc_walk_b32(bool nz, int xy, int par0, McWalk &k, const McConst &g, int n_steps) {
    if (!nz) mc_walk_par<WT, true, -1, CHK>(par0, k, g, n_steps);
    else if (xy == 0) mc_walk_par<WT, true, 0, CHK>(par0, k, g, n_steps);
    else if (xy == 1) mc_walk_par<WT, true, 1, CHK>(par0, k, g, n_steps);
    else if (xy == 2) mc_walk_par<WT, true, 2, CHK>(par0, k, g, n_steps);
    else mc_walk_par<WT, true, 3, CHK>(par0, k, g, n_steps);
}

template <int WT, bool REQUEUE = false>
__device__ __forceinline__ void mc_half_sweep_t(const Strip0 &s, int c, int lr_lo, int nrows, int lw, uint32_t anti,
                                                const McTable *tab, const McQueue &q, uint64_t seed, uint32_t replica,
                                                unsigned long long sweep) {
    const int W = WT > 0 ? WT : s.W, o = 1 - c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *plane_c = s.base + c * s.rows * W;
    const uint32_t *plane_o = s.base + o * s.rows * W;
    McConst g;
    g.my_q = q.ent + q.cap * warp;  // this warp's segment: q.cap entries
    {
        const uint32_t t0 = smem_u32(&mc_table());
        asm volatile("mov.u32 %0, %1;" : "=r"(g.tab_s) : "r"(t0));  // opaque: the compiler cannot rematerialise it inside the row loop
    }
    g.seed = seed;
    g.replica = replica;
    g.t_lo = (uint32_t)sweep;
    g.ck = (((uint32_t)PURPOSE_MC << 28) | (uint32_t)((sweep >> 32) & 0xFFFFFu)) ^ (uint32_t)(seed >> 32);
    g.head = mc_philox_head(seed, replica, g.t_lo);
    g.anti = anti;
    g.mask = s.mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(g.lanes_below));  // one S2R if the compiler rematerialises it (it did: S2R tid + MOV + SHF)
    g.W = W;
    g.bits = s.bits;
    g.qcap = q.cap;
    const int w = threadIdx.x & (W - 1), grp = threadIdx.x >> lw;
    const int n_grp = blockDim.x >> lw;  // blockDim is a multiple of W (launch_sweep0)
    g.w_first = w == 0;
    g.w_last = w == W - 1;
    g.d_up = ((w + 1) & (W - 1)) - w;
    g.d_dn = ((w - 1) & (W - 1)) - w;
    g.wid_c = (uint32_t)(c * s.L * W + w);
    g.yw_mask = (uint32_t)((s.L - 1) << lw);
    // Rows per group.  nrows is even (strips have an even number of rows and lose two per half-sweep; a resident lattice has
    // L rows).  W >= 32, one row group per warp: the row PAIRS are dealt out as evenly as possible, every group gets whole
    // pairs, so a warp runs exactly its own rows with no per-row "do I still have a row" test (`whole_pairs`).  W < 32, several
    // groups per warp: equal even chunks (the groups of a warp must start on rows of the same parity and run the same number
    // of steps for the full-warp ballots), the last groups may run short and test every step.
    const bool whole_pairs = W >= 32;
    int lr0, n_steps;
    if (whole_pairs) {
        const int pairs = nrows >> 1, base = pairs / n_grp, rem = pairs - base * n_grp;
        const int first = grp * base + (grp < rem ? grp : rem), mine = base + (grp < rem ? 1 : 0);
        lr0 = lr_lo + 2 * first;
        n_steps = 2 * mine;  // warp-uniform
        g.n_act = n_steps;
    } else {
        int chunk = (nrows + n_grp - 1) / n_grp;
        chunk = (chunk + 1) & ~1;
        n_steps = chunk;  // identical for every thread
        const int lr_hi = lr_lo + nrows;
        int lr1;
        lr0 = lr_lo + grp * chunk;
        lr1 = lr0 + chunk;
        if (lr1 > lr_hi) lr1 = lr_hi;
        if (lr0 >= lr_hi) lr0 = lr1 = lr_lo;  // nothing to do: park on a valid row
        g.n_act = lr1 - lr0;
    }
    McWalk k;
    k.off = (uint32_t)(lr0 * W + w);
    k.pc = plane_c + k.off;
    k.po = plane_o + k.off;
    k.u = k.po[-W];
    k.n0 = k.po[0];
    k.yw = (uint32_t)((s.y_first + lr0) << lw);
    k.n_queued = 0;
    const int par0 = (s.y_first + lr0 + c) & 1;  // warp-uniform (W >= 32: one group per warp; W < 32: chunk even)
    // thresholds below 1/4 (every coupling of the critical region): cheaper first compare, see mc_compare4_nz
    const McTable &tb = mc_table();  // == *tab: every caller of the strip / resident kernels passes &mc_table()
    const bool nz = (tb.tm[0][0] | tb.tm[1][0] | tb.tm[0][1] | tb.tm[1][1] | tb.tm[2][1] | tb.tm[3][1]) == 0u;
    const int xy = (tb.tm[2][0] ? 2 : 0) | (tb.tm[3][0] ? 1 : 0);
    if (s.bits == 32) {
        if (WT >= 32) mc_walk_b32<WT, false>(nz, xy, par0, k, g, n_steps);        // W is a compile-time constant >= 32
        else if (WT > 0) mc_walk_b32<WT, true>(nz, xy, par0, k, g, n_steps);      // ... < 32
        else if (whole_pairs) mc_walk_b32<0, false>(nz, xy, par0, k, g, n_steps);
        else mc_walk_b32<0, true>(nz, xy, par0, k, g, n_steps);
    } else {
        mc_walk_par<0, false, -1, true>(par0, k, g, n_steps);
    }
    __syncwarp();
    if (!REQUEUE) {  // resident kernels: a warp rarely queues more than one batch per half-sweep; the plain form is faster there
        const int total = min(k.n_queued, q.cap);
        for (int e = lane; e < total; e += 32) {
            const uint4 ent = g.my_q[e];
            const uint32_t yw = ((uint32_t)s.y_first << lw) + (ent.x & ~(uint32_t)(W - 1));
            const uint32_t word_id = (yw & g.yw_mask) | ((uint32_t)(c * s.L * W) + (ent.x & (uint32_t)(W - 1)));
            plane_c[ent.x] ^= mc_finish_s(ent.y, ent.z, 2, g.tab_s, g.head, seed, word_id, g.ck);
        }
    } else {
        // Pass 2, one queue entry per lane.  A batch of 32 entries would run as many Philox calls as its unluckiest entry (after
        // call 2 a lane is still undecided with probability 1/16: 90 % of the full batches need a second round for one or two lanes).
        // So every batch but the last runs exactly ONE call and appends the entries that still have undecided lanes to the tail of
        // the queue (entry word 3 = the next call); they are picked up by the last, partly filled batch, which finishes its entries
        // completely.  Which call decides a lane does not depend on who executes it: same results as mc_finish per entry.
        int total = min(k.n_queued, q.cap);
        for (int base = 0; base < total; base += 32) {
            const int e = base + lane;
            const bool last = total <= base + 32;  // warp-uniform: nothing appended now would be reached by a later batch
            uint32_t eq = 0u, sel = 0u, off = 0u, word_id = 0u;
            int j = 2;
            if (e < total) {
                const uint4 ent = g.my_q[e];
                off = ent.x;
                eq = ent.y;
                sel = ent.z;
                j = 2 + (int)ent.w;
                const uint32_t yw = ((uint32_t)s.y_first << lw) + (off & ~(uint32_t)(W - 1));
                word_id = (yw & g.yw_mask) | ((uint32_t)(c * s.L * W) + (off & (uint32_t)(W - 1)));
                if (last) {
                    plane_c[off] ^= mc_finish_s(eq, sel, j, g.tab_s, g.head, seed, word_id, g.ck);
                    eq = 0u;
                } else {
                    uint32_t lt = 0u;
                    mc_compare4_s(mc_philox_j_ck(g.head, seed, word_id, g.ck, j), g.tab_s, 4 * j, sel, eq, lt);
                    plane_c[off] ^= lt;
                    if (j >= 7) eq = 0u;  // call 7 was the last one (32 planes): whatever is still equal does not flip (U == T)
                }
            }
            if (!last) {
                const unsigned again = __ballot_sync(0xFFFFFFFFu, eq != 0u);
                if (again != 0u) {  // warp-uniform
                    if (eq != 0u) {
                        const int slot = total + __popc(again & g.lanes_below);
                        if (slot < q.cap) g.my_q[slot] = make_uint4(off, eq, sel, (uint32_t)(j + 1 - 2));
                        else plane_c[off] ^= mc_finish_s(eq, sel, j + 1, g.tab_s, g.head, seed, word_id, g.ck);  // segment full (see mc_push)
                    }
                    total = min(total + __popc(again), q.cap);
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void mc_half_sweep(const Strip0 &s, int c, int lr_lo, int nrows, int lw, uint32_t anti,
                                              const McTable *tab, const McQueue &q, uint64_t seed, uint32_t replica,
                                              unsigned long long sweep) {
    if (s.W == 64) mc_half_sweep_t<64>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);  // L = 4096
    else mc_half_sweep_t<0>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);
}

// the strip kernel's dispatcher: the row body with the words per row as a compile-time constant (immediate neighbour offsets)
// for the named lattice sizes L = 1024, 4096, 16384; any other size takes the general body
__device__ __forceinline__ void mc_half_sweep_strip(const Strip0 &s, int c, int lr_lo, int nrows, int lw, uint32_t anti,
                                                    const McTable *tab, const McQueue &q, uint64_t seed, uint32_t replica,
                                                    unsigned long long sweep) {
    if (s.W == 64) mc_half_sweep_t<64, true>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);
#if !defined(MCRG_FEWER_INSTANTIATIONS)
    else if (s.W == 256) mc_half_sweep_t<256, true>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);
    else if (s.W == 16) mc_half_sweep_t<16, true>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);
#endif
    else mc_half_sweep_t<0, true>(s, c, lr_lo, nrows, lw, anti, tab, q, seed, replica, sweep);
}

// MEASURE phase of a strip whose words are full (bits == 32, L >= 64): the same counts and block words as measure_pair0
// (tile.cuh), organised like the sweep.  A thread keeps ONE column w and visits the pair rows i0, i0 + di, ... (item
// index = threadIdx.x + k * blockDim.x, so consecutive items of a thread are 256 output words apart and share a
// tie-coin Philox call four at a time, see tie_group); every shared-memory address is "pointer + constant", the in-row
// neighbours are one funnel shift each and the periodic wrap of w +- 1 is a per-thread constant offset.
struct MeasureAcc {
    uint32_t nn, nnn, pq, up;
};

// one row-pair word: counts into `m`, returns the majority word and the tie mask of the 32 blocks
template <int WT>
__device__ __forceinline__ uint32_t measure_item_b32(const uint32_t *pb, const uint32_t *pw, int Wrt, int d_up, int d_dn,
                                                     MeasureAcc &m, uint32_t &tie) {
    const int W = WT > 0 ? WT : Wrt;
    const uint32_t b0 = pb[0], b1 = pb[W], b2 = pb[2 * W];
    const uint32_t w0 = pw[0], w1 = pw[W], w2 = pw[2 * W];
    const uint32_t b0u = __funnelshift_r(b0, pb[d_up], 1);          // black row y, index x'+1
    const uint32_t w1u = __funnelshift_r(w1, pw[W + d_up], 1);
    const uint32_t b2u = __funnelshift_r(b2, pb[2 * W + d_up], 1);
    const uint32_t b1d = __funnelshift_l(pb[W + d_dn], b1, 1);      // black row y+1, index x'-1
    const uint32_t w2d = __funnelshift_l(pw[2 * W + d_dn], w2, 1);
    const uint32_t e00 = b0 ^ w0, e11 = w1 ^ b1;                    // shared by the bond and the plaquette words
    const uint32_t f0 = w0 ^ b0u, f1 = b1 ^ w1u;
    // even row y: black sites x = 2x', white sites x = 2x'+1
    m.nn += __popc(e00) + __popc(f0) + __popc(b0 ^ w1) + __popc(w0 ^ b1);
    m.nnn += __popc(b0 ^ b1) + __popc(b0 ^ b1d) + __popc(w0 ^ w1u) + __popc(w0 ^ w1);
    m.pq += __popc(e00 ^ e11) + __popc(f0 ^ f1);
    // odd row y+1: black sites x = 2x'+1, white sites x = 2x'
    m.nn += __popc(f1) + __popc(e11) + __popc(b1 ^ w2) + __popc(w1 ^ b2);
    m.nnn += __popc(b1 ^ b2u) + __popc(b1 ^ b2) + __popc(w1 ^ w2) + __popc(w1 ^ w2d);
    m.pq += __popc(f1 ^ w2 ^ b2u) + __popc(e11 ^ b2 ^ w2);
    m.up += __popc(b0) + __popc(w0) + __popc(b1) + __popc(w1);
    uint32_t maj;
    majority4(b0, w0, b1, w1, maj, tie);
    return maj;
}

template <int WT, bool FAST4>
__device__ __forceinline__ void measure_strip_b32(const Strip0 &s, int H, int R, int lw, int y0, uint32_t *lev1, uint64_t seed,
                                                  uint32_t replica, unsigned long long t, Counts &cnt,
                                                  const uint32_t *tie_src = nullptr) {
    const int W = WT > 0 ? WT : s.W;
    const int w = threadIdx.x & (W - 1), i0 = threadIdx.x >> lw, di = blockDim.x >> lw;  // blockDim is a multiple of W
    const int d_up = ((w + 1) & (W - 1)) - w, d_dn = ((w - 1) & (W - 1)) - w;
    const uint32_t *pb = s.base + (H + 2 * i0) * W + w;  // black plane, first row of the pair
    const uint32_t *pw = pb + s.rows * W;                // white plane
    const int step = 2 * di * W;
    uint32_t q = (uint32_t)(((y0 >> 1) + i0) << lw) + (uint32_t)w;
    const uint32_t dq = (uint32_t)(di << lw);
    MeasureAcc m = {0u, 0u, 0u, 0u};
    const int half = R >> 1;
    if (FAST4 && blockDim.x == 256 && tie_src == nullptr) {
        // The CTA's 256 threads x 4 items cover one 1024-word chunk of level-1 output (4 * di pair rows): a thread's items are
        // the words q = 1024 c + threadIdx.x + 256 e, e = 0..3 — exactly the four words that share one tie-coin call
        // (tie_group), in the order x, y, z, w.  Chunks are aligned to the LATTICE (not to the strip), so a strip of any even
        // height and any first row works: items whose pair row falls outside the strip are skipped (strip edges only; the
        // test is warp-uniform for W >= 32).
        const int P0 = y0 >> 1, chunk = 4 * di;
        const int c_first = P0 / chunk, c_last = (P0 + half - 1) / chunk;
        int pl = c_first * chunk + i0 - P0;  // this thread's pair row relative to the strip; may start negative
        pb += 2 * (pl - i0) * W;             // (not dereferenced while out of range)
        pw += 2 * (pl - i0) * W;
        q = (uint32_t)(c_first * 1024) + threadIdx.x;
        for (int c = c_first; c <= c_last; ++c, q += 1024u) {
            const U4 r = philox_keyed(seed, tie_group(q), replica, t, PURPOSE_TIE, 1);
            const uint32_t coin[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e, pb += step, pw += step, pl += di) {
                if ((unsigned)pl >= (unsigned)half) continue;
                uint32_t tie;
                const uint32_t maj = measure_item_b32<WT>(pb, pw, W, d_up, d_dn, m, tie);
                lev1[q + 256u * e] = maj | (tie & coin[e]);
            }
        }
    } else {
        TieCache coins;
        coins.init();
        for (int i = i0; i < half; i += di, pb += step, pw += step, q += dq) {
            uint32_t tie;
            uint32_t maj = measure_item_b32<WT>(pb, pw, W, d_up, d_dn, m, tie);
            if (tie) maj |= tie & (tie_src ? tie_src[q] : coins.get(seed, q, replica, t, 1));
            lev1[q] = maj;
        }
    }
    cnt.anti_nn += m.nn;
    cnt.anti_nnn += m.nnn;
    cnt.odd_plaq += m.pq;
    cnt.up += m.up;
}

// Correlator popcounts of rows [0, n_rows) of a natural-layout lattice with full words (Ln >= 32) that sits in shared
// memory at x[lr * Wn + w]: the same counts as measure_rowN (tile.cuh), organised like the sweep.  A thread keeps one
// column w and walks down a contiguous block of rows; the row below and its two in-row shifts slide along in registers
// (one row = 3 loads, 2 funnel shifts, 5 XORs, 6 popcounts), the periodic wrap of w +- 1 is a per-thread constant
// offset.  The row below the last one is row n_rows (WRAP = false: a staged halo row) or row 0 (WRAP = true: the whole
// periodic lattice is in shared memory).  Any block size; columns beyond blockDim are visited in further passes.
template <bool WRAP>
__device__ __forceinline__ void measure_rows_b32(const uint32_t *x, int Wn, int n_rows, Counts &cnt) {
    const int cols = Wn < (int)blockDim.x ? Wn : (int)blockDim.x;  // both are powers of two or multiples of 32 >= Wn
    if ((blockDim.x % cols) != 0) {  // (not reached with the block sizes used here; kept for exactness)
        StripN s;
        s.x = x;
        s.W = Wn;
        s.bits = 32;
        s.mask = 0xFFFFFFFFu;
        for (int idx = threadIdx.x; idx < n_rows * Wn; idx += blockDim.x) {
            const int lr = idx / Wn;
            measure_rowN(s, lr, (WRAP && lr + 1 == n_rows) ? 0 : lr + 1, idx - lr * Wn, cnt);
        }
        return;
    }
    const int n_grp = blockDim.x / cols, grp = threadIdx.x / cols;
    const int chunk = (n_rows + n_grp - 1) / n_grp;
    const int lr0 = grp * chunk, lr1 = min(n_rows, lr0 + chunk);
    uint32_t nn = 0, nnn = 0, pq = 0, up = 0;
    for (int w = threadIdx.x & (cols - 1); w < Wn && lr0 < lr1; w += cols) {
        const int d_up = ((w + 1) & (Wn - 1)) - w, d_dn = ((w - 1) & (Wn - 1)) - w;
        const uint32_t *p = x + lr0 * Wn + w;
        uint32_t r0 = p[0], r0u = __funnelshift_r(r0, p[d_up], 1);
        for (int lr = lr0; lr < lr1; ++lr) {
            p = (WRAP && lr + 1 == n_rows) ? x + w : p + Wn;
            const uint32_t r1 = p[0];
            const uint32_t r1u = __funnelshift_r(r1, p[d_up], 1), r1d = __funnelshift_l(p[d_dn], r1, 1);
            const uint32_t h = r0 ^ r0u;
            nn += __popc(h) + __popc(r0 ^ r1);
            nnn += __popc(r0 ^ r1u) + __popc(r0 ^ r1d);
            pq += __popc(h ^ r1 ^ r1u);
            up += __popc(r0);
            r0 = r1;
            r0u = r1u;
        }
    }
    cnt.anti_nn += nn;
    cnt.anti_nnn += nnn;
    cnt.odd_plaq += pq;
    cnt.up += up;
}

template <bool MEASURE>
__global__ void __launch_bounds__(SWEEP_THREADS, MCRG_SWEEP_MIN_BLOCKS) k_sweep0(const SweepArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ unsigned int red[4];
    McTable &tab = mc_table();
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.y, strip = blockIdx.x;
    const int L = a.L, W = a.W, lw = ilog2(W);
    const int y0 = strip * a.R;
    const int Rs = min(a.R, L - y0);  // R need not divide L: the last strip takes what is left (even, like R and L)
    const int rows = Rs + 2 * a.H;
    Strip0 s;
    s.base = smem;
    s.rows = rows;
    s.W = W;
    s.bits = a.bits;
    s.mask = valid_mask(a.bits);
    s.L = L;
    s.y_first = (y0 - a.H) & (L - 1);
    const uint32_t *src_r = a.src + (size_t)r * 2 * L * W;
    // staging: TMA bulk copies signalled through an mbarrier when rows are 16-byte multiples (L >= 256), else plain loads
    const bool tma = tile_stage_begin(&bar, W, 2u * (uint32_t)rows * (uint32_t)W * 4u);
    if (MEASURE && threadIdx.x < 4) red[threadIdx.x] = 0;
    pdl_wait();  // nothing above reads or writes global memory
    tile_stage_plane(tma, &bar, s0_plane(s, 0), src_r, s.y_first, rows, W, L);
    tile_stage_plane(tma, &bar, s0_plane(s, 1), src_r + (size_t)L * W, s.y_first, rows, W, L);
    if (a.nsw > 0) {
        for (int k = threadIdx.x; k < 64; k += blockDim.x) {  // blockDim may be as small as 32
            const uint32_t T = (k & 1) ? a.T8[r] : a.T4[r];
            tab.tm[k >> 1][k & 1] = ((T >> (31 - (k >> 1))) & 1u) ? 0xFFFFFFFFu : 0u;
        }
    }
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    tile_stage_wait(tma, &bar);
    __syncthreads();

    if (MEASURE) {
        Counts c = {0u, 0u, 0u, 0u};
        const int npairs = (Rs >> 1) << lw;
        uint32_t *lev1 = a.level1 + (size_t)r * (L >> 1) * W;
        const uint32_t *tie_src = a.ties ? a.ties + (size_t)r * a.tie_stride : nullptr;  // level 1 comes first
        if (a.bits == 32) {  // L >= 64
            if (W == 64) measure_strip_b32<64, true>(s, a.H, Rs, lw, y0, lev1, a.seed, replica, t, c, tie_src);  // L = 4096
#if !defined(MCRG_FEWER_INSTANTIATIONS)
            else if (W == 256) measure_strip_b32<256, true>(s, a.H, Rs, lw, y0, lev1, a.seed, replica, t, c, tie_src);  // L = 16384
            else if (W == 16) measure_strip_b32<16, true>(s, a.H, Rs, lw, y0, lev1, a.seed, replica, t, c, tie_src);    // L = 1024
#endif
            else measure_strip_b32<0, true>(s, a.H, Rs, lw, y0, lev1, a.seed, replica, t, c, tie_src);
        } else {
            TieCache coins;
            coins.init();
            for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
                const int i = idx >> lw, w = idx & (W - 1);
                uint32_t maj, tie;
                measure_pair0(s, a.H + 2 * i, w, c, maj, tie);
                const uint32_t q = (uint32_t)(((y0 >> 1) + i) << lw) + (uint32_t)w;
                uint32_t out = maj;
                if (tie) out |= tie & (tie_src ? tie_src[q] : coins.get(a.seed, q, replica, t, 1));
                lev1[q] = out;
            }
        }
        warp_reduce_to(c, red);
        __syncthreads();
        if (threadIdx.x < 4)
            atomicAdd(&a.cnt[((size_t)r * (MAX_LEVELS + 1) + 0) * 4 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
    }

    if (a.nsw > 0) {
        McQueue q;
        q.ent = reinterpret_cast<uint4 *>(smem + ((2 * rows * W + 3) & ~3));
        q.cap = sweep0_queue_cap(rows * W, blockDim.x >> 5);
        const uint32_t anti = a.anti[r];
        for (int h = 0; h < 2 * a.nsw; ++h)
            mc_half_sweep_strip(s, h & 1, 1 + h, rows - 2 - 2 * h, lw, anti, &tab, q, a.seed, replica,
                          t + (unsigned long long)(h >> 1));
        uint32_t *dst_r = a.dst + (size_t)r * 2 * L * W;
        pdl_trigger();  // the next kernel of the stream may be scheduled from here on (see launch_sweep0)
        if (tma) {  // the R rows of a strip are contiguous in global memory: one bulk store per colour plane
            fence_proxy_async();  // this thread's shared-memory writes become visible to the copy engine ...
            __syncthreads();      // ... and so do everybody else's
            if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)Rs * (uint32_t)W * 4u;
                bulk_s2g(dst_r + (size_t)y0 * W, s0_plane(s, 0) + a.H * W, bytes);
                bulk_s2g(dst_r + (size_t)L * W + (size_t)y0 * W, s0_plane(s, 1) + a.H * W, bytes);
                bulk_commit_wait_read();  // shared memory must stay alive until the engine has read it
            }
        } else {
            unstage_rows(dst_r, s0_plane(s, 0) + a.H * W, y0, Rs, W, L);
            unstage_rows(dst_r + (size_t)L * W, s0_plane(s, 1) + a.H * W, y0, Rs, W, L);
        }
    }
}

__global__ void __launch_bounds__(MCRG_LEVEL_THREADS, MCRG_LEVEL_MIN_BLOCKS) k_level(const LevelArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ unsigned int red[4];
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.y, strip = blockIdx.x;
    const int Ln = a.Ln, Wn = nat_words(Ln), lw = ilog2(Wn);
    const int y0 = strip * a.R;
    StripN s;
    s.x = smem;
    s.W = Wn;
    s.bits = nat_bits(Ln);
    s.mask = valid_mask(s.bits);
    pdl_trigger();
    const bool tma = tile_stage_begin(&bar, Wn, (uint32_t)(a.R + 1) * (uint32_t)Wn * 4u);
    if (threadIdx.x < 4) red[threadIdx.x] = 0;
    pdl_wait();
    tile_stage_plane(tma, &bar, smem, a.in + (size_t)r * Ln * Wn, y0, a.R + 1, Wn, Ln);
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    tile_stage_wait(tma, &bar);
    __syncthreads();

    Counts c = {0u, 0u, 0u, 0u};
    measure_rows_b32<false>(smem, Wn, a.R, c);  // Ln > TAIL_MAX_L here: full words; row R is the staged halo row
    if (a.out != nullptr) {
        const int Lb = Ln >> 1, Wb = nat_words(Lb), lwb = ilog2(Wb);
        uint32_t *out_r = a.out + (size_t)r * Lb * Wb;
        const uint32_t *tie_src = a.ties ? a.ties + (size_t)r * a.tie_stride + tie_level_off(a.L, a.level + 1) : nullptr;
        const int nb = (a.R >> 1) << lwb;
        TieCache coins;
        coins.init();
        // Output words q0 .. q0+nb-1, visited so that a thread does the four words that share one Philox call (q, q+256,
        // q+512, q+768 inside a 1024-aligned chunk, see tie_group) back to back, whatever the block size.
        const uint32_t q0 = (uint32_t)((y0 >> 1) << lwb), q1 = q0 + (uint32_t)nb;
        for (uint32_t qc = q0 & ~1023u; qc < q1; qc += 1024u)
            for (uint32_t p = threadIdx.x; p < 256u; p += blockDim.x)
#pragma unroll
                for (uint32_t e = 0; e < 4u; ++e) {
                    const uint32_t q = qc + (e << 8) + p;
                    if (q < q0 || q >= q1) continue;
                    const int idx = (int)(q - q0), i = idx >> lwb, wb = idx & (Wb - 1);
                    uint32_t maj, tie;
                    block_pairN(s, 2 * i, wb, maj, tie);
                    uint32_t o = maj;
                    if (tie) o |= tie & (tie_src ? tie_src[q] : coins.get(a.seed, q, replica, t, a.level + 1));
                    out_r[q] = o;
                }
    }
    warp_reduce_to(c, red);
    __syncthreads();
    if (threadIdx.x < 4)
        atomicAdd(&a.cnt[((size_t)r * (MAX_LEVELS + 1) + a.level) * 4 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
}

// ---- pieces shared by k_tail and k_resident ---------------------------------------------------------------------

// Levels lv0..n_levels with at most 32 rows of one word each, held one row per lane in registers of warp 0: neighbours
// by shuffle, counters by warp reduction, no block barrier until the end.  Same formulas as measure_rowN / block_pairN
// (W == 1: a row's horizontal neighbours wrap inside its own word), same tie-coin keys.  All threads of the CTA call
// this; ends synchronised.
__device__ __forceinline__ void pyramid_tail_warp(const uint32_t *cur, int L, int lv0, int n_levels, unsigned int *red,
                                                  uint32_t *levels_out, const size_t *level_off, int r, uint64_t seed,
                                                  uint32_t replica, unsigned long long t, const uint32_t *ties) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int Ln = L >> lv0;
        uint32_t row = lane < Ln ? cur[lane] : 0u;
        // Tie coins of ALL blocking steps of the tail in one Philox call per lane: step lv -> lv+1 has (L >> lv) / 2 output
        // rows (each one word, its own call: tie_group(q) = q below 256); laid back to back they occupy at most Ln - 1 <= 31
        // lanes.  The consumer of step lv (lane 2q) fetches the coins of row q from lane coin_base + q by shuffle.
        uint32_t coin = 0u;
        {
            int base = 0, my_lv = -1, my_q = 0;
            for (int lv = lv0, n = Ln >> 1; lv < n_levels; ++lv, n >>= 1) {
                if (lane >= base && lane < base + n) {
                    my_lv = lv + 1;
                    my_q = lane - base;
                }
                base += n;
            }
            if (my_lv >= 0) coin = ties ? ties[tie_level_off(L, my_lv) + (size_t)my_q] : tie_word(seed, (uint32_t)my_q, replica, t, my_lv);
        }
        int coin_base = 0;
        for (int lv = lv0; lv <= n_levels; ++lv, Ln >>= 1) {
            const int bits = Ln;  // nat_bits(Ln) for Ln <= 32
            const uint32_t mask = valid_mask(bits);
            const int below = (lane + 1 >= Ln) ? 0 : lane + 1;
            const uint32_t r0 = row, r1 = __shfl_sync(0xFFFFFFFFu, row, below);
            Counts c = {0u, 0u, 0u, 0u};
            if (lane < Ln) {
                const uint32_t r0u = shift_up_index(r0, r0, bits, mask), r1u = shift_up_index(r1, r1, bits, mask);
                const uint32_t r1d = shift_down_index(r1, r1, bits, mask);
                c.anti_nn = popc32(r0 ^ r0u) + popc32(r0 ^ r1);
                c.anti_nnn = popc32(r0 ^ r1u) + popc32(r0 ^ r1d);
                c.odd_plaq = popc32(r0 ^ r0u ^ r1 ^ r1u);
                c.up = popc32(r0);
            }
            warp_reduce_to(c, red + lv * 4);
            if (lv < n_levels) {
                const int Lb = Ln >> 1;
                const uint32_t cw = __shfl_sync(0xFFFFFFFFu, coin, (coin_base + (lane >> 1)) & 31);
                coin_base += Lb;
                uint32_t o = 0u;
                if (lane < Ln && !(lane & 1)) {  // rows (lane, lane + 1) -> block row lane / 2
                    uint32_t m, tie;
                    majority4(r0, r0 >> 1, r1, r1 >> 1, m, tie);
                    o = compress_even(m) | (compress_even(tie) & cw);
                }
                row = __shfl_sync(0xFFFFFFFFu, o, (2 * lane) & 31);  // block row i was computed by lane 2i
                if (lane >= Lb) row = 0u;
                if (levels_out && lane < Lb) levels_out[level_off[lv + 1] + (size_t)r * Lb + lane] = row;
            }
        }
    }
    __syncthreads();
}

// Levels start..n_levels of one replica, level `start` already in `cur` (natural layout): correlator popcounts of
// every level into red[lv*4..], blocking to the next level ping-ponging between cur and nxt.  Optionally mirrors
// every produced level to global memory.  All threads of the CTA call this; ends synchronised.
__device__ __forceinline__ void pyramid_in_smem(uint32_t *cur, uint32_t *nxt, int L, int start, int n_levels, unsigned int *red,
                                                uint32_t *levels_out, const size_t *level_off, int r, uint64_t seed,
                                                uint32_t replica, unsigned long long t, const uint32_t *ties = nullptr) {
    for (int lv = start; lv <= n_levels; ++lv) {
        const int Ln = L >> lv, Wn = nat_words(Ln), lw = ilog2(Wn);
        if (Ln <= 32) {  // one word per row and at most 32 rows: the rest of the pyramid is one warp's business
            pyramid_tail_warp(cur, L, lv, n_levels, red, levels_out, level_off, r, seed, replica, t, ties);
            return;
        }
        StripN s;
        s.x = cur;
        s.W = Wn;
        s.bits = nat_bits(Ln);
        s.mask = valid_mask(s.bits);
        Counts c = {0u, 0u, 0u, 0u};
        measure_rows_b32<true>(cur, Wn, Ln, c);  // Ln >= 64 here: full words, the whole periodic lattice is in `cur`
        warp_reduce_to(c, red + lv * 4);
        if (lv < n_levels) {
            const int Lb = Ln >> 1, Wb = nat_words(Lb), lwb = ilog2(Wb);
            uint32_t *glob = levels_out ? levels_out + level_off[lv + 1] + (size_t)r * Lb * Wb : nullptr;
            const int nb = Lb << lwb;
            const uint32_t *tie_src = ties ? ties + tie_level_off(L, lv + 1) : nullptr;
            TieCache coins;
            coins.init();
            for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
                const int yb = idx >> lwb, wb = idx & (Wb - 1);
                uint32_t maj, tie;
                block_pairN(s, 2 * yb, wb, maj, tie);
                uint32_t o = maj;
                if (tie) o |= tie & (tie_src ? tie_src[idx] : coins.get(seed, (uint32_t)idx, replica, t, lv + 1));
                nxt[idx] = o;
                if (glob) glob[idx] = o;
            }
        }
        __syncthreads();
        uint32_t *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
}

__device__ __forceinline__ void add128(unsigned long long *lo, long long *hi, __int128 v) {
    const unsigned long long vlo = (unsigned long long)v;
    const long long vhi = (long long)(v >> 64);
    const unsigned long long old = *lo;
    const unsigned long long nl = old + vlo;
    *lo = nl;
    *hi = *hi + vhi + (nl < old ? 1 : 0);
}

// The same with atomics, for sums that several kernels in flight add to (the k_tail launches of up to four samples): every
// addition to the low word sees a consistent old value, so the carries it hands to the high word add up to exactly the number of
// wraps — the final 128-bit sum is exact whatever the interleaving; it is read only after all of them are done.
__device__ __forceinline__ void atomic_add128(unsigned long long *lo, long long *hi, __int128 v) {
    const unsigned long long vlo = (unsigned long long)v;
    const long long vhi = (long long)(v >> 64);
    const unsigned long long old = atomicAdd(lo, vlo);
    const long long carry = (old + vlo) < old ? 1 : 0;
    if (vhi + carry != 0) atomicAdd(reinterpret_cast<unsigned long long *>(hi), (unsigned long long)(vhi + carry));
}

// The accumulator slots that are live for a pyramid of n_levels blocking levels, in a compact order:
// k -> (slot index in the public layout, this sample's contribution).  S_sh[lv*4 + {nn, nnn, plaq, sum}].
// mcrg.cpp:86-97 with the column-major flatten of definitions.cpp:9-19 (index b*NOP+a holds X_a * Y_b).
constexpr int ACC_FIXED = 6;  // N, |M|, M^2 and the three M^4 parts
__device__ __forceinline__ int acc_live_slots(int n_levels) { return ACC_FIXED + (NOP + NOP * NOP) * (n_levels + 1) + 2 * NOP * NOP * n_levels; }

// Slot k contributes X[ia] * X[ib] per sample, X being S_sh extended by pseudo-entries: X_ONE = 1, X_ABSM = |M|, and the two
// halves of M^2 = X_M2H * 2^20 + X_M2L (the exact sum of M^4 is kept as the three products of the halves, see SLOT_M4).
constexpr int X_ONE = (MAX_LEVELS + 1) * 4, X_ABSM = X_ONE + 1, X_M2H = X_ABSM + 1, X_M2L = X_M2H + 1, X_LEN = X_M2L + 1;

__device__ __forceinline__ void acc_slot_decode(int k, int n_levels, int &slot, int &ia, int &ib) {
    if (k < 3) {
        slot = k;  // SLOT_N, SLOT_ABSM, SLOT_M2;  M = S_sh[3]
        ia = k == 0 ? X_ONE : (k == 1 ? X_ABSM : 3);
        ib = k == 2 ? 3 : X_ONE;
        return;
    }
    if (k < ACC_FIXED) {
        slot = SLOT_M4 + (k - 3);  // h*h, h*l, l*l
        ia = k == 5 ? X_M2L : X_M2H;
        ib = k == 3 ? X_M2H : X_M2L;
        return;
    }
    k -= ACC_FIXED;
    if (k < NOP * (n_levels + 1)) {
        const int lv = k / NOP, op = k - lv * NOP;
        slot = SLOT_S + lv * NOP + op;
        ia = lv * 4 + op;
        ib = X_ONE;
        return;
    }
    k -= NOP * (n_levels + 1);
    if (k < NOP * NOP * (n_levels + 1)) {
        const int lv = k / (NOP * NOP), e = k - lv * NOP * NOP, b = e / NOP, al = e - b * NOP;
        slot = SLOT_SS + lv * NOP * NOP + e;
        ia = lv * 4 + al;
        ib = lv * 4 + b;
        return;
    }
    k -= NOP * NOP * (n_levels + 1);
    const bool vs_prev = k < NOP * NOP * n_levels;  // S(n) x S(n-1), then S(n) x S(0)
    if (!vs_prev) k -= NOP * NOP * n_levels;
    const int n1 = k / (NOP * NOP), e = k - n1 * NOP * NOP, b = e / NOP, al = e - b * NOP, n = n1 + 1;
    slot = (vs_prev ? SLOT_SBS : SLOT_SB0) + n1 * NOP * NOP + e;
    ia = n * 4 + al;
    ib = (vs_prev ? n - 1 : 0) * 4 + b;
}

__device__ __forceinline__ long long acc_x(const long long *S_sh, int i) {
    if (i == X_ONE) return 1;
    if (i == X_ABSM) return S_sh[3] < 0 ? -S_sh[3] : S_sh[3];
    if (i == X_M2H) return (S_sh[3] * S_sh[3]) >> M4_SPLIT_BITS;  // |M| <= 2^28: M^2 fits
    if (i == X_M2L) return (S_sh[3] * S_sh[3]) & ((1ll << M4_SPLIT_BITS) - 1);
    return S_sh[i];
}

__device__ __forceinline__ void acc_slot_value(int k, int n_levels, const long long *S_sh, int &slot, __int128 &v) {
    int ia, ib;
    acc_slot_decode(k, n_levels, slot, ia, ib);
    v = (__int128)acc_x(S_sh, ia) * acc_x(S_sh, ib);
}

__global__ void __launch_bounds__(256) k_tail(const TailArgs a) {
    __shared__ __align__(16) uint32_t bufA[TAIL_MAX_L * (TAIL_MAX_L / 32)];
    __shared__ __align__(16) uint32_t bufB[(TAIL_MAX_L / 2) * (TAIL_MAX_L / 64)];
    __shared__ unsigned int red[(MAX_LEVELS + 1) * 4];
    __shared__ long long S_sh[(MAX_LEVELS + 1) * 4];
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.x;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    pdl_trigger();
    for (int k = threadIdx.x; k < (MAX_LEVELS + 1) * 4; k += blockDim.x) red[k] = 0;
    pdl_wait();
    const unsigned long long t = *a.d_t + a.t_off;
    if (a.start <= a.n_levels) {  // uniform for the CTA
        const int Ln = a.L >> a.start, Wn = nat_words(Ln);
        const bool tma = tile_stage_begin(&bar, Wn, (uint32_t)Ln * (uint32_t)Wn * 4u);
        tile_stage_plane(tma, &bar, bufA, a.in + (size_t)r * Ln * Wn, 0, Ln, Wn, Ln);
        tile_stage_wait(tma, &bar);
    }
    __syncthreads();
    pyramid_in_smem(bufA, bufB, a.L, a.start, a.n_levels, red, a.levels_out, a.level_off, r, a.seed, replica, t,
                    a.ties ? a.ties + (size_t)r * a.tie_stride : nullptr);
    // raw popcounts -> the reference's sums; levels below `start` were counted by k_sweep0 / k_level
    if (threadIdx.x <= a.n_levels) {
        const int lv = threadIdx.x;
        unsigned long long q[4];
        for (int k = 0; k < 4; ++k) {
            if (lv < a.start) {
                unsigned long long *g = &a.cnt[((size_t)r * (MAX_LEVELS + 1) + lv) * 4 + k];
                q[k] = *g;
                *g = 0ull;
            } else {
                q[k] = red[lv * 4 + k];
            }
        }
        long long S[4];
        counts_to_S((long long)(a.L >> lv), q[0], q[1], q[2], q[3], S);
        for (int k = 0; k < 4; ++k) {
            S_sh[lv * 4 + k] = S[k];
            a.S_out[((size_t)r * (MAX_LEVELS + 1) + lv) * 4 + k] = S[k];
        }
    }
    __syncthreads();
    if (!a.accumulate) return;
    const size_t base = ((size_t)r * a.n_bins + a.bin) * N_SLOTS;
    const int n_live = acc_live_slots(a.n_levels);
    for (int k = threadIdx.x; k < n_live; k += blockDim.x) {
        int slot;
        __int128 v;
        acc_slot_value(k, a.n_levels, S_sh, slot, v);
        atomic_add128(&a.acc_lo[base + slot], &a.acc_hi[base + slot], v);
    }
}

// ---- resident kernel: a whole replica (L <= RESIDENT_MAX_L) lives in one CTA's shared memory -------------------------
// One launch = n_samples x { [measure level 0 + block, pyramid in shared memory, accumulate], m Metropolis sweeps }.
// Local row lr holds global row y = lr-1; rows 0 and L+1 are periodic halo copies, refreshed after every half-sweep
// (2W words) instead of recomputed.  Global memory is touched at the start (load), at the end (store, accumulator
// flush) and nowhere in between; the accumulators of the launch live in shared memory as exact 128-bit sums.
//
// Block size = the number of column walkers a half-sweep can keep busy (resident_threads).  Lattices up to 64^2 are one
// warp's work (32 walkers x 2 rows): SMALL instantiates the kernel for one-warp CTAs with a register budget that lets
// RESIDENT_SMALL_BLOCKS of them share an SM, so that 4096 replicas of 64^2 (BASELINE config 2) are ONE wave of CTAs
// (148 x 28 = 4144) and every scheduler has 7 warps to hide the Philox dependency chains behind.
template <bool MEASURE, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 32 : SWEEP_THREADS, SMALL ? RESIDENT_SMALL_BLOCKS : MCRG_SWEEP_MIN_BLOCKS)
    k_resident(const ResidentArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ unsigned int red[(MAX_LEVELS + 1) * 4];
    __shared__ long long S_sh[X_LEN];  // the sums of the sample + the two pseudo-entries of acc_slot_decode
    McTable &tab = mc_table();
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.x;
    const int L = a.L, W = a.W, lw = ilog2(W);
    const int rows = L + 2;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    const unsigned long long t0 = *a.d_t + a.t_off;
    Strip0 s;
    s.base = smem;
    s.rows = rows;
    s.W = W;
    s.bits = a.bits;
    s.mask = valid_mask(a.bits);
    s.L = L;
    s.y_first = L - 1;
    const ResidentLayout lay = resident_layout(L, blockDim.x, a.n_levels);
    McQueue q;
    q.ent = reinterpret_cast<uint4 *>(smem + lay.queue_off);
    q.cap = lay.cap;
    uint32_t *bufA = smem + lay.bufA_off, *bufB = smem + lay.bufB_off;
    // Accumulators of this launch: 64-bit sums are exact here — |S| <= 4 L^2 <= 2^20, products <= 2^40, and the host
    // layer splits runs into launches of at most 2^22 samples — and flushed into the 128-bit global sums at the end.
    // The slot decoding (which two sums a slot multiplies) is tabulated once per launch.
    long long *acc64 = reinterpret_cast<long long *>(smem + lay.acc_off);
    const int n_live = acc_live_slots(a.n_levels);
    int2 *dec = reinterpret_cast<int2 *>(acc64 + n_live);  // {public slot, ia | ib << 8}

    uint32_t *gl = a.planes + (size_t)r * 2 * L * W;
    const bool tma = tile_stage_begin(&bar, W, 2u * (uint32_t)rows * (uint32_t)W * 4u);
    tile_stage_plane(tma, &bar, s0_plane(s, 0), gl, s.y_first, rows, W, L);
    tile_stage_plane(tma, &bar, s0_plane(s, 1), gl + (size_t)L * W, s.y_first, rows, W, L);
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        const uint32_t T = (k & 1) ? a.T8[r] : a.T4[r];
        tab.tm[k >> 1][k & 1] = ((T >> (31 - (k >> 1))) & 1u) ? 0xFFFFFFFFu : 0u;
    }
    if (MEASURE)
        for (int k = threadIdx.x; k < n_live; k += blockDim.x) {
            int slot, ia, ib;
            acc_slot_decode(k, a.n_levels, slot, ia, ib);
            acc64[k] = 0ll;
            dec[k] = make_int2(slot, ia | (ib << 8));
        }
    if (threadIdx.x == 0) S_sh[X_ONE] = 1;
    const uint32_t anti = a.anti[r];
    tile_stage_wait(tma, &bar);
    __syncthreads();

    for (int smp = 0; smp < a.n_samples; ++smp) {
        const unsigned long long t = t0 + (unsigned long long)smp * a.m;
        if (MEASURE) {
            for (int k = threadIdx.x; k < (MAX_LEVELS + 1) * 4; k += blockDim.x) red[k] = 0;
            __syncthreads();
            Counts c = {0u, 0u, 0u, 0u};
            if (a.bits == 32) {  // L >= 64; the level-1 lattice goes to shared memory (word idx = i * W + w)
                if (SMALL) measure_strip_b32<1, false>(s, 1, L, lw, 0, bufA, a.seed, replica, t, c);
                else measure_strip_b32<0, false>(s, 1, L, lw, 0, bufA, a.seed, replica, t, c);
            } else {
                const int npairs = (L >> 1) << lw;
                TieCache coins;
                coins.init();
                for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
                    const int i = idx >> lw, w = idx & (W - 1);
                    uint32_t maj, tie;
                    measure_pair0(s, 1 + 2 * i, w, c, maj, tie);
                    uint32_t out = maj;
                    if (tie) out |= tie & coins.get(a.seed, (uint32_t)idx, replica, t, 1);
                    bufA[idx] = out;
                }
            }
            warp_reduce_to(c, red);
            __syncthreads();
            pyramid_in_smem(bufA, bufB, L, 1, a.n_levels, red, nullptr, nullptr, r, a.seed, replica, t);
            if (threadIdx.x <= a.n_levels) {
                const int lv = threadIdx.x;
                long long S[4];
                counts_to_S((long long)(L >> lv), red[lv * 4 + 0], red[lv * 4 + 1], red[lv * 4 + 2], red[lv * 4 + 3], S);
                for (int k = 0; k < 4; ++k) S_sh[lv * 4 + k] = S[k];
                if (lv == 0) {
                    S_sh[X_ABSM] = S[3] < 0 ? -S[3] : S[3];
                    S_sh[X_M2H] = (S[3] * S[3]) >> M4_SPLIT_BITS;
                    S_sh[X_M2L] = (S[3] * S[3]) & ((1ll << M4_SPLIT_BITS) - 1);
                }
            }
            __syncthreads();
            if (a.accumulate) {
                for (int k = threadIdx.x; k < n_live; k += blockDim.x) {
                    const int e = dec[k].y;
                    acc64[k] += S_sh[e & 255] * S_sh[e >> 8];
                }
            }
            // S_sh / red are rewritten only after the barriers inside the sweeps below (or at the loop top)
        }
        for (int h = 0; h < 2 * a.m; ++h) {
            const int c = h & 1;
            if (SMALL) mc_half_sweep_t<1>(s, c, 1, L, lw, anti, &tab, q, a.seed, replica, t + (unsigned long long)(h >> 1));
            else mc_half_sweep(s, c, 1, L, lw, anti, &tab, q, a.seed, replica, t + (unsigned long long)(h >> 1));
            uint32_t *pc = s0_plane(s, c);  // refresh this colour's periodic halo rows
            for (int w = threadIdx.x; w < 2 * W; w += blockDim.x) {
                if (w < W) pc[w] = pc[L * W + w];
                else pc[(L + 1) * W + (w - W)] = pc[W + (w - W)];
            }
            __syncthreads();
        }
    }

    if (a.m > 0) {
        unstage_rows(gl, s0_plane(s, 0) + W, 0, L, W, L);
        unstage_rows(gl + (size_t)L * W, s0_plane(s, 1) + W, 0, L, W, L);
    }
    if (MEASURE) {
        __syncthreads();
        if (threadIdx.x <= a.n_levels)
            for (int k = 0; k < 4; ++k) a.S_out[((size_t)r * (MAX_LEVELS + 1) + threadIdx.x) * 4 + k] = S_sh[threadIdx.x * 4 + k];
        if (a.accumulate) {
            const size_t base = ((size_t)r * a.n_bins + a.bin) * N_SLOTS;
            for (int k = threadIdx.x; k < n_live; k += blockDim.x) {
                const int slot = dec[k].x;
                add128(&a.acc_lo[base + slot], &a.acc_hi[base + slot], (__int128)acc64[k]);
            }
        }
    }
}

// ---- resident kernel for the smallest lattices: several replicas per warp ----------------------------------------------
// L <= 32: a replica has one word per row and colour and L/2 column walkers (two rows each per half-sweep), so a one-warp CTA
// of k_resident keeps L/2 of its 32 lanes busy (4 at L = 8, BASELINE config 1).  Here a warp takes 64/L replicas at once:
// lane = (replica within the warp, walker), every lane runs its own replica's Philox stream with its own thresholds, the
// warp collectives become segment-wide (shuffles of width L/2), and the few words whose lanes are still undecided after the
// first calls are finished in place by the lane that owns them (no queue: a warp has at most 32 words in flight).  Same
// specification, same results as k_resident — the update is the scalar form metropolis_flip_mask itself.
struct MultiLayout {
    int planes_off, bufA_off, S_off, red_off, tab_off, per_replica, dec_off, total_words, sdim;
};
MCRG_HD MultiLayout multi_layout(int L, int n_levels) {
    MultiLayout o;
    const int rpw = L >= 2 ? 64 / L : 32;                   // replicas per warp
    const int n_live = 6 + (NOP + NOP * NOP) * (n_levels + 1) + 2 * NOP * NOP * n_levels;
    o.sdim = (n_levels + 1) * 4 + 4;                        // the sums of a sample + the four pseudo-entries (acc_slot_decode)
    o.planes_off = 0;                                       // [colour][L + 2] words
    o.bufA_off = 2 * (L + 2);                               // level-1 lattice, L/2 words (at least 1)
    o.S_off = o.bufA_off + (L / 2 > 0 ? L / 2 : 1);         // int [sdim]: |S| <= 4 L^2 <= 4096 here
    o.red_off = o.S_off + o.sdim;                           // unsigned [(n_levels + 1) * 4]
    o.tab_off = (o.red_off + (n_levels + 1) * 4 + 3) & ~3;  // McTable (64 words), 16-byte aligned
    o.per_replica = o.tab_off + 64;
    o.dec_off = o.per_replica * rpw;                        // int2 [n_live], shared by the replicas of the warp
    o.total_words = o.dec_off + 2 * n_live;
    return o;
}

// accumulator slots a lane of k_resident_multi owns at most: ceil(n_live / (L/2)) over L = 2 .. 32
constexpr int MULTI_KMAX = 24;

// sum over the lanes of a segment of width G (a power of two), result in every lane of the segment
__device__ __forceinline__ uint32_t seg_sum(uint32_t v, int G) {
    for (int d = 1; d < G; d <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

template <bool MEASURE>
__global__ void __launch_bounds__(32, 16) k_resident_multi(const __grid_constant__ ResidentArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int L = a.L, G = L >= 2 ? (L / 2 > 0 ? L / 2 : 1) : 1, rpw = 32 / G;
    const int lane = threadIdx.x, gi = lane / G, li = lane - gi * G;
    const int r = blockIdx.x * rpw + gi;
    const bool live = r < a.n_replicas;
    const MultiLayout lay = multi_layout(L, a.n_levels);
    const int rows = L + 2, n_lv = a.n_levels;
    uint32_t *mine = smem + gi * lay.per_replica;
    Strip0 s;
    s.base = mine + lay.planes_off;
    s.rows = rows;
    s.W = 1;
    s.bits = a.bits;
    s.mask = valid_mask(a.bits);
    s.L = L;
    s.y_first = L - 1;
    uint32_t *bufA = mine + lay.bufA_off;
    int *S_sh = reinterpret_cast<int *>(mine + lay.S_off);
    unsigned int *red = mine + lay.red_off;
    McTable *tab = reinterpret_cast<McTable *>(mine + lay.tab_off);
    // this lane's accumulator slots li, li + G, ..: 64-bit sums in registers for the whole launch (products <= 2^24, the host
    // layer splits runs at 2^22 samples), one IMAD.WIDE per slot and sample
    long long acc[MULTI_KMAX];
#pragma unroll
    for (int j = 0; j < MULTI_KMAX; ++j) acc[j] = 0ll;
    int2 *dec = reinterpret_cast<int2 *>(smem + lay.dec_off);
    const int n_live = acc_live_slots(n_lv), s_one = (n_lv + 1) * 4;
    const uint32_t replica = a.replica_base + (uint32_t)(live ? r : 0);
    const unsigned long long t0 = *a.d_t + a.t_off;
    const uint32_t T4 = live ? a.T4[r] : 0u, T8 = live ? a.T8[r] : 0u, anti = live ? a.anti[r] : 0u;
    uint32_t *gl = a.planes + (size_t)(live ? r : 0) * 2 * L;
    for (int k = li; k < 64; k += G) tab->tm[k >> 1][k & 1] = ((((k & 1) ? T8 : T4) >> (31 - (k >> 1))) & 1u) ? 0xFFFFFFFFu : 0u;

    if (live)
        for (int idx = li; idx < 2 * rows; idx += G) {  // both planes with their periodic halo rows
            const int c = idx / rows, lr = idx - c * rows;
            s.base[idx] = gl[c * L + ((L - 1 + lr) & (L - 1))];
        }
    if (MEASURE) {
        for (int k = lane; k < n_live; k += 32) {  // slot decoding, once per launch, shared by the warp's replicas
            int slot, ia, ib;
            acc_slot_decode(k, n_lv, slot, ia, ib);
            if (ia >= X_ONE) ia = s_one + (ia - X_ONE);
            if (ib >= X_ONE) ib = s_one + (ib - X_ONE);
            dec[k] = make_int2(slot, ia | (ib << 8));
        }
        if (li == 0) S_sh[s_one] = 1;  // X_ONE
    }
    __syncwarp();

    for (int smp = 0; smp < a.n_samples; ++smp) {
        const unsigned long long t = t0 + (unsigned long long)smp * a.m;
        if (MEASURE) {
            // level 0: one row pair per lane (L >= 4; the 2 x 2 lattice has a single pair), block to level 1
            Counts c = {0u, 0u, 0u, 0u};
            if (live && 2 * li < L) {
                uint32_t maj, tie;
                measure_pair0(s, 1 + 2 * li, 0, c, maj, tie);
                if (tie) maj |= tie & tie_word(a.seed, (uint32_t)li, replica, t, 1);
                bufA[li] = maj;
            }
            const uint32_t c0 = seg_sum(c.anti_nn, G), c1 = seg_sum(c.anti_nnn, G), c2 = seg_sum(c.odd_plaq, G), c3 = seg_sum(c.up, G);
            if (li == 0) {
                red[0] = c0;
                red[1] = c1;
                red[2] = c2;
                red[3] = c3;
            }
            __syncwarp();
            // levels 1 .. n_lv: one row per lane of the segment (pyramid_tail_warp, segment-wide)
            if (n_lv >= 1) {
                int Ln = L >> 1;
                uint32_t row = (live && li < Ln) ? bufA[li] : 0u;
                uint32_t coin = 0u;
                {
                    int base = 0, my_lv = -1, my_q = 0;
                    for (int lv = 1, n = Ln >> 1; lv < n_lv; ++lv, n >>= 1) {
                        if (li >= base && li < base + n) {
                            my_lv = lv + 1;
                            my_q = li - base;
                        }
                        base += n;
                    }
                    if (my_lv >= 0) coin = tie_word(a.seed, (uint32_t)my_q, replica, t, my_lv);
                }
                int coin_base = 0;
                for (int lv = 1; lv <= n_lv; ++lv, Ln >>= 1) {
                    const int bits = Ln;
                    const uint32_t mask = valid_mask(bits);
                    const int below = (li + 1 >= Ln) ? 0 : li + 1;
                    const uint32_t r0 = row, r1 = __shfl_sync(0xFFFFFFFFu, row, below, G);
                    uint32_t q0 = 0u, q1 = 0u, q2 = 0u, q3 = 0u;
                    if (li < Ln) {
                        const uint32_t r0u = shift_up_index(r0, r0, bits, mask), r1u = shift_up_index(r1, r1, bits, mask);
                        const uint32_t r1d = shift_down_index(r1, r1, bits, mask);
                        q0 = popc32(r0 ^ r0u) + popc32(r0 ^ r1);
                        q1 = popc32(r0 ^ r1u) + popc32(r0 ^ r1d);
                        q2 = popc32(r0 ^ r0u ^ r1 ^ r1u);
                        q3 = popc32(r0);
                    }
                    q0 = seg_sum(q0, G);
                    q1 = seg_sum(q1, G);
                    q2 = seg_sum(q2, G);
                    q3 = seg_sum(q3, G);
                    if (li == 0) {
                        red[lv * 4 + 0] = q0;
                        red[lv * 4 + 1] = q1;
                        red[lv * 4 + 2] = q2;
                        red[lv * 4 + 3] = q3;
                    }
                    if (lv < n_lv) {
                        const int Lb = Ln >> 1;
                        const uint32_t cw = __shfl_sync(0xFFFFFFFFu, coin, (coin_base + (li >> 1)) & (G - 1), G);
                        coin_base += Lb;
                        uint32_t o = 0u;
                        if (li < Ln && !(li & 1)) {  // rows (li, li + 1) -> block row li / 2
                            uint32_t m, tie;
                            majority4(r0, r0 >> 1, r1, r1 >> 1, m, tie);
                            o = compress_even(m) | (compress_even(tie) & cw);
                        }
                        row = __shfl_sync(0xFFFFFFFFu, o, (2 * li) & (G - 1), G);
                        if (li >= Lb) row = 0u;
                    }
                }
            }
            __syncwarp();
            for (int lv = li; lv <= n_lv; lv += G) {
                long long S[4];
                counts_to_S((long long)(L >> lv), red[lv * 4 + 0], red[lv * 4 + 1], red[lv * 4 + 2], red[lv * 4 + 3], S);
                for (int k = 0; k < 4; ++k) S_sh[lv * 4 + k] = (int)S[k];
                if (lv == 0) {
                    S_sh[s_one + 1] = (int)(S[3] < 0 ? -S[3] : S[3]);                            // X_ABSM
                    S_sh[s_one + 2] = (int)((S[3] * S[3]) >> M4_SPLIT_BITS);                     // X_M2H
                    S_sh[s_one + 3] = (int)((S[3] * S[3]) & ((1ll << M4_SPLIT_BITS) - 1));      // X_M2L
                }
            }
            __syncwarp();
            if (a.accumulate) {
#pragma unroll
                for (int j = 0; j < MULTI_KMAX; ++j) {
                    const int k = li + j * G;
                    if (k < n_live) {
                        const int e = dec[k].y;
                        acc[j] += (long long)S_sh[e & 255] * (long long)S_sh[e >> 8];
                    }
                }
            }
            __syncwarp();
        }
        for (int h = 0; h < 2 * a.m; ++h) {
            const int c = h & 1, o = 1 - c;
            uint32_t *pc = s.base + c * rows;
            const uint32_t *po = s.base + o * rows;
            const unsigned long long sweep = t + (unsigned long long)(h >> 1);
            const uint32_t c3_base = ((uint32_t)PURPOSE_MC << 28) | (uint32_t)((sweep >> 32) & 0xFFFFFu);
            const McPhiloxHead head = mc_philox_head(a.seed, replica, (uint32_t)sweep);
            if (live)
                for (int lr = 1 + 2 * li; lr < 1 + 2 * li + 2 && lr <= L; ++lr) {  // two rows per lane (both rows of the 2 x 2 lattice)
                    const int y = lr - 1;
                    const uint32_t tw = pc[lr], n0 = po[lr];
                    const uint32_t n1 = ((y + c) & 1) ? shift_up_index(n0, n0, s.bits, s.mask) : shift_down_index(n0, n0, s.bits, s.mask);
                    const uint32_t word_id = (uint32_t)(c * L + y);
                    // the fast forms of mcfast.cuh (same decisions as metropolis_flip_mask): two calls and eight planes at once,
                    // what is left — rarely anything with at most 16 sites per word — is finished in place
                    uint32_t ge2, sel;
                    mc_neighbour_count(tw ^ po[lr - 1] ^ anti, tw ^ po[lr + 1] ^ anti, tw ^ n0 ^ anti, tw ^ n1 ^ anti, ge2, sel);
                    uint32_t eq = ~ge2 & s.mask, lt = 0u;
                    U4 r0, r1;
                    mc_philox_pair(head, a.seed, word_id, c3_base, r0, r1);
                    mc_compare4(r0, tab, 0, sel, eq, lt);
                    mc_compare4(r1, tab, 4, sel, eq, lt);
                    if (eq != 0u) lt |= mc_finish(eq, sel, 2, tab, head, a.seed, word_id, c3_base);
                    pc[lr] = tw ^ ((ge2 & s.mask) | lt);
                }
            __syncwarp();
            if (live && li == 0) {  // refresh this colour's periodic halo rows
                pc[0] = pc[L];
                pc[L + 1] = pc[1];
            }
            __syncwarp();
        }
    }

    if (live && a.m > 0)
        for (int idx = li; idx < 2 * L; idx += G) {
            const int c = idx / L, y = idx - c * L;
            gl[c * L + y] = s.base[c * rows + 1 + y];
        }
    if (MEASURE && live) {
        for (int lv = li; lv <= n_lv; lv += G)
            for (int k = 0; k < 4; ++k) a.S_out[((size_t)r * (MAX_LEVELS + 1) + lv) * 4 + k] = S_sh[lv * 4 + k];
        if (a.accumulate) {
            const size_t base = ((size_t)r * a.n_bins + a.bin) * N_SLOTS;
#pragma unroll
            for (int j = 0; j < MULTI_KMAX; ++j) {
                const int k = li + j * G;
                if (k < n_live) {
                    const int slot = dec[k].x;
                    add128(&a.acc_lo[base + slot], &a.acc_hi[base + slot], (__int128)acc[j]);
                }
            }
        }
    }
}

int g_max_smem = -1;

}  // namespace

static int pick_threads(long long work_items, int max_threads) {
    long long t = (work_items + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > max_threads) t = max_threads;
    return (int)t;
}

int sweep0_threads(int L, int R, int H) {
    const int W = l0_words(L);
    int threads = pick_threads((long long)(R + 2 * H) * W, SWEEP_THREADS);
    if (threads < W) threads = W;  // mc_half_sweep: a thread owns one column, blockDim is a multiple of W
    return threads;
}

size_t sweep0_smem_bytes(int L, int R, int H) {
    const int words = (R + 2 * H) * l0_words(L);
    const int warps = sweep0_threads(L, R, H) / 32;
    return ((((size_t)2 * words + 3) & ~(size_t)3) + (size_t)4 * warps * sweep0_queue_cap(words, warps)) * sizeof(uint32_t);
}

// resident k_sweep0 CTAs per SM for a strip geometry (registers: at most MCRG_SWEEP_MIN_BLOCKS x 256 threads; shared memory)
int sweep0_occupancy(int L, int R, int H) {
    sweep0_max_smem();
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sweep0<true>, sweep0_threads(L, R, H), sweep0_smem_bytes(L, R, H));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int sweep0_max_smem() {
    if (g_max_smem < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        // dynamic limit = opt-in maximum minus the kernels' static shared memory (k_resident: 1032 bytes)
        const int dyn = v - 2048;
        cudaError_t e1 = cudaFuncSetAttribute(k_sweep0<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        cudaError_t e2 = cudaFuncSetAttribute(k_sweep0<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        cudaError_t e3 = cudaFuncSetAttribute(k_level, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        cudaError_t e4 = cudaFuncSetAttribute(k_resident<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        cudaError_t e5 = cudaFuncSetAttribute(k_resident<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        g_max_smem = (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && e4 == cudaSuccess && e5 == cudaSuccess) ? dyn : 48 * 1024;
        (void)cudaGetLastError();  // a refused opt-in only lowers the limit we plan with
    }
    return g_max_smem;
}


// Threads of a resident CTA = column walkers of a half-sweep: W columns x (L rows / rows per walker), where a walker
// takes at least two rows when a warp spans several rows (W < 32, see mc_half_sweep_t).  More threads than that would
// idle through the sweeps and only cost registers, i.e. resident CTAs per SM.  `forced` (MCRG_RESIDENT_THREADS, read per
// context) overrides the choice for tuning and for the block-size independence test.
int resident_threads(int L, int forced) {
    const int W = l0_words(L);
    int threads = pick_threads((long long)W * ((L + 1) / 2), SWEEP_THREADS);
    if (forced >= 32 && (forced & 31) == 0 && forced <= SWEEP_THREADS) threads = forced;
    if (threads < W) threads = W;
    return threads;
}

void launch_resident(const ResidentArgs &a, int n_replicas, bool measure, int forced_threads, cudaStream_t st) {
    sweep0_max_smem();
    if (a.L <= 32 && forced_threads == 0) {  // several replicas per warp (k_resident_multi)
        const int rpw = 64 / a.L;
        const size_t smem = (size_t)multi_layout(a.L, a.n_levels).total_words * sizeof(uint32_t);
        ResidentArgs b = a;
        b.n_replicas = n_replicas;
        if (measure) k_resident_multi<true><<<(n_replicas + rpw - 1) / rpw, 32, smem, st>>>(b);
        else k_resident_multi<false><<<(n_replicas + rpw - 1) / rpw, 32, smem, st>>>(b);
        return;
    }
    const int threads = resident_threads(a.L, forced_threads);
    const size_t smem = (size_t)resident_layout(a.L, threads, a.n_levels).total_words * sizeof(uint32_t);
    if (threads == 32 && l0_words(a.L) == 1) {  // one-warp CTAs of one-word rows (L <= 64), 28 per SM
        if (measure) k_resident<true, true><<<n_replicas, threads, smem, st>>>(a);
        else k_resident<false, true><<<n_replicas, threads, smem, st>>>(a);
    } else {
        if (measure) k_resident<true, false><<<n_replicas, threads, smem, st>>>(a);
        else k_resident<false, false><<<n_replicas, threads, smem, st>>>(a);
    }
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): when the previous operation of the stream is a
// kernel, this kernel's CTAs may be scheduled — and run their prologue: shared-memory tables, mbarrier set-up — while that one is
// still draining.  Only for kernels that call pdl_wait() before their first global access.  pdl = false launches the plain way
// (per context: MCRG_PDL=0).
template <typename Args>
void launch_pdl(void (*kernel)(Args), dim3 grid, int threads, size_t smem, cudaStream_t st, bool pdl, const Args &a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

void launch_sweep0(const SweepArgs &a, int n_replicas, bool measure, cudaStream_t st, bool pdl) {
    sweep0_max_smem();
    const size_t smem = sweep0_smem_bytes(a.L, a.R, a.H);
    const dim3 grid(a.strips, n_replicas);
    const int threads = sweep0_threads(a.L, a.R, a.H);
    // Where k_sweep0 releases the next kernel (pdl_trigger), measured (profiles/r2/pdl_ab.txt): at the top of the kernel the next
    // grid's CTAs settle on whatever slots are free while this grid still runs — uneven over the SMs when the grid is less than a
    // wave (one 4096^2 replica: -8 %); just before the store, the next grid is scheduled into the slots of a finished wave and only
    // its launch latency and prologue overlap the drain: +21 % for one 4096^2 replica, +1..3 % for C3 / C4 / C5 sweep-only.
    if (measure) launch_pdl(k_sweep0<true>, grid, threads, smem, st, pdl, a);
    else launch_pdl(k_sweep0<false>, grid, threads, smem, st, pdl, a);
}

void launch_level(const LevelArgs &a, int n_replicas, cudaStream_t st, bool pdl) {
    sweep0_max_smem();
    const int Wn = nat_words(a.Ln);
    const size_t smem = (size_t)(a.R + 1) * Wn * sizeof(uint32_t);
    const dim3 grid(a.strips, n_replicas);
    launch_pdl(k_level, grid, pick_threads((long long)a.R * Wn, MCRG_LEVEL_THREADS), smem, st, pdl, a);
}

void launch_tail(const TailArgs &a, int n_replicas, cudaStream_t st, bool pdl) { launch_pdl(k_tail, dim3(n_replicas), 256, 0, st, pdl, a); }

}  // namespace mcrg
