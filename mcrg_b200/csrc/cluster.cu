// Swendsen-Wang cluster update on the bit-packed lattices (stands in for the reference's Wolff update,
// ising.cpp:87-155).  Scalar specification: orc_swendsen_wang in oracle/mcrg_oracle.c.
#include "kernels.cuh"

namespace mcrg {

namespace {

// Three kernels per update.  k_sw_tile: draw the +x and +y bonds of every site (one Philox call per site) and label the
// clusters inside 64 x 64 tiles with a union-find in shared memory.  k_sw_border: merge across tile edges in global
// memory.  Both use a lock-free union-find whose hooks always point from the larger to the smaller root (atomicMin), so
// the final root of a cluster is its smallest site index whatever the interleaving — the result is bit-identical to the
// scalar specification.  k_sw_border_compress points every tile-local root at its final root, k_sw_coins draws the cluster
// coins (128 per Philox call) and k_sw_flip flips the sites whose root's coin is set (ballots + one atomic XOR per word).
__device__ __forceinline__ int sw_find(const int *parent, int x) {
    int p = __ldcg(parent + x);
    while (p != x) {
        x = p;
        p = __ldcg(parent + x);
    }
    return x;
}

// find with path halving.  parent[x] <= x always (hooks and shortcuts only ever point to a smaller index of the same
// cluster-to-be), so a racing plain store of an ancestor can neither create a cycle nor disconnect anything for good:
// whoever replaced parent[x] by a hook keeps uniting the previous parent with the hook target (sw_unite below).
__device__ __forceinline__ int sw_find_halving(int *parent, int x) {
    int p = __ldcg(parent + x);
    while (p != x) {
        const int g = __ldcg(parent + p);
        if (g != p) __stcg(parent + x, g);
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ void sw_unite(int *parent, int a, int b) {
    while (true) {
        a = sw_find_halving(parent, a);
        b = sw_find_halving(parent, b);
        if (a == b) return;
        if (a < b) {
            const int tmp = a;
            a = b;
            b = tmp;
        }
        const int old = atomicMin(parent + a, b);  // hook root a under the smaller root b
        if (old == a) return;                      // a was still a root: done
        a = old;                                   // somebody hooked a first: continue from where it points now
    }
}

// spin of site (x, y) as a bit, from the colour planes of one replica in global memory
__device__ __forceinline__ uint32_t sw_spin_bit(const uint32_t *pl, int L, int W, int x, int y) {
    const int c = (x + y) & 1, xh = x >> 1;
    return (pl[((size_t)c * L + y) * W + (xh >> 5)] >> (xh & 31)) & 1u;
}

// Tile phase: one CTA labels the clusters INSIDE one TW x TH tile (TW = TH = min(64, L)) with a union-find in shared
// memory (shared atomics), then writes for every site the global index of its tile-local root: parent[] becomes a
// forest of depth 1 whose roots are the smallest site index of every tile-local cluster.  The bonds that leave the tile
// through its right and bottom edges (including the periodic wrap) are left to k_sw_border.  Every bond decision is the
// same Philox draw as in orc_swendsen_wang: element 0 (+x) / 1 (+y) of the call keyed by the site index.
constexpr int SW_TILE = 64;

__device__ __forceinline__ int sw_find_smem(int *lab, int x) {
    int p = lab[x];
    while (p != x) {
        const int g = lab[p];
        if (g != p) lab[x] = g;  // path halving; labels only ever decrease, a stale store is still an ancestor
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ void sw_unite_smem(int *lab, int a, int b) {
    while (true) {
        a = sw_find_smem(lab, a);
        b = sw_find_smem(lab, b);
        if (a == b) return;
        if (a < b) {
            const int tmp = a;
            a = b;
            b = tmp;
        }
        const int old = atomicMin(lab + a, b);
        if (old == a) return;
        a = old;
    }
}

// insert a zero between the bits of a 32-bit word: bit k -> bit 2k
__device__ __forceinline__ unsigned long long spread_bits(uint32_t v) {
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// T = 64 (L >= 64).  Rows of the tile are 64-bit masks in shared memory: spins S, active horizontal bonds HB (bit lx =
// bond between lx and lx+1) and vertical bonds VB (bit lx = bond between rows ly and ly+1).  Horizontal runs need no
// union-find at all: the first site of the run through (lx, ly) follows from HB[ly] with a count-leading-zeros.  The
// union-find then only merges RUNS through vertical bonds, and only through the first bond of every stretch of bonds that
// joins the same two runs.  Everything else as in the generic kernel below.
__global__ void __launch_bounds__(256) k_sw_tile64(const SwArgs a, int tiles_x, int tiles_per_replica, int n_sites) {
    constexpr int T = SW_TILE;
    __shared__ int lab[T * T];
    __shared__ unsigned long long S[T + 1], HB[T], VB[T];
    __shared__ unsigned short pairs[T * T];
    __shared__ int n_pairs;
    const int L = a.L, W = a.W;
    const int r = blockIdx.x / tiles_per_replica, tile = blockIdx.x - r * tiles_per_replica;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = tx * T, y0 = ty * T;
    const uint32_t *pl = a.planes + (size_t)r * 2 * L * W;
    const int w0 = x0 >> 6;
    if (threadIdx.x == 0) n_pairs = 0;
    if (threadIdx.x <= T) {  // natural-order row: the plane with (y + c) even holds the even x
        const int y = (y0 + threadIdx.x) & (L - 1), ce = y & 1;
        const uint32_t ev = pl[((size_t)ce * L + y) * W + w0], od = pl[((size_t)(1 - ce) * L + y) * W + w0];
        S[threadIdx.x] = spread_bits(ev) | (spread_bits(od) << 1);
    }
    __syncthreads();
    const uint32_t TP = a.TP[r];
    const unsigned long long anti = a.anti[r] ? ~0ull : 0ull;
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    const int lx = threadIdx.x & (T - 1), half = (threadIdx.x >> 5) & 1;
    for (int k = 0; k < T / 4; ++k) {
        const int ly = 4 * k + (threadIdx.x >> 6);
        const unsigned long long row = S[ly];
        const unsigned long long sat_r = ~(row ^ (row >> 1) ^ anti) & 0x7FFFFFFFFFFFFFFFull;  // bit lx: (lx, lx+1), lx < 63
        const unsigned long long sat_d = ly + 1 < T ? ~(row ^ S[ly + 1] ^ anti) : 0ull;
        const bool sr = (sat_r >> lx) & 1ull, sd = (sat_d >> lx) & 1ull;
        bool br = false, bd = false;
        if (sr || sd) {
            const int i = (y0 + ly) * L + (x0 + lx);
            const U4 u = philox_keyed(a.seed, (uint32_t)i, replica, t, PURPOSE_SW_BOND, 0);
            br = sr && u.x < TP;
            bd = sd && u.y < TP;
        }
        const uint32_t mr = __ballot_sync(0xFFFFFFFFu, br), md = __ballot_sync(0xFFFFFFFFu, bd);
        if ((threadIdx.x & 31) == 0) {
            reinterpret_cast<uint32_t *>(&HB[ly])[half] = mr;
            reinterpret_cast<uint32_t *>(&VB[ly])[half] = md;
        }
    }
    __syncthreads();
    for (int k = 0; k < T / 4; ++k) {  // run labels: first site of the horizontal run
        const int ly = 4 * k + (threadIdx.x >> 6);
        const unsigned long long z = ~HB[ly] & ((1ull << lx) - 1ull);  // broken links below lx
        const int rs = z ? 64 - __clzll((long long)z) : 0;
        lab[ly * T + lx] = ly * T + rs;
    }
    __syncthreads();
    // the vertical bonds that still have to be merged are sparse (about one site in five): gather them into a list
    // (warp-aggregated append) and run the union-find densely over the list instead of with mostly idle warps
    for (int k = 0; k < T / 4; ++k) {
        const int ly = 4 * k + (threadIdx.x >> 6);
        bool need = false;
        if (ly + 1 < T) {
            const unsigned long long vb = VB[ly];
            // the bond one column to the left joins the same two runs if both runs continue and it is active too
            need = ((vb >> lx) & 1ull) && !(lx > 0 && (((vb & HB[ly] & HB[ly + 1]) >> (lx - 1)) & 1ull));
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
        if (m) {
            int base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(&n_pairs, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (need) pairs[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = (unsigned short)(ly * T + lx);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_pairs; e += blockDim.x) {
        const int up = pairs[e];
        sw_unite_smem(lab, lab[up], lab[up + T]);
    }
    __syncthreads();
    // flatten by pointer jumping (every site in lock step, no divergent pointer chasing): lab[s] <- lab[lab[s]] until
    // nothing changes; labels only move towards the root, so concurrent reads of half-updated entries are harmless
    bool changed;
    do {
        changed = false;
        for (int k = 0; k < T / 4; ++k) {
            const int s = (4 * k + (threadIdx.x >> 6)) * T + lx;
            const int p = lab[s], g = lab[p];
            if (g != p) {
                lab[s] = g;
                changed = true;
            }
        }
    } while (__syncthreads_or(changed));
    int *parent = a.parent + (size_t)r * n_sites;
    for (int k = 0; k < T / 4; ++k) {
        const int ly = 4 * k + (threadIdx.x >> 6);
        const int root = lab[ly * T + lx];
        parent[(y0 + ly) * L + (x0 + lx)] = (y0 + (root >> 6)) * L + (x0 + (root & (T - 1)));
    }
}

// Generic tile kernel (T = L < 64: the tile is the whole lattice): one union-find entry per site, bonds merged one by one.
__global__ void __launch_bounds__(256) k_sw_tile(const SwArgs a, int T, int tiles_x, int tiles_per_replica, int n_sites) {
    __shared__ int lab[SW_TILE * SW_TILE];
    __shared__ uint32_t stage[SW_TILE + 1][2][2];  // [local row][colour][own word, word of the column right of the tile]
    const int L = a.L, W = a.W;
    const int r = blockIdx.x / tiles_per_replica, tile = blockIdx.x - r * tiles_per_replica;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = tx * T, y0 = ty * T;
    const uint32_t *pl = a.planes + (size_t)r * 2 * L * W;
    const int w0 = x0 >> 6, w1 = ((x0 + T) & (L - 1)) >> 6;
    for (int k = threadIdx.x; k < (T + 1) * 4; k += blockDim.x) {
        const int ly = k >> 2, c = (k >> 1) & 1, which = k & 1;
        const int y = (y0 + ly) & (L - 1);
        stage[ly][c][which] = pl[((size_t)c * L + y) * W + (which ? w1 : w0)];
    }
    const int n_tile = T * T, lt = ilog2(T);
    for (int s = threadIdx.x; s < n_tile; s += blockDim.x) lab[s] = s;
    __syncthreads();
    const uint32_t anti = a.anti[r] & 1u, TP = a.TP[r];
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    auto spin = [&](int lx, int ly) -> uint32_t {  // lx in [0, T], ly in [0, T]
        const int x = (x0 + lx) & (L - 1), y = y0 + ly;  // parity of y0 + ly equals the parity of the wrapped row
        const int c = (x + y) & 1, xh = x >> 1;
        return (stage[ly][c][lx == T ? 1 : 0] >> (xh & 31)) & 1u;
    };
    for (int s = threadIdx.x; s < n_tile; s += blockDim.x) {
        const int ly = s >> lt, lx = s & (T - 1);
        const uint32_t me = spin(lx, ly);
        const bool sat_r = lx + 1 < T && ((me ^ spin(lx + 1, ly) ^ anti) == 0u);
        const bool sat_d = ly + 1 < T && ((me ^ spin(lx, ly + 1) ^ anti) == 0u);
        if (!sat_r && !sat_d) continue;
        const int i = (y0 + ly) * L + (x0 + lx);
        const U4 u = philox_keyed(a.seed, (uint32_t)i, replica, t, PURPOSE_SW_BOND, 0);
        if (sat_r && u.x < TP) sw_unite_smem(lab, s, s + 1);
        if (sat_d && u.y < TP) sw_unite_smem(lab, s, s + T);
    }
    __syncthreads();
    int *parent = a.parent + (size_t)r * n_sites;
    for (int s = threadIdx.x; s < n_tile; s += blockDim.x) {
        const int root = sw_find_smem(lab, s);
        const int ly = s >> lt, lx = s & (T - 1), ry = root >> lt, rx = root & (T - 1);
        parent[(y0 + ly) * L + (x0 + lx)] = (y0 + ry) * L + (x0 + rx);
    }
}

// Border phase: the +x bonds of every tile's last column and the +y bonds of every tile's last row (periodic), merged in
// global memory with the lock-free union-find above.  One thread per border site and direction.
__global__ void __launch_bounds__(256) k_sw_border(const SwArgs a, int T, size_t n_threads, int n_sites) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n_threads) return;
    const int L = a.L, W = a.W, ll = ilog2(L), lnt = ll - ilog2(T);  // L / T tiles per direction, all powers of two
    const size_t r = idx >> (ll + lnt + 1);
    uint32_t k = (uint32_t)(idx & (((size_t)1 << (ll + lnt + 1)) - 1));
    const int dir = (int)(k >> (ll + lnt));  // 0: +x bond of a last-column site, 1: +y bond of a last-row site
    k &= (1u << (ll + lnt)) - 1u;
    const int along = (int)(k & (uint32_t)(L - 1)), blk = (int)(k >> ll);
    const int x = dir ? along : blk * T + T - 1, y = dir ? blk * T + T - 1 : along;
    const int xn = dir ? x : (x + 1) & (L - 1), yn = dir ? (y + 1) & (L - 1) : y;
    const uint32_t *pl = a.planes + r * 2 * (size_t)L * W;
    if ((sw_spin_bit(pl, L, W, x, y) ^ sw_spin_bit(pl, L, W, xn, yn) ^ (a.anti[r] & 1u)) != 0u) return;
    const int i = y * L + x;
    const U4 u = philox_keyed(a.seed, (uint32_t)i, a.replica_base + (uint32_t)r, *a.d_t + a.t_off, PURPOSE_SW_BOND, 0);
    if ((dir ? u.y : u.x) < a.TP[r]) {
        // unite the tile-local roots, not the sites: the entries of the sites themselves then never lie on a search
        // path, stay equal to their tile-local root, and k_sw_border_compress / k_sw_flip can rely on that
        int *parent = a.parent + r * (size_t)n_sites;
        sw_unite(parent, __ldcg(parent + i), __ldcg(parent + yn * L + xn));
    }
}

// After the border merges a tile-local root may sit at the bottom of a long chain of hooks (one per tile the cluster
// crosses).  Every local root that can have been hooked owns an end point of a border bond, so one thread per border
// bond points the local roots of both end points straight at their final root; k_sw_flip then needs two hops per site.
__global__ void __launch_bounds__(256) k_sw_border_compress(const SwArgs a, int T, size_t n_threads, int n_sites) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n_threads) return;
    const int L = a.L, ll = ilog2(L), lnt = ll - ilog2(T);
    const size_t r = idx >> (ll + lnt + 1);
    uint32_t k = (uint32_t)(idx & (((size_t)1 << (ll + lnt + 1)) - 1));
    const int dir = (int)(k >> (ll + lnt));
    k &= (1u << (ll + lnt)) - 1u;
    const int along = (int)(k & (uint32_t)(L - 1)), blk = (int)(k >> ll);
    const int x = dir ? along : blk * T + T - 1, y = dir ? blk * T + T - 1 : along;
    const int xn = dir ? x : (x + 1) & (L - 1), yn = dir ? (y + 1) & (L - 1) : y;
    int *parent = a.parent + r * (size_t)n_sites;
    const int ends[2] = {y * L + x, yn * L + xn};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        // parent[end point] is its tile-local root — or, when the end point IS a tile-local root that was hooked, already
        // one of its ancestors: point both the end point and that node at the final root
        const int loc = __ldcg(parent + ends[e]);
        const int root = sw_find(parent, loc);
        if (root != loc) {
            __stcg(parent + loc, root);
            __stcg(parent + ends[e], root);
        }
    }
}

// coin bitmap: bit i of a replica's map = coin of the cluster whose root is site i (used only where i is a root).
// One Philox call yields the coins of 128 consecutive site indices; one thread per call.
__global__ void __launch_bounds__(256) k_sw_coins(const SwArgs a, size_t n_calls, int calls_per_replica) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n_calls) return;
    const size_t r = idx / (size_t)calls_per_replica;
    const uint32_t g = (uint32_t)(idx - r * (size_t)calls_per_replica);
    const U4 u = philox_keyed(a.seed, g, a.replica_base + (uint32_t)r, *a.d_t + a.t_off, PURPOSE_SW_FLIP, 0);
    reinterpret_cast<uint4 *>(a.coins)[idx] = make_uint4(u.x, u.y, u.z, u.w);
}

// Flip pass.  One warp per 64 consecutive sites of a row (both colours): lane l owns x = 64w + 2l and x + 1, so the
// parent[] entries are read as coalesced 8-byte pairs and every lane has two independent dependent-load chains in
// flight.  After k_sw_border_compress every site is at most two hops from its root — site -> tile-local root -> root —
// hence root = parent[parent[i]] without a search loop.  The coins of the 2 x 32 sites are gathered with two ballots
// and applied to the two colour words with one fire-and-forget atomic XOR each (each word has exactly one owner).
__global__ void __launch_bounds__(256) k_sw_flip(const SwArgs a, size_t n_warps, int n_sites, int calls_per_replica) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_warps) return;  // whole warps leave together
    const int lane = threadIdx.x & 31;
    const int L = a.L, W = a.W, lw = ilog2(W), ll = ilog2(L);
    const int w = (int)(warp & (size_t)(W - 1));
    const size_t ry = warp >> lw;
    const int y = (int)(ry & (size_t)(L - 1));
    const size_t r = ry >> ll;
    const int x = 64 * w + 2 * lane;
    bool f0 = false, f1 = false;
    if (x < L) {
        const int *parent = a.parent + r * (size_t)n_sites;
        const uint32_t *coins = a.coins + r * (size_t)calls_per_replica * 4;
        const int2 p = __ldcs(reinterpret_cast<const int2 *>(parent + y * L + x));  // read once: streaming
        const uint32_t r0 = (uint32_t)__ldg(parent + p.x), r1 = (uint32_t)__ldg(parent + p.y);
        f0 = (__ldg(coins + (r0 >> 5)) >> (r0 & 31u)) & 1u;
        f1 = (__ldg(coins + (r1 >> 5)) >> (r1 & 31u)) & 1u;
    }
    const uint32_t even = __ballot_sync(0xFFFFFFFFu, f0), odd = __ballot_sync(0xFFFFFFFFu, f1);
    if (lane == 0) {
        const int ce = y & 1;  // the plane whose row offset is 0 holds the even-x sites
        uint32_t *pl = a.planes + r * 2 * (size_t)L * W;
        if (even) atomicXor(pl + ((size_t)ce * L + y) * W + w, even);
        if (odd) atomicXor(pl + ((size_t)(1 - ce) * L + y) * W + w, odd);
    }
}

}  // namespace

void launch_sw_update(const SwArgs &a, int n_replicas, cudaStream_t st) {
    const size_t n_sites = (size_t)a.L * a.L;
    const int T = a.L < SW_TILE ? a.L : SW_TILE, tiles_x = a.L / T, tiles = tiles_x * tiles_x;
    if (T == SW_TILE) k_sw_tile64<<<(unsigned)(n_replicas * tiles), 256, 0, st>>>(a, tiles_x, tiles, (int)n_sites);
    else k_sw_tile<<<(unsigned)(n_replicas * tiles), 256, 0, st>>>(a, T, tiles_x, tiles, (int)n_sites);
    const size_t n_border = (size_t)n_replicas * 2 * a.L * tiles_x;
    k_sw_border<<<(unsigned)((n_border + 255) / 256), 256, 0, st>>>(a, T, n_border, (int)n_sites);
    k_sw_border_compress<<<(unsigned)((n_border + 255) / 256), 256, 0, st>>>(a, T, n_border, (int)n_sites);
    const int calls_per_replica = (int)((n_sites + 127) / 128);
    const size_t n_calls = (size_t)n_replicas * calls_per_replica;
    k_sw_coins<<<(unsigned)((n_calls + 255) / 256), 256, 0, st>>>(a, n_calls, calls_per_replica);
    const size_t n_warps = (size_t)n_replicas * a.L * a.W;  // one warp per 64 sites of a row
    k_sw_flip<<<(unsigned)((n_warps + 7) / 8), 256, 0, st>>>(a, n_warps, (int)n_sites, calls_per_replica);
}

}  // namespace mcrg
