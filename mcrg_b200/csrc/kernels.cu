// sm_100a kernels of the MCRG hot path (no tensor cores: nothing here is a contraction; the work is bitwise
// integer + Philox, staged through shared memory, reduced with warp primitives and one atomic pass per CTA).
//
//   k_sweep0<MEASURE>  one strip of R rows of one replica: stage strip + halo in shared memory (128-bit loads),
//                      [MEASURE: level-0 correlator popcounts + b=2 majority block to level 1 (Philox ties)],
//                      nsw full checkerboard Metropolis sweeps with halo recomputation (counter-based RNG makes
//                      the redundant halo updates bit-identical to the owning strip's), store the strip.
//                      Replaces IsingModel::sample_new_configuration (ising.cpp:87-155, Wolff there, Metropolis
//                      here per the north_star), Lattice::calc_interactions (lattice.cpp:102-120) and the first
//                      block_spin_transformation (mcrg.cpp:314-348) of the sample loop mcrg.cpp:72-98.
//   k_level            natural-layout level n: correlator popcounts + block to level n+1, strips of rows.
//   k_tail             one CTA per replica: remaining (small) levels entirely in shared memory, then the
//                      accumulation of mcrg.cpp:86-97 (S, S(n) x S(n-1), S(n) x S(n)) into exact 128-bit sums.
#include "kernels.cuh"

namespace mcrg {

namespace {

__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }

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

// rows [row_lo, row_lo+nrows) of a [*, W] word array: global -> shared, periodic in y (L a power of two)
__device__ __forceinline__ void stage_rows(uint32_t *dst, const uint32_t *src_plane, int y_first, int nrows, int W,
                                           int L) {
    if ((W & 3) == 0) {
        const int W4 = W >> 2, n4 = nrows * W4;
        for (int idx = threadIdx.x; idx < n4; idx += blockDim.x) {
            const int lr = idx / W4, w4 = idx - lr * W4;
            const int y = (y_first + lr) & (L - 1);
            reinterpret_cast<uint4 *>(dst)[idx] = __ldg(reinterpret_cast<const uint4 *>(src_plane + (size_t)y * W) + w4);
        }
    } else {
        const int n = nrows * W;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int lr = idx / W, w = idx - lr * W;
            const int y = (y_first + lr) & (L - 1);
            dst[idx] = __ldg(src_plane + (size_t)y * W + w);
        }
    }
}

__device__ __forceinline__ void unstage_rows(uint32_t *dst_plane, const uint32_t *src, int y_first, int nrows, int W,
                                             int L) {
    if ((W & 3) == 0) {
        const int W4 = W >> 2, n4 = nrows * W4;
        for (int idx = threadIdx.x; idx < n4; idx += blockDim.x) {
            const int lr = idx / W4, w4 = idx - lr * W4;
            const int y = (y_first + lr) & (L - 1);
            reinterpret_cast<uint4 *>(dst_plane + (size_t)y * W)[w4] = reinterpret_cast<const uint4 *>(src)[idx];
        }
    } else {
        const int n = nrows * W;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int lr = idx / W, w = idx - lr * W;
            const int y = (y_first + lr) & (L - 1);
            dst_plane[(size_t)y * W + w] = src[idx];
        }
    }
}

template <bool MEASURE>
__global__ void __launch_bounds__(SWEEP_THREADS) k_sweep0(const SweepArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ unsigned int red[4];
    const int r = blockIdx.y, strip = blockIdx.x;
    const int L = a.L, W = a.W, lw = ilog2(W);
    const int rows = a.R + 2 * a.H;
    const int y0 = strip * a.R;
    Strip0 s;
    s.base = smem;
    s.rows = rows;
    s.W = W;
    s.bits = a.bits;
    s.mask = valid_mask(a.bits);
    s.L = L;
    s.y_first = (y0 - a.H) & (L - 1);
    const uint32_t *src_r = a.src + (size_t)r * 2 * L * W;
    stage_rows(s0_plane(s, 0), src_r, s.y_first, rows, W, L);
    stage_rows(s0_plane(s, 1), src_r + (size_t)L * W, s.y_first, rows, W, L);
    if (MEASURE && threadIdx.x < 4) red[threadIdx.x] = 0;
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    __syncthreads();

    if (MEASURE) {
        Counts c = {0u, 0u, 0u, 0u};
        const int npairs = (a.R >> 1) << lw;
        uint32_t *lev1 = a.level1 + (size_t)r * (L >> 1) * W;
        for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
            const int i = idx >> lw, w = idx & (W - 1);
            uint32_t maj, tie;
            measure_pair0(s, a.H + 2 * i, w, c, maj, tie);
            const uint32_t q = (uint32_t)(((y0 >> 1) + i) << lw) + (uint32_t)w;
            uint32_t out = maj;
            if (tie) out |= tie & tie_word(a.seed, q, replica, t, 1);
            lev1[q] = out;
        }
        warp_reduce_to(c, red);
        __syncthreads();
        if (threadIdx.x < 4)
            atomicAdd(&a.cnt[((size_t)r * (MAX_LEVELS + 1) + 0) * 4 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
    }

    if (a.nsw > 0) {
        McParams p;
        p.seed = a.seed;
        p.T4 = a.T4[r];
        p.T8 = a.T8[r];
        p.anti = a.anti[r];
        for (int h = 0; h < 2 * a.nsw; ++h) {
            const int c = h & 1;
            const int lr_lo = 1 + h;
            const int n = (rows - 2 - 2 * h) << lw;
            for (int idx = threadIdx.x; idx < n; idx += blockDim.x)
                update_word0(s, c, lr_lo + (idx >> lw), idx & (W - 1), p, replica, t + (unsigned long long)(h >> 1));
            __syncthreads();
        }
        uint32_t *dst_r = a.dst + (size_t)r * 2 * L * W;
        unstage_rows(dst_r, s0_plane(s, 0) + a.H * W, y0, a.R, W, L);
        unstage_rows(dst_r + (size_t)L * W, s0_plane(s, 1) + a.H * W, y0, a.R, W, L);
    }
}

__global__ void __launch_bounds__(256) k_level(const LevelArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ unsigned int red[4];
    const int r = blockIdx.y, strip = blockIdx.x;
    const int Ln = a.Ln, Wn = nat_words(Ln), lw = ilog2(Wn);
    const int y0 = strip * a.R;
    StripN s;
    s.x = smem;
    s.W = Wn;
    s.bits = nat_bits(Ln);
    s.mask = valid_mask(s.bits);
    stage_rows(smem, a.in + (size_t)r * Ln * Wn, y0, a.R + 1, Wn, Ln);
    if (threadIdx.x < 4) red[threadIdx.x] = 0;
    const unsigned long long t = *a.d_t + a.t_off;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    __syncthreads();

    Counts c = {0u, 0u, 0u, 0u};
    const int n = a.R << lw;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int lr = idx >> lw;
        measure_rowN(s, lr, lr + 1, idx & (Wn - 1), c);
    }
    if (a.out != nullptr) {
        const int Lb = Ln >> 1, Wb = nat_words(Lb), lwb = ilog2(Wb);
        uint32_t *out_r = a.out + (size_t)r * Lb * Wb;
        const int nb = (a.R >> 1) << lwb;
        for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
            const int i = idx >> lwb, wb = idx & (Wb - 1);
            uint32_t maj, tie;
            block_pairN(s, 2 * i, wb, maj, tie);
            const uint32_t q = (uint32_t)(((y0 >> 1) + i) << lwb) + (uint32_t)wb;
            uint32_t o = maj;
            if (tie) o |= tie & tie_word(a.seed, q, replica, t, a.level + 1);
            out_r[q] = o;
        }
    }
    warp_reduce_to(c, red);
    __syncthreads();
    if (threadIdx.x < 4)
        atomicAdd(&a.cnt[((size_t)r * (MAX_LEVELS + 1) + a.level) * 4 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
}

__device__ __forceinline__ void add128(unsigned long long *lo, long long *hi, __int128 v) {
    const unsigned long long vlo = (unsigned long long)v;
    const long long vhi = (long long)(v >> 64);
    const unsigned long long old = *lo;
    const unsigned long long nl = old + vlo;
    *lo = nl;
    *hi = *hi + vhi + (nl < old ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_tail(const TailArgs a) {
    __shared__ __align__(16) uint32_t bufA[TAIL_MAX_L * (TAIL_MAX_L / 32)];
    __shared__ __align__(16) uint32_t bufB[(TAIL_MAX_L / 2) * (TAIL_MAX_L / 64)];
    __shared__ unsigned int red[(MAX_LEVELS + 1) * 4];
    __shared__ long long S_sh[(MAX_LEVELS + 1) * 4];
    const int r = blockIdx.x;
    const uint32_t replica = a.replica_base + (uint32_t)r;
    const unsigned long long t = *a.d_t + a.t_off;
    for (int k = threadIdx.x; k < (MAX_LEVELS + 1) * 4; k += blockDim.x) red[k] = 0;
    if (a.start <= a.n_levels) {
        const int Ln = a.L >> a.start, Wn = nat_words(Ln);
        stage_rows(bufA, a.in + (size_t)r * Ln * Wn, 0, Ln, Wn, Ln);
    }
    __syncthreads();
    uint32_t *cur = bufA, *nxt = bufB;
    for (int lv = a.start; lv <= a.n_levels; ++lv) {
        const int Ln = a.L >> lv, Wn = nat_words(Ln), lw = ilog2(Wn);
        StripN s;
        s.x = cur;
        s.W = Wn;
        s.bits = nat_bits(Ln);
        s.mask = valid_mask(s.bits);
        Counts c = {0u, 0u, 0u, 0u};
        const int n = Ln << lw;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int lr = idx >> lw;
            measure_rowN(s, lr, (lr + 1 == Ln) ? 0 : lr + 1, idx & (Wn - 1), c);
        }
        warp_reduce_to(c, red + lv * 4);
        if (lv < a.n_levels) {
            const int Lb = Ln >> 1, Wb = nat_words(Lb), lwb = ilog2(Wb);
            uint32_t *glob = a.levels_out + a.level_off[lv + 1] + (size_t)r * Lb * Wb;
            const int nb = Lb << lwb;
            for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
                const int yb = idx >> lwb, wb = idx & (Wb - 1);
                uint32_t maj, tie;
                block_pairN(s, 2 * yb, wb, maj, tie);
                uint32_t o = maj;
                if (tie) o |= tie & tie_word(a.seed, (uint32_t)idx, replica, t, lv + 1);
                nxt[idx] = o;
                glob[idx] = o;
            }
        }
        __syncthreads();
        uint32_t *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
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
    const long long M = S_sh[3];
    for (int slot = threadIdx.x; slot < N_SLOTS; slot += blockDim.x) {
        __int128 v = 0;
        bool live = true;
        if (slot == SLOT_N) v = 1;
        else if (slot == SLOT_ABSM) v = M < 0 ? -M : M;
        else if (slot == SLOT_M2) v = (__int128)M * M;
        else if (slot < SLOT_SS) {
            const int k = slot - SLOT_S, lv = k / NOP, op = k - lv * NOP;
            live = lv <= a.n_levels;
            if (live) v = S_sh[lv * 4 + op];
        } else if (slot < SLOT_SBS) {
            const int k = slot - SLOT_SS, lv = k / (NOP * NOP), e = k - lv * NOP * NOP, b = e / NOP, al = e - b * NOP;
            live = lv <= a.n_levels;
            if (live) v = (__int128)S_sh[lv * 4 + al] * S_sh[lv * 4 + b];
        } else {
            const int k = slot - SLOT_SBS, n1 = k / (NOP * NOP), e = k - n1 * NOP * NOP, b = e / NOP, al = e - b * NOP;
            const int n = n1 + 1;
            live = n <= a.n_levels;
            if (live) v = (__int128)S_sh[n * 4 + al] * S_sh[(n - 1) * 4 + b];  // flatten: index b*NOP+a holds Sb_a*S_b
        }
        if (live) add128(&a.acc_lo[base + slot], &a.acc_hi[base + slot], v);
    }
    if (threadIdx.x == 0) {
        const double m = (double)M;
        a.acc_d[((size_t)r * a.n_bins + a.bin) * N_DSLOTS + 0] += m * m * m * m;
    }
}

__global__ void k_advance_t(unsigned long long *d_t, unsigned long long by) { *d_t += by; }

__global__ void k_init_hot(uint32_t *planes, int L, int W, int bits, size_t n_words, uint64_t seed, uint32_t replica_base) {
    const size_t per = (size_t)2 * L * W;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n_words; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(idx / per);
        const uint32_t word_id = (uint32_t)(idx - (size_t)r * per);
        planes[idx] = philox_keyed(seed, word_id, replica_base + r, 0ull, PURPOSE_INIT, 0).x & valid_mask(bits);
    }
}

__global__ void k_fill(uint32_t *p, size_t n, uint32_t v) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) p[idx] = v;
}

// int32 column-major (an internal row is contiguous) -> colour planes.  One warp per (replica, row, word):
// lane l owns x = 64w+2l (even) and x+1 (odd); two ballots give the two colour words of that row.
__global__ void k_pack0(const int32_t *spins, uint32_t *planes, int L, int W, size_t n_warps) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_warps) return;
    const int w = (int)(warp % W);
    const size_t ry = warp / W;
    const int y = (int)(ry % L);
    const size_t r = ry / L;
    const int x = 64 * w + 2 * lane;
    int2 v = make_int2(0, 0);
    if (x < L) v = *reinterpret_cast<const int2 *>(spins + (r * L + y) * (size_t)L + x);
    const uint32_t even = __ballot_sync(0xFFFFFFFFu, v.x > 0), odd = __ballot_sync(0xFFFFFFFFu, v.y > 0);
    if (lane == 0) {
        const int ce = y & 1;  // plane whose row offset is 0 holds the even-x sites
        planes[((r * 2 + ce) * L + y) * W + w] = even;
        planes[((r * 2 + (1 - ce)) * L + y) * W + w] = odd;
    }
}

__global__ void k_unpack0(const uint32_t *planes, int32_t *spins, int L, int W, size_t n_warps) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_warps) return;
    const int w = (int)(warp % W);
    const size_t ry = warp / W;
    const int y = (int)(ry % L);
    const size_t r = ry / L;
    const int ce = y & 1;
    const uint32_t even = planes[((r * 2 + ce) * L + y) * W + w], odd = planes[((r * 2 + (1 - ce)) * L + y) * W + w];
    const int x = 64 * w + 2 * lane;
    if (x < L) {
        int2 v;
        v.x = ((even >> lane) & 1u) ? 1 : -1;
        v.y = ((odd >> lane) & 1u) ? 1 : -1;
        *reinterpret_cast<int2 *>(spins + (r * L + y) * (size_t)L + x) = v;
    }
}

__global__ void k_unpackN(const uint32_t *lev, int32_t *spins, int Ln, int Wn, size_t n_warps) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_warps) return;
    const uint32_t word = lev[warp];
    const int w = (int)(warp % Wn);
    const size_t ry = warp / Wn;
    const int x = 32 * w + lane;
    if (x < Ln) spins[ry * (size_t)Ln + x] = ((word >> lane) & 1u) ? 1 : -1;
}

// totals over (replica, bin) of every slot, as four 32-bit limbs in int64 (top limb signed)
__global__ void k_total_limbs(const unsigned long long *lo, const long long *hi, int n_rb, long long *out) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= N_SLOTS) return;
    __int128 tot = 0;
    for (int k = 0; k < n_rb; ++k) {
        const size_t i = (size_t)k * N_SLOTS + slot;
        tot += ((__int128)hi[i] << 64) | (__int128)lo[i];
    }
    const unsigned long long tl = (unsigned long long)tot;
    const long long th = (long long)(tot >> 64);
    out[4 * slot + 0] = (long long)(tl & 0xFFFFFFFFull);
    out[4 * slot + 1] = (long long)(tl >> 32);
    out[4 * slot + 2] = (long long)((unsigned long long)th & 0xFFFFFFFFull);
    out[4 * slot + 3] = th >> 32;
}

int g_max_smem = -1;

}  // namespace

size_t sweep0_smem_bytes(int L, int R, int H) { return (size_t)2 * (R + 2 * H) * l0_words(L) * sizeof(uint32_t); }

int sweep0_max_smem() {
    if (g_max_smem < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        g_max_smem = v;
        cudaFuncSetAttribute(k_sweep0<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
        cudaFuncSetAttribute(k_sweep0<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
        cudaFuncSetAttribute(k_level, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    }
    return g_max_smem;
}

static int pick_threads(long long work_items, int max_threads) {
    long long t = (work_items + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > max_threads) t = max_threads;
    return (int)t;
}

void launch_sweep0(const SweepArgs &a, int n_replicas, bool measure, cudaStream_t st) {
    sweep0_max_smem();
    const size_t smem = sweep0_smem_bytes(a.L, a.R, a.H);
    const dim3 grid(a.strips, n_replicas);
    const int threads = pick_threads((long long)(a.R + 2 * a.H) * a.W, SWEEP_THREADS);
    if (measure) k_sweep0<true><<<grid, threads, smem, st>>>(a);
    else k_sweep0<false><<<grid, threads, smem, st>>>(a);
}

void launch_level(const LevelArgs &a, int n_replicas, cudaStream_t st) {
    sweep0_max_smem();
    const int Wn = nat_words(a.Ln);
    const size_t smem = (size_t)(a.R + 1) * Wn * sizeof(uint32_t);
    const dim3 grid(a.strips, n_replicas);
    k_level<<<grid, pick_threads((long long)a.R * Wn, 256), smem, st>>>(a);
}

void launch_tail(const TailArgs &a, int n_replicas, cudaStream_t st) { k_tail<<<n_replicas, 256, 0, st>>>(a); }

void launch_total_limbs(const unsigned long long *lo, const long long *hi, int n_rb, long long *out, cudaStream_t st) {
    k_total_limbs<<<(N_SLOTS + 127) / 128, 128, 0, st>>>(lo, hi, n_rb, out);
}

void launch_advance_t(unsigned long long *d_t, unsigned long long by, cudaStream_t st) { k_advance_t<<<1, 1, 0, st>>>(d_t, by); }

void launch_init_hot(uint32_t *planes, int L, int n_replicas, uint64_t seed, uint32_t replica_base, cudaStream_t st) {
    const int W = l0_words(L);
    const size_t n = (size_t)n_replicas * 2 * L * W;
    const int blocks = (int)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
    k_init_hot<<<blocks, 256, 0, st>>>(planes, L, W, l0_bits(L), n, seed, replica_base);
}

void launch_init_cold(uint32_t *planes, int L, int n_replicas, cudaStream_t st) {
    const size_t n = (size_t)n_replicas * 2 * L * l0_words(L);
    const int blocks = (int)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
    k_fill<<<blocks, 256, 0, st>>>(planes, n, valid_mask(l0_bits(L)));
}

void launch_pack0(const int32_t *spins, uint32_t *planes, int L, int n_replicas, cudaStream_t st) {
    const int W = l0_words(L);
    const size_t n_warps = (size_t)n_replicas * L * W;
    k_pack0<<<(unsigned)((n_warps + 7) / 8), 256, 0, st>>>(spins, planes, L, W, n_warps);
}

void launch_unpack0(const uint32_t *planes, int32_t *spins, int L, int n_replicas, cudaStream_t st) {
    const int W = l0_words(L);
    const size_t n_warps = (size_t)n_replicas * L * W;
    k_unpack0<<<(unsigned)((n_warps + 7) / 8), 256, 0, st>>>(planes, spins, L, W, n_warps);
}

void launch_unpackN(const uint32_t *lev, int32_t *spins, int Ln, int n_replicas, cudaStream_t st) {
    const int Wn = nat_words(Ln);
    const size_t n_warps = (size_t)n_replicas * Ln * Wn;
    k_unpackN<<<(unsigned)((n_warps + 7) / 8), 256, 0, st>>>(lev, spins, Ln, Wn, n_warps);
}

}  // namespace mcrg
