// Small kernels around the hot path: initial states, format conversion at the boundary (the reference's int32 column-major
// imat <-> colour planes), sweep-counter bookkeeping, exact accumulator totals for the all-reduce.
#include "kernels.cuh"

namespace mcrg {

namespace {

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

// packed natural transport format (hostpack.cpp: row y, bit x, max(1, L/32) words per row) -> colour planes.
// One thread per (replica, row, colour-plane word w): the word's 32 sites x = 2x'+par come from natural words 2w, 2w+1.
__global__ void k_pack_nat(const uint32_t *nat, uint32_t *planes, int L, int W, int bits, size_t n) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int w = (int)(idx % W);
    const size_t ry = idx / W;
    const int y = (int)(ry % L);
    const size_t r = ry / L;
    const int Wn = nat_words(L);
    const uint32_t *row = nat + (r * L + y) * (size_t)Wn;
    uint32_t even, odd;
    if (Wn == 1) {
        even = compress_even(row[0]);
        odd = compress_even(row[0] >> 1);
    } else {
        const uint32_t a = row[2 * w], b = row[2 * w + 1];
        even = compress_even(a) | (compress_even(b) << 16);
        odd = compress_even(a >> 1) | (compress_even(b >> 1) << 16);
    }
    const uint32_t mask = valid_mask(bits);
    const int ce = y & 1;  // plane whose row offset is 0 holds the even-x sites
    planes[((r * 2 + ce) * L + y) * W + w] = even & mask;
    planes[((r * 2 + (1 - ce)) * L + y) * W + w] = odd & mask;
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

// totals over (replica, bin) of every slot, as four 32-bit limbs in int64 (top limb signed).  blockIdx.y takes a chunk of the
// (replica, bin) range, sums it exactly in 128 bits and adds its four limbs to `out` (zeroed by the launcher) with 64-bit
// atomics: limbs may then exceed 32 bits, which the consumers allow for anyway (they are summed over ranks next) — any
// order of the additions gives the same integers.  (One thread per slot walking all replicas took 12 ms for the 65535 replicas
// of BASELINE config 1.)
__global__ void k_total_limbs(const unsigned long long *lo, const long long *hi, int n_rb, int chunk, long long *out) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= N_SLOTS) return;
    const int k0 = blockIdx.y * chunk, k1 = min(n_rb, k0 + chunk);
    __int128 tot = 0;
    for (int k = k0; k < k1; ++k) {
        const size_t i = (size_t)k * N_SLOTS + slot;
        tot += ((__int128)hi[i] << 64) | (__int128)lo[i];
    }
    const unsigned long long tl = (unsigned long long)tot;
    const long long th = (long long)(tot >> 64);
    unsigned long long *o = reinterpret_cast<unsigned long long *>(out) + 4 * slot;
    atomicAdd(o + 0, tl & 0xFFFFFFFFull);
    atomicAdd(o + 1, tl >> 32);
    atomicAdd(o + 2, (unsigned long long)th & 0xFFFFFFFFull);
    atomicAdd(o + 3, (unsigned long long)(th >> 32));  // two's complement: adds the signed top limb
}

// Instruction-throughput probe for the roofline note of bench.py: the arithmetic core of the Metropolis row body and
// nothing else — two interleaved Philox4x32-10 calls per step, each followed by the 4-plane lazy threshold compare with
// the per-lane threshold select — on every SM with 8 warps per scheduler.  Reports Philox calls per second.
__global__ void __launch_bounds__(256) k_probe_philox(uint32_t *out, uint64_t seed, int iters) {
    __shared__ uint2 tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = make_uint2((threadIdx.x * 37u) & 1u ? ~0u : 0u, (threadIdx.x * 11u) & 2u ? ~0u : 0u);
    __syncthreads();
    uint32_t lt = 0, eq = 0xFFFFFFFFu;
    const uint32_t sel = threadIdx.x * 2654435761u, w = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        const U4 r0 = philox_keyed(seed, w, 7u, (uint64_t)i, PURPOSE_MC, 0), r1 = philox_keyed(seed, w, 7u, (uint64_t)i, PURPOSE_MC, 1);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint2 t = tab[(8 * i + e) & 31];
            const uint32_t tm = (sel & t.x) | (~sel & t.y);
            lt |= eq & ~rr[e] & tm;
            eq &= ~(rr[e] ^ tm);
        }
        eq |= r0.x;  // keep lanes alive so that the compare is never skipped
    }
    out[w] = lt ^ eq;
}

}  // namespace

int probe_philox_rate(cudaStream_t st, double *calls_per_s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, iters = 2048;
    uint32_t *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * 256 * 4) != cudaSuccess) return -1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_probe_philox<<<blocks, 256, 0, st>>>(out, 1234567ull, iters);  // warm-up
    cudaEventRecord(a, st);
    for (int k = 0; k < 3; ++k) k_probe_philox<<<blocks, 256, 0, st>>>(out, 1234567ull + k, iters);
    cudaEventRecord(b, st);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    if (e != cudaSuccess || ms <= 0.f) return -1;
    *calls_per_s = 3.0 * 2.0 * (double)blocks * 256.0 * iters / (ms * 1e-3);
    return 0;
}

void launch_total_limbs(const unsigned long long *lo, const long long *hi, int n_rb, long long *out, cudaStream_t st) {
    int n_chunks = n_rb < 512 ? n_rb : 512;
    if (n_chunks < 1) n_chunks = 1;
    const int chunk = (n_rb + n_chunks - 1) / n_chunks;
    n_chunks = (n_rb + chunk - 1) / chunk > 0 ? (n_rb + chunk - 1) / chunk : 1;
    cudaMemsetAsync(out, 0, sizeof(long long) * 4 * N_SLOTS, st);
    k_total_limbs<<<dim3((N_SLOTS + 127) / 128, n_chunks), 128, 0, st>>>(lo, hi, n_rb, chunk, out);
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

void launch_pack_nat(const uint32_t *nat, uint32_t *planes, int L, int n_replicas, cudaStream_t st) {
    const int W = l0_words(L);
    const size_t n = (size_t)n_replicas * L * W;
    k_pack_nat<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nat, planes, L, W, l0_bits(L), n);
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
