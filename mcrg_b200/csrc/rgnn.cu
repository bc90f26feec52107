// RGNN forward and finite-difference gradient on the device (rgnn.cpp:281-339).
#include "kernels.cuh"

namespace mcrg {

namespace {

// ---- RGNN (rgnn.cpp:281-339): scalar output of the b=2 filter pyramid and its central-difference gradient ---------
// One thread per (replica, variant): variant 0 = W, variants 1..8 = W +- h on one weight (rgnn.cpp:321-331).  The
// pyramid is walked depth-first in Morton order with a log2(L)-deep stack, in the oracle's operation order and with
// explicit round-to-nearest multiplies/adds (no FMA contraction), so results equal the scalar code bit for bit.
// W is column-major: W[k*2 + r] = W(r,k).  Internal coordinates: x = reference row i, y = reference column j.
__device__ __forceinline__ double rgnn_block(const double W[4], double b00, double b10, double b01, double b11) {
    // block(k,c): k = row offset (x), c = column offset (y); conv(r,c) = W(r,0) block(0,c) + W(r,1) block(1,c)
    double l1 = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double x0 = c == 0 ? b00 : b01, x1 = c == 0 ? b10 : b11;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const double acc = __dadd_rn(__dmul_rn(W[0 * 2 + r], x0), __dmul_rn(W[1 * 2 + r], x1));
            l1 = __dadd_rn(l1, fabs(acc));
        }
    }
    return l1;
}

__device__ __forceinline__ double spin_at(const uint32_t *planes, int L, int W, int x, int y) {
    const int c = (x + y) & 1, xh = x >> 1;
    return ((planes[((size_t)c * L + y) * W + (xh >> 5)] >> (xh & 31)) & 1u) ? 1.0 : -1.0;
}

__global__ void __launch_bounds__(128) k_rgnn(const uint32_t *planes_all, int L, int n_replicas, const double *W_in, double h,
                                              double *u_out, double *grad_out, double *acc, int accumulate) {
    const int slot = threadIdx.x / 9, v = threadIdx.x - slot * 9;
    const int r = threadIdx.x < 126 ? blockIdx.x * 14 + slot : n_replicas;  // threads 126, 127 idle
    __shared__ double sh_u[128];
    double u = 0.0;
    if (r < n_replicas) {
        double W[4] = {W_in[0], W_in[1], W_in[2], W_in[3]};
        if (v > 0) {  // element e = (v-1)/2 in the order i outer, j inner of rgnn.cpp:318-319: (i,j) -> W[j*2+i]
            const int e = (v - 1) >> 1, i = e >> 1, j = e & 1;
            W[j * 2 + i] = (v & 1) ? __dadd_rn(W[j * 2 + i], h) : __dadd_rn(W[j * 2 + i], -h);
        }
        const uint32_t *planes = planes_all + (size_t)r * 2 * L * l0_words(L);
        const int Wd = l0_words(L);
        int depth = 0;
        for (int n = L; n > 1; n >>= 1) ++depth;  // number of filter applications
        // stack[level][slot]: results of the four children of the node being assembled at `level`
        double stack[MAX_LEVELS][4];
        int cnt[MAX_LEVELS];
        for (int l = 0; l < depth; ++l) cnt[l] = 0;
        const int n_blocks = (L / 2) * (L / 2);
        for (int m = 0; m < n_blocks; ++m) {
            // Morton decode: block (bi, bj) at the first level, bi from the even bits, bj from the odd bits
            int bi = 0, bj = 0;
            for (int k = 0; k < depth - 1; ++k) {
                bi |= ((m >> (2 * k)) & 1) << k;
                bj |= ((m >> (2 * k + 1)) & 1) << k;
            }
            double val = rgnn_block(W, spin_at(planes, L, Wd, 2 * bi, 2 * bj), spin_at(planes, L, Wd, 2 * bi + 1, 2 * bj),
                                    spin_at(planes, L, Wd, 2 * bi, 2 * bj + 1), spin_at(planes, L, Wd, 2 * bi + 1, 2 * bj + 1));
            // push upwards: child slot = (k offset, c offset) = (bi & 1, bj & 1) at each level
            int lvl = 1, ci = bi, cj = bj;
            while (lvl < depth) {
                stack[lvl][(cj & 1) * 2 + (ci & 1)] = val;
                if (++cnt[lvl] < 4) break;
                cnt[lvl] = 0;
                val = rgnn_block(W, stack[lvl][0], stack[lvl][1], stack[lvl][2], stack[lvl][3]);
                ci >>= 1;
                cj >>= 1;
                ++lvl;
            }
            if (lvl == depth) u = val;
        }
    }
    sh_u[threadIdx.x] = u;
    __syncthreads();
    // variant 0 of each replica gathers its 8 neighbours (same block: 128 is not a multiple of 9, so guard the edge
    // by recomputing nothing — blocks are sized in whole replicas by the launcher: 126 = 14 * 9 threads are used)
    if (r < n_replicas && v == 0) {
        const double *uu = &sh_u[threadIdx.x];
        const double inv2h = 2.0 * h;
        double g[4];
        for (int e = 0; e < 4; ++e) g[e] = __ddiv_rn(__dadd_rn(uu[1 + 2 * e], -uu[2 + 2 * e]), inv2h);  // (out1-out2)/(2h)
        // g[e] with e = i*2 + j  ->  column-major grad[j*2 + i]
        if (u_out) u_out[r] = uu[0];
        if (grad_out)
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j) grad_out[(size_t)r * 4 + j * 2 + i] = g[i * 2 + j];
        if (accumulate) {
            double *a = acc + (size_t)r * 6;
            a[0] = __dadd_rn(a[0], uu[0]);
            a[1] = __dadd_rn(a[1], __dmul_rn(uu[0], uu[0]));
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j) a[2 + j * 2 + i] = __dadd_rn(a[2 + j * 2 + i], g[i * 2 + j]);
        }
    }
}

}  // namespace

void launch_rgnn(const uint32_t *planes, int L, int n_replicas, const double *W, double h, double *u_out, double *grad_out,
                 double *acc, int accumulate, cudaStream_t st) {
    // 14 replicas x 9 variants = 126 of the 128 threads of a block: a replica never straddles two blocks
    const int blocks = (n_replicas + 13) / 14;
    k_rgnn<<<blocks, 128, 0, st>>>(planes, L, n_replicas, W, h, u_out, grad_out, acc, accumulate);
}

}  // namespace mcrg
