// Pipe-throughput microbenchmarks for the MCRG sweep kernel's instruction mix on B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Reports warp-instructions per cycle per SM (peak issue = 4) for: LOP3 chains, IMAD.WIDE chains, a 1:1 mix,
// bare Philox4x32-10 (calls/s) and Philox + the 4-plane lazy compare — the ceilings the sweep kernel is judged by.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
// MODE 0: 16 LOP3   1: 8 IMAD.WIDE (mul.wide.u32)   2: both   3: 16 IMAD (mad.lo.u32)   4: 16 POPC+IADD   5: 16 SHF (funnel shift)
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x0 = threadIdx.x + a0, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3;
    uint32_t y0 = b0 ^ x0, y1 = b0 + x1;
    uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x0) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x1) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x2) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x3) : "r"(y0), "r"(y1));
            }
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x0), "=r"(h0) : "r"(0xD2511F53u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x1), "=r"(h1) : "r"(0xCD9E8D57u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x2), "=r"(h2) : "r"(0xD2511F53u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x3), "=r"(h3) : "r"(0xCD9E8D57u));
            }
        }
        if (MODE == 3) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(y0), "r"(y1));
            }
        }
        if (MODE == 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                asm volatile("popc.b32 %0, %0;" : "+r"(x0)); asm volatile("popc.b32 %0, %0;" : "+r"(x1));
                asm volatile("popc.b32 %0, %0;" : "+r"(x2)); asm volatile("popc.b32 %0, %0;" : "+r"(x3));
            }
        }
        if (MODE == 5) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x0) : "r"(y0)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x1) : "r"(y0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x2) : "r"(y0)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x3) : "r"(y0));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ h0 ^ h1 ^ h2 ^ h3;
}

// Mixes of two instruction types on independent registers: do the pipes overlap?
//   A: 0 none, 1 = 16 LOP3, 2 = 16 IMAD (mad.lo), 3 = 8 IMAD.WIDE, 4 = 16 IADD3, 5 = 16 mul.hi.u32
//   B: 0 none, 1 = 16 POPC, 2 = 4 POPC, 3 = 16 SHF, 4 = 16 IMAD (mad.lo), 5 = 16 PRMT
template <int A, int B>
__global__ void __launch_bounds__(256) kmix(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x0 = threadIdx.x + a0, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3;
    uint32_t z0 = x0 ^ 0x1234567u, z1 = x1 ^ 0x89ABCDEu, z2 = x2 + 77u, z3 = x3 + 99u;
    uint32_t y0 = b0 ^ x0, y1 = b0 + x1;
    uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (A == 1) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x0) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x1) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x2) : "r"(y0), "r"(y1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x3) : "r"(y0), "r"(y1));
            }
            if (A == 2) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(y0), "r"(y1));
            }
            if (A == 3 && (u & 1)) {
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x0), "=r"(h0) : "r"(0xD2511F53u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x1), "=r"(h1) : "r"(0xCD9E8D57u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x2), "=r"(h2) : "r"(0xD2511F53u));
                asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0,%1}, p;\n\t}" : "+r"(x3), "=r"(h3) : "r"(0xCD9E8D57u));
            }
            if (A == 4) {
                asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(x0) : "r"(y0), "r"(y1));
                asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(x1) : "r"(y0), "r"(y1));
                asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(x2) : "r"(y0), "r"(y1));
                asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(x3) : "r"(y0), "r"(y1));
            }
            if (A == 5) {
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x0) : "r"(0xD2511F53u));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x1) : "r"(0xCD9E8D57u));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x2) : "r"(0xD2511F53u));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x3) : "r"(0xCD9E8D57u));
            }
            if (B == 1 || (B == 2 && u == 0)) {
                asm volatile("popc.b32 %0, %0;" : "+r"(z0)); asm volatile("popc.b32 %0, %0;" : "+r"(z1));
                asm volatile("popc.b32 %0, %0;" : "+r"(z2)); asm volatile("popc.b32 %0, %0;" : "+r"(z3));
            }
            if (B == 3) {
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(z0) : "r"(y0)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(z1) : "r"(y0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(z2) : "r"(y0)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(z3) : "r"(y0));
            }
            if (B == 4) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z0) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z1) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z2) : "r"(y0), "r"(y1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z3) : "r"(y0), "r"(y1));
            }
            if (B == 5) {
                asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(z0) : "r"(y0)); asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(z1) : "r"(y0));
                asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(z2) : "r"(y0)); asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(z3) : "r"(y0));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ h0 ^ h1 ^ h2 ^ h3 ^ z0 ^ z1 ^ z2 ^ z3;
}

__device__ __forceinline__ void philox(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1,%0}, p;\n\t}" : "=r"(h0), "=r"(l0) : "r"(0xD2511F53u), "r"(c0));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1,%0}, p;\n\t}" : "=r"(h1), "=r"(l1) : "r"(0xCD9E8D57u), "r"(c2));
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c1 = l1; c3 = l0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

template <int COMPARE, int ILP>
__global__ void __launch_bounds__(256) kphilox(uint32_t *out, uint32_t k0, uint32_t k1, const uint2 *tab) {
    __shared__ uint2 stab[32];
    if (threadIdx.x < 32) stab[threadIdx.x] = tab[threadIdx.x];
    __syncthreads();
    uint32_t acc = 0, eq = 0xFFFFFFFFu, sel = threadIdx.x * 2654435761u;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < ITER / 16; i += ILP) {
        uint32_t c[ILP][4];
#pragma unroll
        for (int q = 0; q < ILP; ++q) { c[q][0] = w; c[q][1] = 7; c[q][2] = i + q; c[q][3] = 0x10000000u; philox(c[q][0], c[q][1], c[q][2], c[q][3], k0, k1); }
#pragma unroll
        for (int q = 0; q < ILP; ++q) {
            if (COMPARE) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint2 t = stab[(4 * q + e) & 31];
                    const uint32_t tm = (sel & t.x) | (~sel & t.y);
                    acc |= eq & ~c[q][e] & tm;
                    eq &= ~(c[q][e] ^ tm);
                }
                eq |= c[q][0];
            } else acc ^= c[q][0] ^ c[q][1] ^ c[q][2] ^ c[q][3];
        }
    }
    out[w] = acc ^ eq;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8;
    uint32_t *out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    uint2 h[32]; for (int i = 0; i < 32; ++i) h[i] = make_uint2((i * 37) & 1 ? ~0u : 0u, (i * 11) & 2 ? ~0u : 0u);
    uint2 *tab; cudaMalloc(&tab, sizeof h); cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice);
    printf("%s, %d SMs, nominal max clock %.0f MHz; rates below assume the clock the GPU actually ran (see ncu/nvml)\n", p.name, sms, clk_khz / 1e3);
    const double warps = (double)blocks * 8;
    auto rep = [&](const char *name, float ms, double winstr_per_warp) {
        const double wi = warps * winstr_per_warp;
        printf("%-34s %8.3f ms  %7.1f G warp-instr/s  = %5.2f warp-instr/clk/SM at 1.92 GHz\n", name, ms, wi / ms / 1e6, wi / (ms * 1e-3) / sms / 1.92e9);
    };
    rep("LOP3 x16 (4 chains)", timeit([&] { k<0><<<blocks, 256>>>(out, 1, 2); }), ITER * 16.0);
    rep("IMAD.WIDE x8 (4 chains)", timeit([&] { k<1><<<blocks, 256>>>(out, 1, 2); }), ITER * 8.0);
    rep("LOP3 x16 + IMAD.WIDE x8", timeit([&] { k<2><<<blocks, 256>>>(out, 1, 2); }), ITER * 24.0);
    rep("IMAD (mad.lo) x16", timeit([&] { k<3><<<blocks, 256>>>(out, 1, 2); }), ITER * 16.0);
    rep("POPC x16", timeit([&] { k<4><<<blocks, 256>>>(out, 1, 2); }), ITER * 16.0);
    rep("SHF x16", timeit([&] { k<5><<<blocks, 256>>>(out, 1, 2); }), ITER * 16.0);
    // mixes: does the second type ride along for free (separate pipe) or add its own time (same pipe)?
    auto mix = [&](const char *name, float ms) { printf("%-44s %8.3f ms\n", name, ms); };
    mix("mix: 16 LOP3 alone", timeit([&] { kmix<1, 0><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 POPC alone", timeit([&] { kmix<0, 1><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 LOP3 + 16 POPC", timeit([&] { kmix<1, 1><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 LOP3 + 4 POPC", timeit([&] { kmix<1, 2><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 LOP3 + 16 SHF", timeit([&] { kmix<1, 3><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 LOP3 + 16 IMAD", timeit([&] { kmix<1, 4><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 LOP3 + 16 PRMT", timeit([&] { kmix<1, 5><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 PRMT alone", timeit([&] { kmix<0, 5><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 IMAD + 16 POPC", timeit([&] { kmix<2, 1><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 8 IMAD.WIDE alone", timeit([&] { kmix<3, 0><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 8 IMAD.WIDE + 16 POPC", timeit([&] { kmix<3, 1><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 8 IMAD.WIDE + 16 IMAD", timeit([&] { kmix<3, 4><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 8 IMAD.WIDE + 16 SHF", timeit([&] { kmix<3, 3><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 IADD3 (2 adds) alone", timeit([&] { kmix<4, 0><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 IADD3 + 16 POPC", timeit([&] { kmix<4, 1><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 mul.hi.u32 alone", timeit([&] { kmix<5, 0><<<blocks, 256>>>(out, 1, 2); }));
    mix("mix: 16 mul.hi.u32 + 16 IMAD", timeit([&] { kmix<5, 4><<<blocks, 256>>>(out, 1, 2); }));
    const double calls = (double)blocks * 256 * (ITER / 16);
    auto repp = [&](const char *name, float ms) {
        printf("%-34s %8.3f ms  %7.2f T Philox calls/s  = %6.1f clk/SM per warp-call at 1.92 GHz  (%.2f T bit-planes of 32 lanes/s)\n", name, ms,
               calls / ms / 1e9, (ms * 1e-3) * 1.92e9 * sms / (calls / 32), 4 * calls / ms / 1e9);
    };
    repp("Philox4x32-10, ILP1", timeit([&] { kphilox<0, 1><<<blocks, 256>>>(out, 5, 6, tab); }));
    repp("Philox4x32-10, ILP2", timeit([&] { kphilox<0, 2><<<blocks, 256>>>(out, 5, 6, tab); }));
    repp("Philox + 4-plane compare, ILP1", timeit([&] { kphilox<1, 1><<<blocks, 256>>>(out, 5, 6, tab); }));
    repp("Philox + 4-plane compare, ILP2", timeit([&] { kphilox<1, 2><<<blocks, 256>>>(out, 5, 6, tab); }));
    return 0;
}
