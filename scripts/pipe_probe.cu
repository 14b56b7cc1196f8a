// Integer-pipe and modular-multiplication probes for sm_100a: the numbers DESIGN.md's issue roofline rests on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/pipe_probe.bin scripts/pipe_probe.cu
//   (on the GPU box)  scripts/pipe_probe.bin > gpurun_out/pipe_probe.txt
// Part 1: warp-instruction issue rate per SM sub-partition (SMSP) of IMAD.WIDE.U32 / IMAD / IADD3 / LOP3 / SHF and of
//         FMA-pipe + ALU-pipe mixes (do the two pipes overlap?).
// Part 2: modular multiplications per second of the library's d_mul and of alternative reductions with the same 4x4 product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../genstark_b200/csrc/fp128.cuh"

using namespace gs;

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096, UNROLL = 8;

// ------------------------------------------------------------------------------------------ part 1
// kind: 0 IMAD.WIDE.U32, 1 IMAD (lo), 2 IADD3, 3 LOP3, 4 SHF, 5 IMAD + IADD3 interleaved 1:1, 6 IMAD.WIDE + 2 IADD3, 7 IMAD.WIDE + 4 IADD3
template <int KIND>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[UNROLL], b[UNROLL];
    uint64_t w[UNROLL];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = t * 2654435761u + i + seed; b[i] = a[i] ^ 0x9E3779B9u; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    const uint32_t m = seed | 3u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (KIND == 0) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi; }" : "+l"(w[i]));
            if (KIND == 6 || KIND == 7) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi; }" : "+l"(w[i]));
            if (KIND == 1 || KIND == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (KIND == 2 || KIND == 5) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[i]) : "r"(m), "r"(seed));
            if (KIND == 6 || KIND == 7) {
                asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[i]) : "r"(m), "r"(seed));
                asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(m), "r"(seed));
            }
            if (KIND == 7) {
                asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[i]) : "r"(seed), "r"(m));
                asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(seed), "r"(m));
            }
            if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (KIND == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (acc == 0x12345678u) out[t] = acc;
}

// ------------------------------------------------------------------------------------------ part 2
// V1: first fold with shifts instead of IMAD.WIDE by 9 (ALU pipe instead of FMA-heavy)
__device__ __forceinline__ fp reduce_v1(const uint32_t (&r)[8]) {
    const uint32_t h0 = r[4], h1 = r[5], h2 = r[6], h3 = r[7];
    // T = 9H = H + (H << 3), five limbs
    const uint32_t s0 = h0 << 3, s1 = __funnelshift_l(h0, h1, 3), s2 = __funnelshift_l(h1, h2, 3), s3 = __funnelshift_l(h2, h3, 3), s4 = h3 >> 29;
    uint32_t t0, t1, t2, t3, t4;
    asm("add.cc.u32 %0, %5, %10;\n\t addc.cc.u32 %1, %6, %11;\n\t addc.cc.u32 %2, %7, %12;\n\t addc.cc.u32 %3, %8, %13;\n\t addc.u32 %4, %9, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4) : "r"(h0), "r"(h1), "r"(h2), "r"(h3), "r"(s4), "r"(s0), "r"(s1), "r"(s2), "r"(s3));
    // U = L + (T << 32)
    uint32_t u0 = r[0], u1, u2, u3, u4, u5;
    asm("add.cc.u32 %0, %5, %8;\n\t addc.cc.u32 %1, %6, %9;\n\t addc.cc.u32 %2, %7, %10;\n\t addc.cc.u32 %3, %11, 0;\n\t addc.u32 %4, %12, 0;"
        : "=r"(u1), "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5) : "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4));
    uint32_t v0, v1, v2, v3, v4, v5;
    asm("sub.cc.u32 %0, %6, %12;\n\t subc.cc.u32 %1, %7, %13;\n\t subc.cc.u32 %2, %8, %14;\n\t subc.cc.u32 %3, %9, %15;\n\t subc.cc.u32 %4, %10, 0;\n\t subc.u32 %5, %11, 0;"
        : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5)
        : "r"(u0), "r"(u1), "r"(u2), "r"(u3), "r"(u4), "r"(u5), "r"(h0), "r"(h1), "r"(h2), "r"(h3));
    const uint64_t yy = (uint64_t)v4 * 9u;
    const uint32_t y0 = (uint32_t)yy, y1 = v5 * 9u + (uint32_t)(yy >> 32);
    uint32_t z0, z1, z2;
    asm("sub.cc.u32 %0, 0, %3;\n\t subc.cc.u32 %1, %4, %5;\n\t subc.u32 %2, %6, 0;" : "=r"(z0), "=r"(z1), "=r"(z2) : "r"(v4), "r"(y0), "r"(v5), "r"(y1));
    uint32_t w0, w1, w2, w3, co;
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, 0;\n\t addc.u32 %4, 0, 0;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(co) : "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(z0), "r"(z1), "r"(z2));
    uint32_t t0b, t1b, t2b, t3b, k;
    asm("add.cc.u32 %0, %5, 0xFFFFFFFF;\n\t addc.cc.u32 %1, %6, 8;\n\t addc.cc.u32 %2, %7, 0;\n\t addc.cc.u32 %3, %8, 0;\n\t addc.u32 %4, 0, 0;"
        : "=r"(t0b), "=r"(t1b), "=r"(t2b), "=r"(t3b), "=r"(k) : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    fp out; const bool sel = (co | k) != 0;
    out.v[0] = sel ? t0b : w0; out.v[1] = sel ? t1b : w1; out.v[2] = sel ? t2b : w2; out.v[3] = sel ? t3b : w3;
    return out;
}

// V2: the library's reduction without the final canonicalisation (result only < 2^128): what a lazy representation could save
__device__ __forceinline__ fp reduce_v2(const uint32_t (&r)[8]) {
    const uint64_t x0 = (uint64_t)r[4] * 9u + r[1], x1 = (uint64_t)r[5] * 9u + r[2], x2 = (uint64_t)r[6] * 9u + r[3], x3 = (uint64_t)r[7] * 9u;
    uint32_t u0 = r[0], u1 = (uint32_t)x0, u2, u3, u4, u5;
    asm("add.cc.u32 %0, %4, %5;\n\t addc.cc.u32 %1, %6, %7;\n\t addc.cc.u32 %2, %8, %9;\n\t addc.u32 %3, %10, 0;"
        : "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5)
        : "r"((uint32_t)x1), "r"((uint32_t)(x0 >> 32)), "r"((uint32_t)x2), "r"((uint32_t)(x1 >> 32)), "r"((uint32_t)x3), "r"((uint32_t)(x2 >> 32)), "r"((uint32_t)(x3 >> 32)));
    uint32_t v0, v1, v2, v3, v4, v5;
    asm("sub.cc.u32 %0, %6, %12;\n\t subc.cc.u32 %1, %7, %13;\n\t subc.cc.u32 %2, %8, %14;\n\t subc.cc.u32 %3, %9, %15;\n\t subc.cc.u32 %4, %10, 0;\n\t subc.u32 %5, %11, 0;"
        : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5)
        : "r"(u0), "r"(u1), "r"(u2), "r"(u3), "r"(u4), "r"(u5), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
    const uint64_t yy = (uint64_t)v4 * 9u;
    const uint32_t y0 = (uint32_t)yy, y1 = v5 * 9u + (uint32_t)(yy >> 32);
    uint32_t z0, z1, z2;
    asm("sub.cc.u32 %0, 0, %3;\n\t subc.cc.u32 %1, %4, %5;\n\t subc.u32 %2, %6, 0;" : "=r"(z0), "=r"(z1), "=r"(z2) : "r"(v4), "r"(y0), "r"(v5), "r"(y1));
    uint32_t w0, w1, w2, w3, co;
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, 0;\n\t addc.u32 %4, 0, 0;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(co) : "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(z0), "r"(z1), "r"(z2));
    // carry-out folds back as + (2^128 mod p) = 9*2^32 - 1; no comparison with p
    const uint32_t c_lo = 0u - co, c_hi = co << 3;
    fp out;
    asm("add.cc.u32 %0, %4, %8;\n\t addc.cc.u32 %1, %5, %9;\n\t addc.cc.u32 %2, %6, 0;\n\t addc.u32 %3, %7, 0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(out.v[3]) : "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(c_lo), "r"(c_hi));
    return out;
}

// V3: plain C with 64-bit limbs and unsigned __int128 (what the compiler makes of the textbook formulation)
__device__ __forceinline__ fp mul_v3(const fp& a, const fp& b) {
    typedef unsigned __int128 u128d;
    const uint64_t a0 = ((uint64_t)a.v[1] << 32) | a.v[0], a1 = ((uint64_t)a.v[3] << 32) | a.v[2];
    const uint64_t b0 = ((uint64_t)b.v[1] << 32) | b.v[0], b1 = ((uint64_t)b.v[3] << 32) | b.v[2];
    const u128d p00 = (u128d)a0 * b0, p01 = (u128d)a0 * b1, p10 = (u128d)a1 * b0, p11 = (u128d)a1 * b1;
    const u128d mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    const u128d L = ((u128d)(uint64_t)mid << 64) | (uint64_t)p00;
    const u128d H = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
    const u128d PP = (((u128d)0xFFFFFFFFFFFFFFFFull) << 64) | 0xFFFFFFF700000001ull, CC = ((u128d)8 << 32) | 0xFFFFFFFFull;
    auto canon = [&](u128d x) { return x >= PP ? x - PP : x; };
    auto add = [&](u128d x, u128d y) { u128d s = x + y; bool c = s < x; u128d t = s + CC; bool k = t < s; return (c | k) ? t : s; };
    auto sub = [&](u128d x, u128d y) { u128d d = x - y; return (x < y) ? d - CC : d; };
    const u128d h9lo = H * 9;
    const uint64_t h9hi = (uint64_t)((((H >> 64) * 9) + ((((u128d)(uint64_t)H) * 9) >> 64)) >> 64);
    const u128d top = (h9lo >> 96) | ((u128d)h9hi << 32);
    u128d r = add(canon(L), canon(h9lo << 32));
    r = sub(r, canon(H));
    r = add(r, top * (((u128d)9 << 32) - 1));
    fp o; o.v[0] = (uint32_t)r; o.v[1] = (uint32_t)(r >> 32); o.v[2] = (uint32_t)(r >> 64); o.v[3] = (uint32_t)(r >> 96);
    return o;
}

// (the library's d_* functions exist in the device pass only)
template <int V>
__device__ __forceinline__ fp mul_variant(const fp& a, const fp& b) {
#ifdef __CUDA_ARCH__
    if (V == 3) return mul_v3(a, b);
    uint32_t r[8];
    d_mul_wide(a, b, r);
    if (V == 1) return reduce_v1(r);
    if (V == 2) return reduce_v2(r);
    return d_reduce256(r);
#else
    return a;
#endif
}
__device__ __forceinline__ fp p_add(const fp& a, const fp& b) { return fp_add(a, b); }
__device__ __forceinline__ fp p_sub(const fp& a, const fp& b) { return fp_sub(a, b); }

// 4 independent chains per thread, like gs_debug_modmul_probe
template <int V>
__global__ void __launch_bounds__(256) modmul_kernel(fp* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    fp a = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1)), b = fp_from_u64(0xD1B54A32D192ED03ull * (t + 3));
    fp c2 = fp_from_u64(0x94D049BB133111EBull * (t + 5)), d = fp_from_u64(0xBF58476D1CE4E5B9ull * (t + 7));
    for (int i = 0; i < iters; ++i) {
        a = mul_variant<V>(a, b); b = mul_variant<V>(b, c2); c2 = mul_variant<V>(c2, d); d = mul_variant<V>(d, a);
    }
    fp r = p_add(p_add(a, b), p_add(c2, d));
    out[t] = r;
}

// butterfly-like mix: (u+v, (u-v)*w) -- one modmul with one add and one sub, the NTT's instruction mix
template <int V>
__global__ void __launch_bounds__(256) butterfly_kernel(fp* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    fp x[4], w = fp_from_u64(0x94D049BB133111EBull * (t + 5));
    for (int i = 0; i < 4; ++i) x[i] = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1 + i));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            fp u = x[2 * k], v = x[2 * k + 1];
            x[2 * k] = p_add(u, v);
            x[2 * k + 1] = mul_variant<V>(p_sub(u, v), w);
        }
        fp tmp = x[1]; x[1] = x[2]; x[2] = tmp;
    }
    out[t] = p_add(p_add(x[0], x[1]), p_add(x[2], x[3]));
}

template <typename F>
static float time_ms(F&& launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
}

int main() {
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, max SM clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
    uint32_t* d32; fp* dfp;
    const int blocks = sms * 8;
    CHECK(cudaMalloc(&d32, (size_t)blocks * 256 * 4)); CHECK(cudaMalloc(&dfp, (size_t)blocks * 256 * sizeof(fp)));
    // the run-time clock: a dependent IADD chain would do; use the nominal max and report rates per nominal cycle
    const double hz = clk_khz * 1e3;
    const char* names[8] = {"IMAD.WIDE.U32", "IMAD (lo)", "IADD3 (2 adds fused)", "LOP3", "SHF", "IMAD + IADD3", "IMAD.WIDE + 2 IADD3", "IMAD.WIDE + 4 IADD3"};
    const double per_iter[8] = {1, 1, 2, 1, 1, 3, 5, 9};
    float ms[8];
    ms[0] = time_ms([&] { pipe_kernel<0><<<blocks, 256>>>(d32, 1); });
    ms[1] = time_ms([&] { pipe_kernel<1><<<blocks, 256>>>(d32, 1); });
    ms[2] = time_ms([&] { pipe_kernel<2><<<blocks, 256>>>(d32, 1); });
    ms[3] = time_ms([&] { pipe_kernel<3><<<blocks, 256>>>(d32, 1); });
    ms[4] = time_ms([&] { pipe_kernel<4><<<blocks, 256>>>(d32, 1); });
    ms[5] = time_ms([&] { pipe_kernel<5><<<blocks, 256>>>(d32, 1); });
    ms[6] = time_ms([&] { pipe_kernel<6><<<blocks, 256>>>(d32, 1); });
    ms[7] = time_ms([&] { pipe_kernel<7><<<blocks, 256>>>(d32, 1); });
    printf("\n# part 1: issue rates (8 CTAs x 256 threads per SM, %d x %d independent ops per thread)\n", ITERS, UNROLL);
    printf("%-22s %10s %34s\n", "group", "ms", "issue cycles per group per SMSP");
    for (int k = 0; k < 8; ++k) {
        const double groups_per_smsp = (double)blocks * 8 /*warps*/ * ITERS * UNROLL / sms / 4.0;
        const double cycles = ms[k] * 1e-3 * hz;                       // nominal clock; every SM runs the same work
        printf("%-22s %10.4f %34.2f\n", names[k], ms[k], cycles / groups_per_smsp);
        (void)per_iter;
    }
    printf("(a group is what one loop body emits for one chain; its SASS content is in profiles/: e.g. 'IMAD + 2 add' = 1 IMAD + 1 IADD3)\n");

    printf("\n# part 2: modular multiplication over p = 2^128 - 9*2^32 + 1 (4 independent chains / thread)\n");
    const int iters = 2000;
    const char* vn[4] = {"V0 library d_mul", "V1 first fold by shifts", "V2 no final canonicalisation", "V3 plain C, u64/u128"};
    float mm[4], bf[4];
    mm[0] = time_ms([&] { modmul_kernel<0><<<blocks, 256>>>(dfp, iters); });
    mm[1] = time_ms([&] { modmul_kernel<1><<<blocks, 256>>>(dfp, iters); });
    mm[2] = time_ms([&] { modmul_kernel<2><<<blocks, 256>>>(dfp, iters); });
    mm[3] = time_ms([&] { modmul_kernel<3><<<blocks, 256>>>(dfp, iters); });
    bf[0] = time_ms([&] { butterfly_kernel<0><<<blocks, 256>>>(dfp, iters); });
    bf[1] = time_ms([&] { butterfly_kernel<1><<<blocks, 256>>>(dfp, iters); });
    bf[2] = time_ms([&] { butterfly_kernel<2><<<blocks, 256>>>(dfp, iters); });
    bf[3] = time_ms([&] { butterfly_kernel<3><<<blocks, 256>>>(dfp, iters); });
    printf("%-30s %10s %16s %20s | %10s %18s\n", "variant", "ms", "G modmul/s", "cyc/warp-modmul/SMSP", "bfly ms", "G butterflies/s");
    for (int v = 0; v < 4; ++v) {
        const double n = (double)blocks * 256 * 4 * iters;
        const double rate = n / (mm[v] * 1e-3);
        const double cyc = (mm[v] * 1e-3 * hz) / ((double)blocks * 8 * 4 * iters / sms / 4.0);
        const double nb = (double)blocks * 256 * 2 * iters;
        printf("%-30s %10.4f %16.1f %20.1f | %10.4f %18.1f\n", vn[v], mm[v], rate / 1e9, cyc, bf[v], nb / (bf[v] * 1e-3) / 1e9);
    }
    // correctness of the variants against V0 on the chain results
    fp *h0 = (fp*)malloc(256 * sizeof(fp)), *h1 = (fp*)malloc(256 * sizeof(fp));
    modmul_kernel<0><<<1, 256>>>(dfp, 50); cudaMemcpy(h0, dfp, 256 * sizeof(fp), cudaMemcpyDeviceToHost);
    for (int v = 1; v < 4; ++v) {
        if (v == 1) modmul_kernel<1><<<1, 256>>>(dfp, 50);
        if (v == 2) modmul_kernel<2><<<1, 256>>>(dfp, 50);
        if (v == 3) modmul_kernel<3><<<1, 256>>>(dfp, 50);
        cudaMemcpy(h1, dfp, 256 * sizeof(fp), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 256; ++i) for (int k = 0; k < 4; ++k) bad += h0[i].v[k] != h1[i].v[k];
        printf("variant V%d vs V0 after 50 iterations: %s\n", v, bad ? (v == 2 ? "differs (expected: not canonical)" : "MISMATCH") : "identical");
    }
    CHECK(cudaDeviceSynchronize());
    return 0;
}
