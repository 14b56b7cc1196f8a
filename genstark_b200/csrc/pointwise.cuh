// K2 (basic part): element-wise vector operations of the FiniteField seam --
// addVectorElements / subVectorElements / mulVectorElements with a vector or a scalar right operand
// (call sites: lib/components/CompositionPolynomial.ts:98,120,136,145; LinearCombination.ts:50,63;
// ZeroPolynomial.ts:41-42).  HBM-bound: 16 B in (x2) + 16 B out per element, 128-bit accesses.
#pragma once
#include "core.cuh"
#include "ntt.cuh"

namespace gs {

enum { VOP_ADD = 0, VOP_SUB = 1, VOP_MUL = 2 };

template <int OP, bool SCALAR>
__global__ void __launch_bounds__(256) vec_binary_kernel(const fp* __restrict__ a, const fp* __restrict__ b,
                                                         fp s, fp* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fp x = ld_fp(a + i);
        fp y = SCALAR ? s : ld_fp(b + i);
        fp r = (OP == VOP_ADD) ? fp_add(x, y) : (OP == VOP_SUB) ? fp_sub(x, y) : fp_mul(x, y);
        st_fp(out + i, r);
    }
}

static inline int vec_binary(Ctx* c, int op, const fp* a, const fp* b, const fp* scalar, fp* out, long long n) {
    const int threads = 256;
    long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)c->sm_count * 16;
    if (blocks > cap) blocks = cap;
    fp s = scalar ? *scalar : fp_zero();
#define GS_LAUNCH_VB(OP)                                                                               \
    if (scalar) vec_binary_kernel<OP, true><<<(unsigned)blocks, threads, 0, c->stream>>>(a, b, s, out, n); \
    else vec_binary_kernel<OP, false><<<(unsigned)blocks, threads, 0, c->stream>>>(a, b, s, out, n)
    switch (op) {
        case VOP_ADD: GS_LAUNCH_VB(VOP_ADD); break;
        case VOP_SUB: GS_LAUNCH_VB(VOP_SUB); break;
        case VOP_MUL: GS_LAUNCH_VB(VOP_MUL); break;
        default: return c->fail(GS_E_ARG, "unknown vector op %d", op);
    }
#undef GS_LAUNCH_VB
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "vec_binary_kernel");
    c->launches++;
    return GS_OK;
}

// combineManyVectors(V, k): out[i] = sum_m k[m] * V[m][i]   (CompositionPolynomial.ts:105,142; LinearCombination.ts:60)
struct CombineParams { const fp* v[64]; fp k[64]; int m; };
__global__ void __launch_bounds__(256) combine_many_kernel(const CombineParams* __restrict__ P, fp* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fp acc = fp_zero();
        for (int m = 0; m < P->m; ++m) acc = fp_add(acc, fp_mul(ld_fp(P->v[m] + i), P->k[m]));
        st_fp(out + i, acc);
    }
}

// getPowerSeries(base, n): out[i] = base^i.  Each thread starts from base^(first index) by square-and-multiply
// and then steps by base^(total threads).
__global__ void __launch_bounds__(256) power_series_kernel(fp base, fp* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fp v = fp_pow(base, (uint64_t)t);
    const fp step = fp_pow(base, (uint64_t)stride);
    for (long long i = t; i < n; i += stride) { st_fp(out + i, v); v = fp_mul(v, step); }
}

// pluckVector(v, skip, times): out[i] = v[(i*skip) mod len]      (ZeroPolynomial.ts:40)
__global__ void pluck_kernel(const fp* __restrict__ v, long long len, long long skip, fp* __restrict__ out, long long times) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < times) st_fp(out + i, ld_fp(v + (unsigned long long)((unsigned __int128)i * skip % len)));
}

// transposeVector(v, columns, step): rows = len/(columns*step); M[i][j] = v[(i + j*rows)*step], row-major out
__global__ void transpose_vector_kernel(const fp* __restrict__ v, long long rows, int columns, long long step, fp* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * columns) return;
    const long long i = t / columns; const int j = (int)(t % columns);
    st_fp(out + t, ld_fp(v + (i + (long long)j * rows) * step));
}

// dependent modmul chains: throughput probe used by bench.py to state the integer roofline
__global__ void __launch_bounds__(256) modmul_probe_kernel(fp* out, int iters) {
    fp a, b, c2, d;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    a = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1)); b = fp_from_u64(0xD1B54A32D192ED03ull * (t + 3));
    c2 = fp_from_u64(0x94D049BB133111EBull * (t + 5)); d = fp_from_u64(0xBF58476D1CE4E5B9ull * (t + 7));
    for (int i = 0; i < iters; ++i) {     // 4 independent chains per thread
        a = fp_mul(a, b); b = fp_mul(b, c2); c2 = fp_mul(c2, d); d = fp_mul(d, a);
    }
    fp r = fp_add(fp_add(a, b), fp_add(c2, d));
    if (r.v[0] == 0xFFFFFFFFu && r.v[3] == 0x12345u) st_fp(out, r);   // keep the chains alive
}

}  // namespace gs
