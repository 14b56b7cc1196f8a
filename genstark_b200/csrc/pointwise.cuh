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
