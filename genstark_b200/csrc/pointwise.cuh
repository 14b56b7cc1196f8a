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

// expVectorElements(v, e): out[i] = v[i]^e, e < 2^128 given as four 32-bit words (examples/poseidon/utils.ts:37: the S-box of the
// plain Poseidon implementation the example checks its STARK against)
__global__ void __launch_bounds__(256) vec_exp_kernel(const fp* __restrict__ a, fp e, fp* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    int top = 127;
    while (top > 0 && !((e.v[top >> 5] >> (top & 31)) & 1u)) --top;
    const bool zero_exp = (e.v[0] | e.v[1] | e.v[2] | e.v[3]) == 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const fp x = ld_fp(a + i);
        fp r = fp_one();
        if (!zero_exp) {
            for (int bit = top; bit >= 0; --bit) {
                r = fp_sqr(r);
                if ((e.v[bit >> 5] >> (bit & 31)) & 1u) r = fp_mul(r, x);
            }
        }
        st_fp(out + i, r);
    }
}
// mulMatrixByVector(m, v): out[r] = sum_c m[r][c] * v[c]   (examples/poseidon/utils.ts:45: the MDS layer); one thread per row
__global__ void __launch_bounds__(128) mat_vec_kernel(const fp* __restrict__ m, const fp* __restrict__ v, fp* __restrict__ out, long long rows, long long cols) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    fp acc = fp_zero();
    for (long long c = 0; c < cols; ++c) acc = fp_add(acc, fp_mul(ld_fp(m + r * cols + c), ld_fp(v + c)));
    st_fp(out + r, acc);
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

// transposeMatrix(M): out[j][i] = M[i][j]   (LowDegreeProver.ts:181) -- tiled through shared memory, 16-byte elements
__global__ void __launch_bounds__(256) transpose_matrix_kernel(const fp* __restrict__ in, long long rows, long long cols, fp* __restrict__ out) {
    __shared__ uint4 tile[16][17];
    const long long c0 = (long long)blockIdx.x * 16, r0 = (long long)blockIdx.y * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    if (r0 + ty < rows && c0 + tx < cols) tile[ty][tx] = *reinterpret_cast<const uint4*>(in + (r0 + ty) * cols + c0 + tx);
    __syncthreads();
    if (c0 + ty < cols && r0 + tx < rows) *reinterpret_cast<uint4*>(out + (c0 + ty) * rows + r0 + tx) = tile[tx][ty];
}

// combineVectors(a, b) = sum_i a[i] * b[i]   (CompositionPolynomial.ts:168,188; LinearCombination.ts:85): per-block partial sums
__global__ void __launch_bounds__(256) dot_partial_kernel(const fp* __restrict__ a, const fp* __restrict__ b, long long n, fp* __restrict__ partial) {
    __shared__ fp red[256];
    fp acc = fp_zero();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc = fp_add(acc, fp_mul(ld_fp(a + i), ld_fp(b + i)));
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] = fp_add(red[threadIdx.x], red[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fp(partial + blockIdx.x, red[0]);
}

// interpolateQuarticBatch(xSets, ySets): per row the cubic through four points (LowDegreeProver.ts:137,191), Lagrange form with
// the four denominators inverted together; a zero denominator (repeated x) inverts to 0 like everywhere else.
// xs, ys, out: rows x 4, row-major.
__global__ void __launch_bounds__(128) quartic_interpolate_kernel(const fp* __restrict__ xs, const fp* __restrict__ ys, long long rows, fp* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    fp x[4], y[4], den[4], pre[4], inv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = ld_fp(xs + 4 * r + i); y[i] = ld_fp(ys + 4 * r + i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fp d = fp_one();
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j != i) d = fp_mul(d, fp_sub(x[i], x[j]));
        den[i] = d;
    }
    fp acc = fp_one();
#pragma unroll
    for (int i = 0; i < 4; ++i) { pre[i] = acc; if (!fp_is_zero(den[i])) acc = fp_mul(acc, den[i]); }
    fp ia = fp_inv(acc);
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        if (fp_is_zero(den[i])) { inv[i] = fp_zero(); continue; }
        inv[i] = fp_mul(ia, pre[i]);
        ia = fp_mul(ia, den[i]);
    }
    // numerator of point i: prod_{j != i} (X - x_j) = X^3 - e1 X^2 + e2 X - e3 over the three other points
    fp c0 = fp_zero(), c1 = fp_zero(), c2 = fp_zero(), c3 = fp_zero();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const fp a = x[(i + 1) & 3], b = x[(i + 2) & 3], c = x[(i + 3) & 3];
        const fp ab = fp_mul(a, b);
        const fp e1 = fp_add(fp_add(a, b), c);
        const fp e2 = fp_add(ab, fp_mul(c, fp_add(a, b)));
        const fp e3 = fp_mul(ab, c);
        const fp f = fp_mul(y[i], inv[i]);
        c3 = fp_add(c3, f);
        c2 = fp_sub(c2, fp_mul(f, e1));
        c1 = fp_add(c1, fp_mul(f, e2));
        c0 = fp_sub(c0, fp_mul(f, e3));
    }
    st_fp(out + 4 * r + 0, c0); st_fp(out + 4 * r + 1, c1); st_fp(out + 4 * r + 2, c2); st_fp(out + 4 * r + 3, c3);
}

// evalQuarticBatch(polys, x): Horner per row; x is one value per row (xs_is_vector) or the same scalar for every row
__global__ void __launch_bounds__(256) quartic_eval_kernel(const fp* __restrict__ polys, const fp* __restrict__ xs, fp x_scalar, int xs_is_vector,
                                                           long long rows, fp* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const fp x = xs_is_vector ? ld_fp(xs + r) : x_scalar;
    fp acc = ld_fp(polys + 4 * r + 3);
#pragma unroll
    for (int k = 2; k >= 0; --k) acc = fp_add(fp_mul(acc, x), ld_fp(polys + 4 * r + k));
    st_fp(out + r, acc);
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

// squaring chains (fp_sqr: 10 limb products) at the same shape as modmul_probe_kernel, and a check of fp_sqr against fp_mul(a, a)
// on edge values and the chain values themselves; *mismatch counts differences
__global__ void __launch_bounds__(256) sqr_probe_kernel(fp* out, int iters, unsigned* mismatch) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    fp a = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1)), b = fp_from_u64(0xD1B54A32D192ED03ull * (t + 3));
    fp c2 = fp_from_u64(0x94D049BB133111EBull * (t + 5)), d = fp_from_u64(0xBF58476D1CE4E5B9ull * (t + 7));
    if (mismatch) {
        fp edge[6];
        edge[0] = fp_zero(); edge[1] = fp_one();
        edge[2].v[0] = 0; edge[2].v[1] = 0xFFFFFFF7u; edge[2].v[2] = 0xFFFFFFFFu; edge[2].v[3] = 0xFFFFFFFFu;      // p - 1
        edge[3].v[0] = 0xFFFFFFFFu; edge[3].v[1] = 0xFFFFFFFFu; edge[3].v[2] = 0xFFFFFFFFu; edge[3].v[3] = 0x7FFFFFFFu;
        edge[4].v[0] = 0; edge[4].v[1] = 0; edge[4].v[2] = 0; edge[4].v[3] = 0x80000000u;
        edge[5].v[0] = 0xFFFFFFFFu; edge[5].v[1] = 8u; edge[5].v[2] = 0; edge[5].v[3] = 0;
        unsigned bad = 0;
        for (int k = 0; k < 6; ++k) bad += fp_eq(fp_sqr(edge[k]), fp_mul(edge[k], edge[k])) ? 0u : 1u;
        fp x = a;
        for (int k = 0; k < 64; ++k) { const fp s1 = fp_sqr(x), s2 = fp_mul(x, x); bad += fp_eq(s1, s2) ? 0u : 1u; x = fp_add(s1, b); }
        if (bad) atomicAdd(mismatch, bad);
    }
    for (int i = 0; i < iters; ++i) {     // 4 independent chains per thread
        a = fp_sqr(a); b = fp_sqr(b); c2 = fp_sqr(c2); d = fp_sqr(d);
    }
    fp r = fp_add(fp_add(a, b), fp_add(c2, d));
    if (r.v[0] == 0xFFFFFFFFu && r.v[3] == 0x12345u) st_fp(out, r);
}

// the NTT's instruction mix: (u, v) -> (u + v, (u - v) * w); 2 butterflies per thread per iteration (scripts/pipe_probe.cu)
__global__ void __launch_bounds__(256) butterfly_probe_kernel(fp* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    fp x[4], w = fp_from_u64(0x94D049BB133111EBull * (t + 5));
    for (int i = 0; i < 4; ++i) x[i] = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1 + i));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const fp u = x[2 * k], v = x[2 * k + 1];
            x[2 * k] = fp_add(u, v);
            x[2 * k + 1] = fp_mul(fp_sub(u, v), w);
        }
        const fp tmp = x[1]; x[1] = x[2]; x[2] = tmp;
    }
    fp r = fp_add(fp_add(x[0], x[1]), fp_add(x[2], x[3]));
    if (r.v[0] == 0xFFFFFFFFu && r.v[3] == 0x12345u) st_fp(out, r);
}

}  // namespace gs
