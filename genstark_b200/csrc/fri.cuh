// K4: FRI layer fold, and K6: row gathers for the query phase.
//
// Replaces, per layer, transposeVector(domain,4,4^d) + interpolateQuarticBatch + evalQuarticBatch
// (lib/components/LowDegreeProver.ts:190-195).  Row i of a layer vector v of length L is
// (v[i], v[i+L/4], v[i+2L/4], v[i+3L/4]) at x-coordinates x_i * iota^j, x_i = w^(i*4^d), iota = w^(N/4)
// (LowDegreeProver.ts:270-282).  The cubic through those four points evaluated at the challenge x* is
//      column[i] = 1/4 * sum_k c_k t^k,   c_k = sum_j y_j iota^(-jk),   t = x* / x_i
// i.e. a radix-4 inverse butterfly followed by Horner -- 7 modmuls per output and no
// inversions, against a generic Lagrange interpolation in the reference.  Exact field arithmetic, so
// the values are identical.
#pragma once
#include "core.cuh"
#include "ntt.cuh"

namespace gs {

struct FriFoldParams {
    const fp* v;          // layer vector, length L
    fp* out;              // next layer, length L/4
    long long quarter;    // L/4
    const fp* special_x;  // device pointer to the challenge (so the host need not sync to launch)
    const fp* tw_lo; const fp* tw_hi; int log_g, log_lo;
    int x_shift;          // exponent of x_i in units of w_G:  e = i << x_shift  (= 4^d * G/N), i = GLOBAL row index
    int log_e, log_el, j0;   // coset sharding (see ComposeParams): local row il <-> global row i
    fp iota_inv;          // w^(-N/4)
    fp quarter_inv;       // 4^-1
};

__global__ void __launch_bounds__(256) fri_fold_kernel(const FriFoldParams P) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const fp xs = ldg_fp(P.special_x);
    const unsigned g_mask = (1u << P.log_g) - 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.quarter; i += stride) {
        const fp y0 = ld_fp(P.v + i), y1 = ld_fp(P.v + i + P.quarter);
        const fp y2 = ld_fp(P.v + i + 2 * P.quarter), y3 = ld_fp(P.v + i + 3 * P.quarter);
        // t = x* * x_i^-1,  x_i^-1 = w_G^(G - e)
        const unsigned ig = (unsigned)(((i >> P.log_el) << P.log_e) + P.j0 + (i & ((1ll << P.log_el) - 1)));
        const unsigned e = (0u - (ig << P.x_shift)) & g_mask;
        fp xinv = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
        if (P.log_g > P.log_lo) xinv = fp_mul(xinv, ldg_fp(P.tw_hi + (e >> P.log_lo)));
        const fp t = fp_mul(xs, xinv);
        // inverse radix-4 butterfly
        const fp s02 = fp_add(y0, y2), d02 = fp_sub(y0, y2);
        const fp s13 = fp_add(y1, y3), d13 = fp_mul(fp_sub(y1, y3), P.iota_inv);
        const fp c0 = fp_add(s02, s13), c2 = fp_sub(s02, s13);
        const fp c1 = fp_add(d02, d13), c3 = fp_sub(d02, d13);
        // Horner in t
        fp acc = fp_add(fp_mul(c3, t), c2);
        acc = fp_add(fp_mul(acc, t), c1);
        acc = fp_add(fp_mul(acc, t), c0);
        st_fp(P.out + i, fp_mul(acc, P.quarter_inv));
    }
}

// ---- K6 ------------------------------------------------------------------------------------------
// out[q][c] = col[c][idx[q]]  (merged leaf values, lib/Stark.ts:284-296; rowsToBuffers of a 4-column matrix)
struct GatherCols {
    const fp* col[64];
    int ncols;
};
__global__ void gather_rows_kernel(const GatherCols cols, const unsigned* __restrict__ idx, int nq, fp* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = t / cols.ncols, cidx = t % cols.ncols;
    if (q < nq) st_fp(out + (long long)q * cols.ncols + cidx, ld_fp(cols.col[cidx] + idx[q]));
}

// out[q] = nodes[idx[q]] (32-byte digests)
__global__ void gather_digests_kernel(const uint32_t* __restrict__ nodes, const unsigned* __restrict__ idx, int nq,
                                      uint32_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = t >> 1, half = t & 1;
    if (q < nq) reinterpret_cast<uint4*>(out)[2 * q + half] = reinterpret_cast<const uint4*>(nodes)[2ll * idx[q] + half];
}

// out[t] = 16 bytes at addr[t]: the whole query phase (every queried row and Merkle node of every tree)
// is one launch + one device->host copy
__global__ void gather_chunks_kernel(const unsigned long long* __restrict__ addr, int n, uint4* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = *reinterpret_cast<const uint4*>(addr[t]);
}

// natural[q*E + r*El + jl] = gathered[r][q*El + jl]  (32-byte digests): puts the all-gathered per-rank digests of a
// commit into leaf order
__global__ void permute_digests_kernel(const uint4* __restrict__ gathered, uint4* __restrict__ natural, long long n_loc, int log_el, int log_w) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per 16-byte half digest
    const long long total = (n_loc << log_w) * 2;
    if (t >= total) return;
    const long long d = t >> 1; const int half = (int)(t & 1);
    const long long r = d / n_loc, il = d % n_loc;
    const long long q = il >> log_el, jl = il & ((1ll << log_el) - 1);
    const long long i = (q << (log_el + log_w)) + (r << log_el) + jl;
    natural[2 * i + half] = gathered[2 * d + half];
}
// same permutation for 16-byte field elements (the FRI remainder)
__global__ void permute_elems_kernel(const uint4* __restrict__ gathered, uint4* __restrict__ natural, long long n_loc, int log_el, int log_w) {
    const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= (n_loc << log_w)) return;
    const long long r = d / n_loc, il = d % n_loc;
    const long long q = il >> log_el, jl = il & ((1ll << log_el) - 1);
    natural[(q << (log_el + log_w)) + (r << log_el) + jl] = gathered[d];
}

}  // namespace gs
