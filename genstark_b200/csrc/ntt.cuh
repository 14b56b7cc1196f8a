// K1: batched NTT / iNTT / coset low-degree extension over GF(p128), natural order in and out.
//
// Replaces galois' interpolateRoots / evalPolyAtRoots / evalPolysAtRoots as called from
// /root/reference/lib/Stark.ts:106,109 and lib/components/CompositionPolynomial.ts:109-110.
//
// Decomposition (Cooley-Tukey, mixed radix, "4-step" generalised to <= 3 passes over HBM):
//   n = N_1 * N_2 * ... * N_P,  input index  = n_1*(N_2..N_P) + ... + n_P,
//                               output index = k_1 + N_1*k_2 + N_1*N_2*k_3 + ...
//   pass p (p < P, "column pass", in place in the work buffer): for every prefix (k_1..k_{p-1}) and
//     every remaining index r in [0, m_p), m_p = N_{p+1}..N_P: size-N_p DFT over the N_p elements at
//     stride m_p, then multiply by w_{N_p*m_p}^(r*k_p).
//   pass P ("final pass", out of place): contiguous size-N_P DFTs whose outputs go to the digit-reversed
//     position; a tile covers C consecutive values of the LOWEST output digit so stores are C*16-byte runs.
//   LDE (T coefficients -> N = E*T evaluations) is the same transform with a leading digit of radix E
//   whose DFT is pruned away (inputs n >= T are zero): coset j reads the coefficients, multiplies by
//   w_N^(pos*j) on load, and j becomes the lowest output digit, so out[q*E + j] is contiguous in j.
//
// A tile is R x C elements (R = N_p <= 256 rows, C columns, 16 B each).  The size-R DFT runs as one
// or two register-resident radix-4/8/16 DIF butterflies (constants from a shared-memory table of
// w_R^i) with a single shared-memory exchange between them; the first step loads straight from global
// memory and the last step stores straight to it (both in >= 128-byte runs).
#pragma once
#include "fp128.cuh"

namespace gs {

struct NttPassParams {
    const fp* src;
    fp* dst;
    long long src_row_stride;     // batch rows (elements)
    long long dst_row_stride;
    // roots: master two-level table of w_G (G = 2^log_g): lo[i] = w^i (i < 2^log_lo), hi[i] = w^(i << log_lo)
    const fp* tw_lo;
    const fp* tw_hi;
    const fp* tw_small;           // w_1024^i, i < 1024 (same generator family)
    int log_g, log_lo;
    int inverse;                  // use w^-1 everywhere
    int final_pass;               // 0 = column pass, 1 = final pass (transposed store)
    int log_c;                    // tile columns
    // column pass
    int log_m;                    // remaining size m_p (row stride); 0 for final pass
    long long src_prefix_stride;  // 0 when the pass reads the shared coefficient vector (pruned LDE pass)
    int log_nsub;                 // log2(R * m): sub-problem size, inter-pass twiddle = w_nsub^(r*k)
    int coset_log_ntot;           // >0: multiply input at position pos of prefix j by w_ntot^(pos*(coset_base+j))
    int coset_base;               // first coset computed by this call (sharded LDE: a rank owns a range of cosets)
    // final pass
    int log_npre;                 // log2(number of prefixes)
    int log_d0, log_d1, log_d2;   // radices of the prefix digits, most significant first
    int log_ntot;                 // log2 of the full output length (incl. pruned digit)
    int has_scale;
    fp scale;                     // multiplied into every output of the final pass (n^-1 for inverse)
    // optional full tables (ntt_host.cuh builds them once per context and shape; null = two-level lookup + 1 modmul)
    const fp* tw_inter;           // [k][col] = w_nsub^(+-col*k), k < R, col < m      (column passes)
    const fp* tw_coset;           // [j-1][pos] = w_ntot^(pos*j), j = 1..E-1, pos < T  (pruned LDE pass)
    int log_t;                    // log2 T: row length of tw_coset
};

GS_DUAL_SOURCE(GS_FP_LDST_SRC,
GS_D fp ldg_fp(const fp* p) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    fp r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
GS_D fp ld_fp(const fp* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    fp r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
GS_D void st_fp(fp* p, const fp& a) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
)

// w_G^e from the two-level table (e < G)
GS_D fp tw_lookup(const NttPassParams& P, unsigned e) {
    const unsigned g_mask = (1u << P.log_g) - 1u;
    if (P.inverse) e = (0u - e) & g_mask;
    fp lo = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
    if (P.log_g <= P.log_lo) return lo;
    fp hi = ldg_fp(P.tw_hi + (e >> P.log_lo));
    return fp_mul(lo, hi);
}

template <int S>
GS_D constexpr int brev(int k) {
    int r = 0;
    for (int i = 0; i < S; ++i) r |= ((k >> i) & 1) << (S - 1 - i);
    return r;
}

// In-register DIF of size 2^S.  Output X[k] ends up in x[brev<S>(k)].
// w_{2^S}^i = tw[i << (log_r - S)]  (tw = shared table of w_R^i)
template <int S>
GS_D void dif_butterfly(fp (&x)[1 << S], const fp* tw, int log_r) {
#pragma unroll
    for (int t = 0; t < S; ++t) {
        const int len = (1 << S) >> t, half = len >> 1;
#pragma unroll
        for (int b = 0; b < (1 << S); b += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                fp u = x[b + i], v = x[b + i + half];
                x[b + i] = fp_add(u, v);
                fp d = fp_sub(u, v);
                if (i == 0) x[b + i + half] = d;
                else x[b + i + half] = fp_mul(d, tw[(i << t) << (log_r - S)]);
            }
        }
    }
}

// TAB: the pass reads its inter-pass (and coset) factors from the full tables; the other instantiation keeps the two-level
// lookups and carries none of the table code (a run-time switch cost both paths registers: measured, profiles/)
template <int LOG_R1, int LOG_R2, int MINB, bool TAB>
__global__ void __launch_bounds__(256, MINB) ntt_pass_kernel(const NttPassParams P) {
    constexpr int R1 = 1 << LOG_R1, R2 = 1 << LOG_R2, LOG_R = LOG_R1 + LOG_R2, R = 1 << LOG_R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fp* s_tw = reinterpret_cast<fp*>(smem_raw);     // R entries: w_R^i
    fp* s_tile = s_tw + R;                          // R x (C+1)
    const int C = 1 << P.log_c;
    const int pitch = C + 1;
    const int tid = threadIdx.x, nthr = blockDim.x;

    const fp* src = P.src + (long long)blockIdx.y * P.src_row_stride;
    fp* dst = P.dst + (long long)blockIdx.y * P.dst_row_stride;

    // table of w_R^i from w_1024^i
    for (int i = tid; i < R; i += nthr) {
        unsigned e = (unsigned)i << (10 - LOG_R);
        if (P.inverse) e = (1024u - e) & 1023u;
        s_tw[i] = ldg_fp(P.tw_small + e);
    }

    // tile coordinates
    long long src_base, dst_base;
    unsigned col0 = 0, pre = 0;        // column pass
    long long out_k_stride = 0;        // final pass: n_tot / R
    long long seg_stride = 0;          // final pass: elements between consecutive columns' segments
    const unsigned tile = blockIdx.x;
    if (!P.final_pass) {
        const int log_tiles_per_pre = P.log_m - P.log_c;
        pre = tile >> log_tiles_per_pre;
        col0 = (tile & ((1u << log_tiles_per_pre) - 1u)) << P.log_c;
        src_base = (long long)pre * P.src_prefix_stride + col0;
        dst_base = ((long long)pre << P.log_nsub) + col0;
    } else {
        const int log_rest = P.log_npre - P.log_d0;
        const unsigned rest = tile & ((1u << log_rest) - 1u);
        const unsigned c0 = (tile >> log_rest) << P.log_c;
        const unsigned d1 = rest >> P.log_d2, d2 = rest & ((1u << P.log_d2) - 1u);
        const unsigned rev = d1 + (d2 << P.log_d1);
        seg_stride = (long long)R << log_rest;
        src_base = (long long)c0 * seg_stride + (long long)rest * R;
        dst_base = (long long)c0 + ((long long)rev << P.log_d0);
        out_k_stride = 1ll << (P.log_ntot - LOG_R);
    }
    __syncthreads();

    // ------------------------------------------------------------------ step 1: radix R1 over a1
    // groups (r, c): r in [0,R2), c in [0,C).  column pass: c fastest; final pass: r fastest.
    const int ngroups1 = R2 * C;
    for (int g = tid; g < ngroups1; g += nthr) {
        int r, c;
        if (!P.final_pass) { c = g & (C - 1); r = g >> P.log_c; }
        else { r = g & (R2 - 1); c = g >> LOG_R2; }
        fp x[R1];
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            const int row = a * R2 + r;
            const fp* ptr = P.final_pass ? (src + src_base + (long long)c * seg_stride + row)
                                         : (src + src_base + ((long long)row << P.log_m) + c);
            x[a] = ld_fp(ptr);
        }
        if (P.coset_log_ntot > 0 && (pre + (unsigned)P.coset_base) != 0) {
            // w_ntot^(pos * j), pos = row*m + col
#pragma unroll
            for (int a = 0; a < R1; ++a) {
                const unsigned pos = ((unsigned)(a * R2 + r) << P.log_m) + col0 + c;
                const unsigned j = pre + (unsigned)P.coset_base;
                if (TAB) x[a] = fp_mul(x[a], ldg_fp(P.tw_coset + ((size_t)(j - 1) << P.log_t) + pos));
                else x[a] = fp_mul(x[a], tw_lookup(P, (pos * j) << (P.log_g - P.coset_log_ntot)));
            }
        }
        dif_butterfly<LOG_R1>(x, s_tw, LOG_R);
        if (LOG_R2 > 0) {
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                fp v = x[brev<LOG_R1>(k1)];
                if (k1 != 0) v = fp_mul(v, s_tw[r * k1]);
                st_fp(&s_tile[(k1 * R2 + r) * pitch + c], v);
            }
        } else {
            // single-step pass: outputs go straight to global memory
#pragma unroll
            for (int k = 0; k < R1; ++k) {
                fp v = x[brev<LOG_R1>(k)];
                if (!P.final_pass) {
                    if (k != 0) {
                        if (TAB) v = fp_mul(v, ldg_fp(P.tw_inter + ((size_t)k << P.log_m) + col0 + c));
                        else v = fp_mul(v, tw_lookup(P, ((col0 + c) * (unsigned)k) << (P.log_g - P.log_nsub)));
                    }
                    st_fp(dst + dst_base + ((long long)k << P.log_m) + c, v);
                } else {
                    if (P.has_scale) v = fp_mul(v, P.scale);
                    st_fp(dst + dst_base + (long long)k * out_k_stride + c, v);
                }
            }
        }
    }
    if (LOG_R2 == 0) return;
    __syncthreads();

    // ------------------------------------------------------------------ step 2: radix R2 over a2
    const int ngroups2 = R1 * C;
    for (int g = tid; g < ngroups2; g += nthr) {
        const int c = g & (C - 1), k1 = g >> P.log_c;
        fp x[R2 > 0 ? R2 : 1];
#pragma unroll
        for (int a = 0; a < R2; ++a) x[a] = ld_fp(&s_tile[(k1 * R2 + a) * pitch + c]);
        dif_butterfly<LOG_R2>(x, s_tw, LOG_R);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            fp v = x[brev<LOG_R2>(k2)];
            const int k = k1 + R1 * k2;
            if (!P.final_pass) {
                if (k != 0) {
                    if (TAB) v = fp_mul(v, ldg_fp(P.tw_inter + ((size_t)k << P.log_m) + col0 + c));
                    else v = fp_mul(v, tw_lookup(P, ((col0 + c) * (unsigned)k) << (P.log_g - P.log_nsub)));
                }
                st_fp(dst + dst_base + ((long long)k << P.log_m) + c, v);
            } else {
                if (P.has_scale) v = fp_mul(v, P.scale);
                st_fp(dst + dst_base + (long long)k * out_k_stride + c, v);
            }
        }
    }
}

// full twiddle tables (built once per context and shape from the two-level table)
__global__ void tw_inter_table_kernel(NttPassParams P, int log_r, fp* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // k * m + col
    if (i >= ((size_t)1 << P.log_nsub)) return;
    const unsigned k = (unsigned)(i >> P.log_m), col = (unsigned)(i & (((size_t)1 << P.log_m) - 1));
    st_fp(out + i, tw_lookup(P, (col * k) << (P.log_g - P.log_nsub)));
}
__global__ void tw_coset_table_kernel(NttPassParams P, int n_cosets_minus_1, fp* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (j-1) * T + pos
    if (i >= ((size_t)n_cosets_minus_1 << P.log_t)) return;
    const unsigned j = (unsigned)(i >> P.log_t) + 1u, pos = (unsigned)(i & (((size_t)1 << P.log_t) - 1));
    st_fp(out + i, tw_lookup(P, (pos * j) << (P.log_g - P.coset_log_ntot)));
}

}  // namespace gs
