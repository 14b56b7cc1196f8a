// K1, two-pass form for 2^16 <= T <= 2^20 (the sizes prove() transforms at for BASELINE configs 4, 5 and the north-star shape).
//
// Replaces galois' interpolateRoots / evalPolysAtRoots as called from /root/reference/lib/Stark.ts:106,109.
//
// T = R_1 * R_2 with R_1 = 2^ceil(log T / 2), R_2 = 2^floor(log T / 2), both in {2^8, 2^9, 2^10}:
//   pass 1 ("column pass"): input index n = n_1 * R_2 + n_2.  For a tile of C consecutive columns n_2 the size-R_1 DFT
//     over n_1 (rows at stride R_2), times the inter-pass twiddle w_T^(n_2 * k_1), stored at [k_1][n_2] of the work buffer.
//   pass 2 ("final pass"): contiguous size-R_2 DFTs over n_2; output k_1 + R_1 * k_2 -- natural order.
// Two trips over HBM instead of three (ntt.cuh), one inter-pass multiplication per point instead of two.
//
// LDE (T coefficients -> E cosets of T evaluations, position q * E + j): coset j is the transform of x_j[pos] =
// coef[pos] * w_N^(pos * j) = x_{j-1}[pos] * w_N^pos.  A CTA of pass 1 walks the cosets of ONE coefficient tile: the tile
// is read from HBM once, the running product x_j lives in a CTA-private scratch tile (L2-resident, thread-private
// addresses, fully coalesced) and the step factors w_N^pos come from one T-entry table -- the (E-1)*T-entry coset table
// of ntt.cuh and its 112 MiB stream per LDE are gone.  Pass 2 takes C consecutive values of u = k_1 * E + j as its
// columns, so the natural-order store out[(k_1 + R_1 k_2) * E + j] is a C*16-byte run.
//
// Inside a tile (4096 elements = R rows x C columns, 256 threads x 16 register-resident elements): radix 16, then radix 8
// (or 16), then radix 8 / 4 -- measured in scripts/ntt_lab.cu: a register-resident radix-16 DIF at 4 warps per scheduler
// runs at 98 % of the modular-arithmetic issue roof, so the tile code keeps that shape and everything else (loads, the two
// exchanges, twiddles) is arranged around it.  Exchange 1 -> 2 goes through shared memory behind a CTA barrier; exchange
// 2 -> 3 stays inside a warp by construction (same k_1 mod 8), so it needs only __syncwarp.  Work is split into units
// (tile x coset) and every CTA takes one contiguous range of units (2 CTAs per SM, 148 SMs: a 2^20 LDE is 2048 units,
// 6.9 per CTA -- 98.8 % balanced, where 256 tiles over 296 CTA slots would be 86 %).
#pragma once
#include <cuda.h>
#include "ntt.cuh"

namespace gs {

struct Ntt2Params {
    const fp* src; long long src_row_stride;
    fp* dst; long long dst_row_stride;
    fp* xs;                       // LDE pass 1: CTA-private scratch, gridDim.x * 4096 elements
    const fp* tw_small;           // w_1024^i
    const fp* tw_inter;           // pass 1: [k_1][n_2] = w_T^(+-n_2 k_1)  (inverse: times T^-1)
    const fp* tw_step;            // LDE pass 1: [pos] = w_N^pos
    const fp* tw_lo; const fp* tw_hi; int log_g, log_lo;      // two-level root table: start factor of a range not at coset 0
    int inverse;
    int log_t, log_m;             // pass 1: m = R_2 columns per row of the input
    int log_ntot;                 // log2(T * E_total)
    int n_cosets, log_cosets, coset_base;      // cosets computed by this call (1, 0, 0 for a plain transform)
    unsigned units;               // rows * tiles * cosets (pass 1), rows * tiles (pass 2)
    int log_r1;                   // pass 2: radix of pass 1
};

template <int LOG_R>
struct TileShape {
    static constexpr int D2 = (LOG_R == 8) ? 4 : 3;
    static constexpr int D3 = LOG_R - 4 - D2;            // 0, 2, 3
    static constexpr int LOG_C = 12 - LOG_R, C = 1 << LOG_C;
    static constexpr int R2 = 1 << D2, R3 = 1 << D3, RR = R2 * R3;
    static constexpr int L3 = D3 + LOG_C;                 // low bits (a3, c) of a step-2 combination
    // shared-memory element index of (row, c); R = 1024 rows x 4 columns pads 64 B per 8 rows so that the step-3 reads
    // (8 rows apart per lane group) fall on different banks
    __device__ static __forceinline__ int addr(int row, int c) { return row * C + c + ((LOG_R == 10) ? ((row >> 3) << 2) : 0); }
};
static constexpr int NTT2_X_ELEMS = 4096 + 512;           // exchange buffer (72 KB)
static constexpr size_t NTT2_SMEM = (1024 + NTT2_X_ELEMS) * sizeof(fp) + 64;      // + mbarriers of the TMA variants

// The multiplication inside the tile code.  Inlined (default): ~80 instructions per site.  As a call (-DGS_NTT2_CALL_MUL) the code
// shrinks 185 KB -> 82 KB and the "no instruction" stalls vanish, but the loads can no longer be hoisted across the calls and 11 %
// more instructions issue: measured slower (profiles/r2_k1_ab.md), so the code is kept small by rolling the group loops instead.
#ifdef GS_NTT2_CALL_MUL
__device__ __noinline__ fp fp_mul_call(const fp a, const fp b) { return fp_mul(a, b); }
#define NTT2_MUL(a, b) fp_mul_call(a, b)
#else
#define NTT2_MUL(a, b) fp_mul(a, b)
#endif

// L2 residency hints: the twiddle tables are re-read by every coset of a tile while 128 MiB of output streams through the L2
GS_D unsigned long long l2_policy_keep() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
GS_D unsigned long long l2_policy_stream() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
GS_D fp ldg_hint_fp(const fp* p, unsigned long long pol) {
    uint4 t;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(pol));
    fp r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
GS_D void st_hint_fp(fp* p, const fp& a, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(p), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "l"(pol) : "memory");
}

// In-register DIF of size 2^S (dif_butterfly of ntt.cuh with the multiplication of this file).  X[k] ends up in x[brev<S>(k)].
template <int S>
GS_D void dif2(fp (&x)[1 << S], const fp* tw, int log_r) {
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
                else x[b + i + half] = NTT2_MUL(d, tw[(i << t) << (log_r - S)]);
            }
        }
    }
}

// w_G^e from the two-level table (e < G), forward direction only
GS_D fp tw2_lookup(const Ntt2Params& P, unsigned e) {
    fp lo = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
    if (P.log_g <= P.log_lo) return lo;
    fp hi = ldg_fp(P.tw_hi + (e >> P.log_lo));
    return fp_mul(lo, hi);
}

// Size-2^LOG_R DFT down the rows of the tile, in two parts.
// tile_front: x[a1] = element (row a1 * RR + rest, column c) of thread t = rest * C + c.  Runs step 1 (radix 16 in registers,
// twiddle, store to X) and, for R > 256, step 2 (radix 8 in place in X).  Leaves the tile in X for the last step.
// tile_final: the last radix over the remaining digit; calls emit(k, column, value) for the 16 outputs of the thread.
// The group loops of steps 2 and 3 are real loops (not unrolled): the unrolled tile was 109 - 185 KB of code and the two
// resident CTAs, out of phase, missed in the instruction caches (ncu "no instruction": 3.2 stall cycles per issue).
template <int LOG_R>
GS_D void tile_front(fp (&x)[16], fp* X, const fp* s_tw, int t) {
    using S = TileShape<LOG_R>;
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    dif2<4>(x, s_tw, 10);
    __syncthreads();                                      // the previous unit's last reads of X are done
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        fp v = x[brev<4>(k1)];
        if (k1 != 0) v = NTT2_MUL(v, s_tw[(rest * k1) << (10 - LOG_R)]);
        st_fp(&X[S::addr(k1 * S::RR + rest, c)], v);
    }
    __syncthreads();
    if (S::D3 == 0) return;
    // ---- step 2: radix R2 over a2 for fixed (k1, a3, c), in place
    constexpr int G2 = 16 / S::R2;
#pragma unroll 1
    for (int g = 0; g < G2; ++g) {
        const int q = g * 256 + t;
        const int lo = q & ((1 << S::L3) - 1), k1 = q >> S::L3;
        const int a3 = lo >> S::LOG_C, cc = lo & (S::C - 1);
        fp y[S::R2];
#pragma unroll
        for (int a2 = 0; a2 < S::R2; ++a2) y[a2] = ld_fp(&X[S::addr(k1 * S::RR + a2 * S::R3 + a3, cc)]);
        dif2<S::D2>(y, s_tw, 10);
#pragma unroll
        for (int k2 = 0; k2 < S::R2; ++k2) {
            fp v = y[brev<S::D2>(k2)];
            if (k2 != 0) v = NTT2_MUL(v, s_tw[(a3 * k2) << (10 - S::D2 - S::D3)]);
            st_fp(&X[S::addr(k1 * S::RR + k2 * S::R3 + a3, cc)], v);          // the addresses this thread just read
        }
    }
    __syncwarp();                                         // exchange 2 -> 3 is warp-local: writer and reader share k1 mod 8
}

template <int LOG_R, typename F>
GS_D void tile_final(const fp* X, const fp* s_tw, int t, F&& emit) {
    using S = TileShape<LOG_R>;
    if (S::D3 == 0) {
        // radix 16 over a2 for fixed (k1, c): one group per thread
        const int k1 = t >> S::L3, cc = t & (S::C - 1);
        fp y[16];
#pragma unroll
        for (int a2 = 0; a2 < 16; ++a2) y[a2] = ld_fp(&X[S::addr(k1 * 16 + a2, cc)]);
        dif2<4>(y, s_tw, 10);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) emit(k1 + 16 * k2, cc, y[brev<4>(k2)]);
    } else {
        constexpr int R3 = S::D3 > 0 ? S::R3 : 2, D3 = S::D3 > 0 ? S::D3 : 1;
        constexpr int G3 = 16 / R3, GU = (S::R2 * S::C) / 32;
        const int warp = t >> 5, lane = t & 31;
#pragma unroll 1
        for (int g = 0; g < G3; ++g) {
            const int k1 = (g / GU) * 8 + warp, u = (g % GU) * 32 + lane;
            const int k2 = u >> S::LOG_C, cc = u & (S::C - 1);
            fp z[R3];
#pragma unroll
            for (int a3 = 0; a3 < R3; ++a3) z[a3] = ld_fp(&X[S::addr((k1 * S::R2 + k2) * R3 + a3, cc)]);
            dif2<D3>(z, s_tw, 10);
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) emit(k1 + 16 * k2 + 16 * S::R2 * k3, cc, z[brev<D3>(k3)]);
        }
    }
}

// tile_final with all 16 inputs of the thread read from X up front: after `after_loads` X is no longer needed by this thread,
// which is what lets the next tile be prefetched into it (TMA variants below).  `mid` runs after the first group.
template <int LOG_R, typename A, typename M, typename F>
GS_D void tile_final_split(const fp* X, const fp* s_tw, int t, A&& after_loads, M&& mid, F&& emit) {
    using S = TileShape<LOG_R>;
    if (S::D3 == 0) {
        const int k1 = t >> S::L3, cc = t & (S::C - 1);
        fp y[16];
#pragma unroll
        for (int a2 = 0; a2 < 16; ++a2) y[a2] = ld_fp(&X[S::addr(k1 * 16 + a2, cc)]);
        after_loads();
        mid();
        dif2<4>(y, s_tw, 10);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) emit(k1 + 16 * k2, cc, y[brev<4>(k2)]);
    } else {
        constexpr int R3 = S::D3 > 0 ? S::R3 : 2, D3 = S::D3 > 0 ? S::D3 : 1;
        constexpr int G3 = 16 / R3, GU = (S::R2 * S::C) / 32;
        const int warp = t >> 5, lane = t & 31;
        fp z[16];
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int k1 = (g / GU) * 8 + warp, u = (g % GU) * 32 + lane;
            const int k2 = u >> S::LOG_C, cc = u & (S::C - 1);
#pragma unroll
            for (int a3 = 0; a3 < R3; ++a3) z[g * R3 + a3] = ld_fp(&X[S::addr((k1 * S::R2 + k2) * R3 + a3, cc)]);
        }
        after_loads();
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int k1 = (g / GU) * 8 + warp, u = (g % GU) * 32 + lane;
            const int k2 = u >> S::LOG_C, cc = u & (S::C - 1);
            fp w[R3];
#pragma unroll
            for (int a3 = 0; a3 < R3; ++a3) w[a3] = z[g * R3 + a3];
            dif2<D3>(w, s_tw, 10);
            if (g == 0) mid();
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) emit(k1 + 16 * k2 + 16 * S::R2 * k3, cc, w[brev<D3>(k3)]);
        }
    }
}

// ---- mbarrier / bulk-copy (TMA) primitives ---------------------------------------------------------------------------
GS_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
GS_D void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
GS_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GS_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
GS_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GS_D void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
GS_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
GS_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 3-D tensor-map tile load (cuTensorMapEncodeTiled on the host): box lands dense, row-major
GS_D void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

GS_D void ntt2_load_small_table(fp* s_tw, const fp* tw_small, int inverse) {
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        unsigned e = (unsigned)i;
        if (inverse) e = (1024u - e) & 1023u;
        s_tw[i] = ldg_fp(tw_small + e);
    }
    __syncthreads();                                      // the first radix-16 of the first unit reads entries other threads wrote
}

// ---------------------------------------------------------------------------------------------- pass 1
template <int LOG_R, bool LDE>
__global__ void __launch_bounds__(256, 2) ntt2_pass1_kernel(const Ntt2Params P) {
    using S = TileShape<LOG_R>;
    extern __shared__ __align__(128) unsigned char ntt2_smem[];
    fp* s_tw = reinterpret_cast<fp*>(ntt2_smem);
    fp* X = s_tw + 1024;
    const int t = threadIdx.x;
    ntt2_load_small_table(s_tw, P.tw_small, P.inverse);
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    const unsigned u_begin = (unsigned)(((unsigned long long)P.units * blockIdx.x) / gridDim.x);
    const unsigned u_end = (unsigned)(((unsigned long long)P.units * (blockIdx.x + 1)) / gridDim.x);
    const int log_tiles = P.log_m - S::LOG_C;
    fp* xs = LDE ? (P.xs + (size_t)blockIdx.x * 4096 + t) : nullptr;
    bool have_prev = false;                               // xs holds x_{jl-1} of the current tile
    const unsigned long long pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    for (unsigned u = u_begin; u < u_end; ++u) {
        const unsigned jl = u & ((1u << P.log_cosets) - 1u);
        const unsigned tile = (u >> P.log_cosets) & ((1u << log_tiles) - 1u);
        const unsigned row = u >> (P.log_cosets + log_tiles);
        const unsigned col0 = tile << S::LOG_C;
        const fp* src = P.src + (long long)row * P.src_row_stride;
        const unsigned pos0 = ((unsigned)rest << P.log_m) + col0 + c;          // position of a1 = 0; a1 adds a1 * RR * m
        fp x[16];
        const unsigned j = (unsigned)P.coset_base + jl;
        if (LDE && !have_prev && j != 0) {
            // a range that starts inside a tile (or a sharded call not at coset 0): xs <- coef * w_N^(pos * (j - 1)), rolled loop
#pragma unroll 1
            for (int a = 0; a < 16; ++a) {
                const unsigned pos = pos0 + ((unsigned)(a * S::RR) << P.log_m);
                fp v = ld_fp(src + pos);
                if (j > 1) v = fp_mul(v, tw2_lookup(P, (pos * (j - 1u)) << (P.log_g - P.log_ntot)));
                st_fp(xs + a * 256, v);
            }
            have_prev = true;
        }
        if (!LDE || !have_prev) {
#pragma unroll
            for (int a = 0; a < 16; ++a) x[a] = ld_fp(src + pos0 + ((unsigned)(a * S::RR) << P.log_m));
        } else {
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                const unsigned pos = pos0 + ((unsigned)(a * S::RR) << P.log_m);
                x[a] = NTT2_MUL(ld_fp(xs + a * 256), ldg_hint_fp(P.tw_step + pos, pol_keep));
            }
        }
        have_prev = LDE && (jl + 1u < (unsigned)P.n_cosets);
        if (have_prev && u + 1 < u_end) {
#pragma unroll
            for (int a = 0; a < 16; ++a) st_fp(xs + a * 256, x[a]);
        }
        tile_front<LOG_R>(x, X, s_tw, t);
        fp* dst = P.dst + (long long)row * P.dst_row_stride + ((size_t)jl << P.log_t) + col0;
        const fp* twi = P.tw_inter + col0;
        const int log_m = P.log_m, inverse = P.inverse;
        tile_final<LOG_R>(X, s_tw, t, [&](int k, int cc, fp v) {
            const size_t idx = ((size_t)k << log_m) + cc;
            if (k != 0 || inverse) v = NTT2_MUL(v, ldg_hint_fp(twi + idx, pol_keep));
            st_hint_fp(dst + idx, v, pol_stream);
        });
    }
}

// ---------------------------------------------------------------------------------------------- pass 2
template <int LOG_R>
__global__ void __launch_bounds__(256, 2) ntt2_pass2_kernel(const Ntt2Params P) {
    using S = TileShape<LOG_R>;
    extern __shared__ __align__(128) unsigned char ntt2_smem[];
    fp* s_tw = reinterpret_cast<fp*>(ntt2_smem);
    fp* X = s_tw + 1024;
    const int t = threadIdx.x;
    ntt2_load_small_table(s_tw, P.tw_small, P.inverse);
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    const unsigned u_begin = (unsigned)(((unsigned long long)P.units * blockIdx.x) / gridDim.x);
    const unsigned u_end = (unsigned)(((unsigned long long)P.units * (blockIdx.x + 1)) / gridDim.x);
    const int log_ucols = P.log_r1 + P.log_cosets;        // columns u = k_1 * cosets + j
    const int log_tiles = log_ucols - S::LOG_C;
    for (unsigned u = u_begin; u < u_end; ++u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        const unsigned ucol0 = tile << S::LOG_C, ucol = ucol0 + c;
        const unsigned jl = ucol & ((1u << P.log_cosets) - 1u), k1p = ucol >> P.log_cosets;
        const fp* col = P.src + (long long)row * P.src_row_stride + ((size_t)jl << P.log_t) + ((size_t)k1p << LOG_R) + rest;
        fp x[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) x[a] = ld_fp(col + a * S::RR);
        tile_front<LOG_R>(x, X, s_tw, t);
        fp* dst = P.dst + (long long)row * P.dst_row_stride + ucol0;
        tile_final<LOG_R>(X, s_tw, t, [&](int k, int cc, fp v) { st_fp(dst + ((size_t)k << log_ucols) + cc, v); });
    }
}

// pass 2 with the next tile prefetched by bulk copies (TMA engine) into the exchange buffer while the last radix of the current
// tile computes: one elected thread arms `full` with the byte count and issues C row copies; every thread arrives on `empty`
// once its last reads of X are in registers.
template <int LOG_R>
__global__ void __launch_bounds__(256, 2) ntt2_pass2_tma_kernel(const Ntt2Params P) {
    using S = TileShape<LOG_R>;
    constexpr int R = 1 << LOG_R, PITCH = R + 2;          // landing pitch: 32 B of padding keeps the column reads off one bank
    extern __shared__ __align__(128) unsigned char ntt2_smem[];
    fp* s_tw = reinterpret_cast<fp*>(ntt2_smem);
    fp* X = s_tw + 1024;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(X + NTT2_X_ELEMS);      // [0] full, [1] empty
    const int t = threadIdx.x;
    ntt2_load_small_table(s_tw, P.tw_small, P.inverse);
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    const unsigned u_begin = (unsigned)(((unsigned long long)P.units * blockIdx.x) / gridDim.x);
    const unsigned u_end = (unsigned)(((unsigned long long)P.units * (blockIdx.x + 1)) / gridDim.x);
    const int log_ucols = P.log_r1 + P.log_cosets;
    const int log_tiles = log_ucols - S::LOG_C;
    if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 256); mbar_fence_init(); }
    __syncthreads();
    auto issue = [&](unsigned u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        const unsigned ucol0 = tile << S::LOG_C;
        fence_proxy_async();
        mbar_expect_tx(&bars[0], 4096u * 16u);
#pragma unroll 1
        for (int cc = 0; cc < S::C; ++cc) {
            const unsigned ucol = ucol0 + cc;
            const unsigned jl = ucol & ((1u << P.log_cosets) - 1u), k1p = ucol >> P.log_cosets;
            const fp* src = P.src + (long long)row * P.src_row_stride + ((size_t)jl << P.log_t) + ((size_t)k1p << LOG_R);
            bulk_g2s(X + cc * PITCH, src, R * 16u, &bars[0]);
        }
    };
    if (t == 0 && u_begin < u_end) issue(u_begin);
    unsigned phase = 0;
    for (unsigned u = u_begin; u < u_end; ++u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        const unsigned ucol0 = tile << S::LOG_C;
        mbar_wait(&bars[0], phase);
        fp x[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) x[a] = ld_fp(X + c * PITCH + a * S::RR + rest);
        tile_front<LOG_R>(x, X, s_tw, t);                  // its first barrier: every thread holds its inputs, X is free for the exchange
        fp* dst = P.dst + (long long)row * P.dst_row_stride + ucol0;
        tile_final_split<LOG_R>(X, s_tw, t,
            [&]() { mbar_arrive(&bars[1]); },
            [&]() { if (t == 0 && u + 1 < u_end) { mbar_wait(&bars[1], phase); issue(u + 1); } },
            [&](int k, int cc, fp v) { st_fp(dst + ((size_t)k << log_ucols) + cc, v); });
        phase ^= 1u;
    }
}

// plain (non-LDE) pass 1 with its tiles fetched by the TMA engine: a tile is R rows of C * 16 contiguous bytes at a stride of
// m * 16 bytes -- a 2-D box of the source seen as a [rows][R][m * 4 x u32] tensor (cuTensorMapEncodeTiled on the host, boxes of at
// most 256 rows).  Same prefetch protocol as ntt2_pass2_tma_kernel; the box lands dense, row-major, which is the layout the
// first radix reads.  A/B against the plain-load kernel in profiles/ (GS_NTT2_TMA bit 1).
template <int LOG_R>
__global__ void __launch_bounds__(256, 2) ntt2_pass1_tma_kernel(const Ntt2Params P, const __grid_constant__ CUtensorMap tmap) {
    using S = TileShape<LOG_R>;
    constexpr int R = 1 << LOG_R, BOX_ROWS = R < 256 ? R : 256;
    extern __shared__ __align__(128) unsigned char ntt2_smem[];
    fp* s_tw = reinterpret_cast<fp*>(ntt2_smem);
    fp* X = s_tw + 1024;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(X + NTT2_X_ELEMS);
    const int t = threadIdx.x;
    ntt2_load_small_table(s_tw, P.tw_small, P.inverse);
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    const unsigned u_begin = (unsigned)(((unsigned long long)P.units * blockIdx.x) / gridDim.x);
    const unsigned u_end = (unsigned)(((unsigned long long)P.units * (blockIdx.x + 1)) / gridDim.x);
    const int log_tiles = P.log_m - S::LOG_C;
    if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 256); mbar_fence_init(); }
    __syncthreads();
    auto issue = [&](unsigned u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        fence_proxy_async();
        mbar_expect_tx(&bars[0], 4096u * 16u);
#pragma unroll 1
        for (int b = 0; b < R / BOX_ROWS; ++b)
            tma_load_3d(X + b * BOX_ROWS * S::C, &tmap, (int)(tile << S::LOG_C) * 4, b * BOX_ROWS, (int)row, &bars[0]);
    };
    if (t == 0 && u_begin < u_end) issue(u_begin);
    unsigned phase = 0;
    const unsigned long long pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    for (unsigned u = u_begin; u < u_end; ++u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        const unsigned col0 = tile << S::LOG_C;
        mbar_wait(&bars[0], phase);
        fp x[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) x[a] = ld_fp(X + (a * S::RR + rest) * S::C + c);
        tile_front<LOG_R>(x, X, s_tw, t);
        fp* dst = P.dst + (long long)row * P.dst_row_stride + col0;
        const fp* twi = P.tw_inter + col0;
        const int log_m = P.log_m, inverse = P.inverse;
        tile_final_split<LOG_R>(X, s_tw, t,
            [&]() { mbar_arrive(&bars[1]); },
            [&]() { if (t == 0 && u + 1 < u_end) { mbar_wait(&bars[1], phase); issue(u + 1); } },
            [&](int k, int cc, fp v) {
                const size_t idx = ((size_t)k << log_m) + cc;
                if (k != 0 || inverse) v = NTT2_MUL(v, ldg_hint_fp(twi + idx, pol_keep));
                st_hint_fp(dst + idx, v, pol_stream);
            });
        phase ^= 1u;
    }
}

// ---- warp-shuffle variant of the inner exchange (north-star item, measured in profiles/r2_k1_ab.md section 5) ---------------------
// For R = 1024 x 4 columns the exchange between radix steps 2 and 3 is an 8 x 8 transpose between the register index k_2 and lane
// bits 2..4 (a_3), for a fixed column c (lane bits 0..1): three rounds of butterfly swaps with __shfl_xor_sync, four shuffles per
// 128-bit element.  Steps 2 and 3 then run back to back on registers for each group and the tile never goes back to shared memory
// after the step-1 exchange.
GS_D fp shfl_xor_fp(const fp& a, int mask) {
    fp r;
    r.v[0] = __shfl_xor_sync(0xFFFFFFFFu, a.v[0], mask); r.v[1] = __shfl_xor_sync(0xFFFFFFFFu, a.v[1], mask);
    r.v[2] = __shfl_xor_sync(0xFFFFFFFFu, a.v[2], mask); r.v[3] = __shfl_xor_sync(0xFFFFFFFFu, a.v[3], mask);
    return r;
}
template <typename F>
GS_D void tile_tail_shfl_1024(const fp* X, const fp* s_tw, int t, F&& emit) {
    using S = TileShape<10>;
    const int warp = t >> 5, lane = t & 31, r = (lane >> 2) & 7, cc = lane & 3;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        const int k1 = g * 8 + warp;
        // step 2: this lane is (a3 = r, c): radix 8 over a2
        fp y[8];
#pragma unroll
        for (int a2 = 0; a2 < 8; ++a2) y[a2] = ld_fp(&X[S::addr(k1 * 64 + a2 * 8 + r, cc)]);
        dif2<3>(y, s_tw, 10);
        fp v[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) { v[k2] = y[brev<3>(k2)]; if (k2 != 0) v[k2] = NTT2_MUL(v[k2], s_tw[(r * k2) << 4]); }
        // transpose: afterwards this lane is (k2 = r, c) and v[a3] = element (k1, k2 = r, a3, c)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const bool bit = (r >> b) & 1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i & (1 << b)) continue;
                const int i2 = i | (1 << b);
                const fp send = bit ? v[i] : v[i2];
                const fp recv = shfl_xor_fp(send, 4 << b);
                if (bit) v[i] = recv; else v[i2] = recv;
            }
        }
        // step 3: radix 8 over a3
        dif2<3>(v, s_tw, 10);
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) emit(k1 + 16 * r + 128 * k3, cc, v[brev<3>(k3)]);
    }
}
__global__ void __launch_bounds__(256, 2) ntt2_pass2_shfl_kernel(const Ntt2Params P) {
    using S = TileShape<10>;
    extern __shared__ __align__(128) unsigned char ntt2_smem[];
    fp* s_tw = reinterpret_cast<fp*>(ntt2_smem);
    fp* X = s_tw + 1024;
    const int t = threadIdx.x;
    ntt2_load_small_table(s_tw, P.tw_small, P.inverse);
    const int rest = t >> S::LOG_C, c = t & (S::C - 1);
    const unsigned u_begin = (unsigned)(((unsigned long long)P.units * blockIdx.x) / gridDim.x);
    const unsigned u_end = (unsigned)(((unsigned long long)P.units * (blockIdx.x + 1)) / gridDim.x);
    const int log_ucols = P.log_r1 + P.log_cosets;
    const int log_tiles = log_ucols - S::LOG_C;
    for (unsigned u = u_begin; u < u_end; ++u) {
        const unsigned tile = u & ((1u << log_tiles) - 1u), row = u >> log_tiles;
        const unsigned ucol0 = tile << S::LOG_C, ucol = ucol0 + c;
        const unsigned jl = ucol & ((1u << P.log_cosets) - 1u), k1p = ucol >> P.log_cosets;
        const fp* col = P.src + (long long)row * P.src_row_stride + ((size_t)jl << P.log_t) + ((size_t)k1p << 10) + rest;
        fp x[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) x[a] = ld_fp(col + a * S::RR);
        // step 1 as in tile_front (radix 16, twiddle, exchange through X)
        dif2<4>(x, s_tw, 10);
        __syncthreads();
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            fp v = x[brev<4>(k1)];
            if (k1 != 0) v = NTT2_MUL(v, s_tw[rest * k1]);
            st_fp(&X[S::addr(k1 * S::RR + rest, c)], v);
        }
        __syncthreads();
        fp* dst = P.dst + (long long)row * P.dst_row_stride + ucol0;
        tile_tail_shfl_1024(X, s_tw, t, [&](int k, int cc, fp v) { st_fp(dst + ((size_t)k << log_ucols) + cc, v); });
    }
}

// tables ------------------------------------------------------------------------------------------
// [k_1][n_2] = w_T^(+-n_2 k_1) (times `scale` = T^-1 for the inverse transform, so the final pass has nothing left to scale)
__global__ void tw2_inter_table_kernel(Ntt2Params P, fp scale, int has_scale, fp* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << P.log_t)) return;
    const unsigned k = (unsigned)(i >> P.log_m), col = (unsigned)(i & (((size_t)1 << P.log_m) - 1));
    unsigned e = (col * k) << (P.log_g - P.log_t);         // col * k < T
    if (P.inverse) e = (0u - e) & ((1u << P.log_g) - 1u);
    fp v = tw2_lookup(P, e);
    if (has_scale) v = fp_mul(v, scale);
    st_fp(out + i, v);
}
// [pos] = w_N^pos, pos < T
__global__ void tw2_step_table_kernel(Ntt2Params P, fp* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << P.log_t)) return;
    st_fp(out + i, tw2_lookup(P, (unsigned)i << (P.log_g - P.log_ntot)));
}

}  // namespace gs
