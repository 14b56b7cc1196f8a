// GF(p), p = 2^128 - 9*2^32 + 1  (the 128-bit STARK field: /root/reference/examples/mimc/mimc128.ts:13,
// assembly/lib128.aa:3).  Device + host arithmetic on canonical residues held as four 32-bit limbs,
// little-endian -- the same bytes the reference's Vector.toBuffer()/copyValue() produce
// (lib/utils/serialization.ts:131-147), so no conversion sits between arithmetic and hashing.
//
// Reduction uses the shape of p directly: 2^128 = 9*2^32 - 1 (mod p).  A 256-bit product
// L + H*2^128 folds to L + 9H*2^32 - H (165 bits), folds once more (top 37 bits), and ends with one
// conditional add of 2^128 - p.  Multiplication is a 4x4 limb product laid out as even/odd columns
// so that ptxas pairs every mad.lo.cc/madc.hi.cc into one IMAD.WIDE.U32 (16 per product).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define GS_HD __host__ __device__ __forceinline__
#define GS_D __device__ __forceinline__
#define GS_ALIGN16 __align__(16)
#else
#define GS_HD inline
#define GS_D inline
#define GS_ALIGN16 alignas(16)
#endif

// Code that exists twice from one spelling: compiled into the library and kept as text for the run-time
// specialised kernels (devjit.cuh hands it to NVRTC together with the code generated for an AIR).
#ifndef GS_DUAL_SOURCE
#define GS_DUAL_SOURCE(name, ...) __VA_ARGS__ static const char name[] = #__VA_ARGS__;
#endif
#ifdef __CUDA_ARCH__
#define GS_DEVICE_DUAL_SOURCE(name, ...) __VA_ARGS__ static const char name[] = #__VA_ARGS__;
#else
#define GS_DEVICE_DUAL_SOURCE(name, ...) static const char name[] = #__VA_ARGS__;
#endif

namespace gs {

GS_DUAL_SOURCE(GS_FP_TYPE_SRC,
struct GS_ALIGN16 fp {
    uint32_t v[4];
};
)

// p and 2^128 - p
static constexpr uint32_t P0 = 0x00000001u, P1 = 0xFFFFFFF7u, P2 = 0xFFFFFFFFu, P3 = 0xFFFFFFFFu;
static constexpr uint32_t C0 = 0xFFFFFFFFu, C1 = 0x00000008u;   // 2^128 - p = 9*2^32 - 1

GS_DUAL_SOURCE(GS_FP_BASIC_SRC,
GS_HD fp fp_zero() { fp r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
GS_HD fp fp_one() { fp r; r.v[0] = 1; r.v[1] = r.v[2] = r.v[3] = 0; return r; }
GS_HD fp fp_from_u64(uint64_t x) { fp r; r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); r.v[2] = r.v[3] = 0; return r; }
GS_HD bool fp_is_zero(const fp& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
GS_HD bool fp_eq(const fp& a, const fp& b) {
    return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0;
}
)

// ------------------------------------------------------------------------------------------------ host
// Host-side twin used by the prover's control path (twiddle roots, challenges, trace generation).
typedef unsigned __int128 u128;
static inline u128 fp_to_u128(const fp& a) {
    return ((u128)a.v[3] << 96) | ((u128)a.v[2] << 64) | ((u128)a.v[1] << 32) | a.v[0];
}
static inline fp fp_from_u128(u128 x) {
    fp r; r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); r.v[2] = (uint32_t)(x >> 64); r.v[3] = (uint32_t)(x >> 96);
    return r;
}
static constexpr u128 HP = (((u128)0xFFFFFFFFFFFFFFFFull) << 64) | 0xFFFFFFF700000001ull;
static constexpr u128 HC = ((u128)8 << 32) | 0xFFFFFFFFull;      // 2^128 - p

static inline u128 h_canon(u128 x) { return x >= HP ? x - HP : x; }
static inline u128 h_add(u128 a, u128 b) {
    u128 s = a + b;
    bool c = s < a;
    u128 t = s + HC;
    bool k = t < s;
    return (c | k) ? t : s;
}
static inline u128 h_sub(u128 a, u128 b) {
    u128 d = a - b;
    return (a < b) ? d - HC : d;
}
static inline u128 h_mul(u128 a, u128 b) {
    uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64), b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    u128 L = ((u128)(uint64_t)mid << 64) | (uint64_t)p00;
    u128 H = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
    // x = L + H*2^128 == L + 9H*2^32 - H
    // 9H*2^32 = (9H mod 2^96) * 2^32  +  (9H >> 96) * 2^128
    u128 h9lo = H * 9;                                         // low 128 bits of 9H
    uint64_t h9hi = (uint64_t)((((H >> 64) * 9) + ((((u128)(uint64_t)H) * 9) >> 64)) >> 64);  // bits 128.. of 9H
    u128 top = (h9lo >> 96) | ((u128)h9hi << 32);              // 9H >> 96  (< 2^36)
    u128 r = h_add(h_canon(L), h_canon(h9lo << 32));
    r = h_sub(r, h_canon(H));
    // top * 2^128 == top * (9*2^32 - 1)
    u128 t2 = top * (((u128)9 << 32) - 1);                     // < 2^72
    return h_add(r, t2);
}
static inline u128 h_pow(u128 b, u128 e) {
    u128 r = 1;
    while (e) { if (e & 1) r = h_mul(r, b); b = h_mul(b, b); e >>= 1; }
    return r;
}
static inline u128 h_inv(u128 a) { return a == 0 ? 0 : h_pow(a, HP - 2); }
// galois getRootOfUnity: smallest i >= 2 whose g = i^((p-1)/order) has exact order.  The test
// g^(order/2) != 1 is i^((p-1)/2) != 1, independent of the order, so every order shares the same i (= 3) and
// w_{n/2} = w_n^2.
static inline u128 h_root_of_unity(int log_order) {
    const u128 pm1 = HP - 1;
    for (u128 i = 2;; ++i) {
        if (h_pow(i, pm1 >> 1) != 1) return h_pow(i, pm1 >> log_order);
    }
}

// ---------------------------------------------------------------------------------------------- device
GS_DEVICE_DUAL_SOURCE(GS_FP_DEVICE_SRC,
GS_D fp d_add(const fp& a, const fp& b) {
    uint32_t s0, s1, s2, s3, c, t0, t1, t2, t3, k;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    // s >= p  <=>  s + (2^128 - p) carries out of 128 bits
    asm("add.cc.u32 %0, %5, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 %1, %6, 8;\n\t"
        "addc.cc.u32 %2, %7, 0;\n\t"
        "addc.cc.u32 %3, %8, 0;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(k)
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3));
    fp r;
    bool sel = (c | k) != 0;
    r.v[0] = sel ? t0 : s0; r.v[1] = sel ? t1 : s1; r.v[2] = sel ? t2 : s2; r.v[3] = sel ? t3 : s3;
    return r;
}

GS_D fp d_sub(const fp& a, const fp& b) {
    uint32_t d0, d1, d2, d3, m;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"          // 0 or 0xFFFFFFFF
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(m)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    // on borrow add p, i.e. subtract 2^128 - p modulo 2^128
    fp r;
    uint32_t m1 = m & 8u;
    asm("sub.cc.u32 %0, %4, %8;\n\t"
        "subc.cc.u32 %1, %5, %9;\n\t"
        "subc.cc.u32 %2, %6, 0;\n\t"
        "subc.u32 %3, %7, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3])
        : "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(m), "r"(m1));
    return r;
}

GS_D fp d_neg(const fp& a) {
    fp z = fp_zero();
    return d_sub(z, a);
}

// fold a 256-bit value r[0..7] to the canonical residue
GS_D fp d_reduce256(const uint32_t (&r)[8]) {
    // X_i = 9*h_i + r_{i+1}   (h = r[4..7]);  U = r0 + (sum X_i << 32(i+1))
    const uint64_t x0 = (uint64_t)r[4] * 9u + r[1];
    const uint64_t x1 = (uint64_t)r[5] * 9u + r[2];
    const uint64_t x2 = (uint64_t)r[6] * 9u + r[3];
    const uint64_t x3 = (uint64_t)r[7] * 9u;
    const uint32_t lo0 = (uint32_t)x0, hi0 = (uint32_t)(x0 >> 32), lo1 = (uint32_t)x1, hi1 = (uint32_t)(x1 >> 32);
    const uint32_t lo2 = (uint32_t)x2, hi2 = (uint32_t)(x2 >> 32), lo3 = (uint32_t)x3, hi3 = (uint32_t)(x3 >> 32);
    uint32_t u0 = r[0], u1 = lo0, u2, u3, u4, u5;
    asm("add.cc.u32 %0, %4, %5;\n\t"
        "addc.cc.u32 %1, %6, %7;\n\t"
        "addc.cc.u32 %2, %8, %9;\n\t"
        "addc.u32 %3, %10, 0;"
        : "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5)
        : "r"(lo1), "r"(hi0), "r"(lo2), "r"(hi1), "r"(lo3), "r"(hi2), "r"(hi3));
    // V = U - H  (non-negative, < 2^165)
    uint32_t v0, v1, v2, v3, v4, v5;
    asm("sub.cc.u32 %0, %6, %12;\n\t"
        "subc.cc.u32 %1, %7, %13;\n\t"
        "subc.cc.u32 %2, %8, %14;\n\t"
        "subc.cc.u32 %3, %9, %15;\n\t"
        "subc.cc.u32 %4, %10, 0;\n\t"
        "subc.u32 %5, %11, 0;"
        : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5)
        : "r"(u0), "r"(u1), "r"(u2), "r"(u3), "r"(u4), "r"(u5), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
    // second fold: Vh = v5:v4 (< 2^37);  Z = (9*Vh << 32) - Vh  (>= 0, < 2^73)
    const uint64_t yy = (uint64_t)v4 * 9u;
    const uint32_t y0 = (uint32_t)yy;
    const uint32_t y1 = v5 * 9u + (uint32_t)(yy >> 32);
    uint32_t z0, z1, z2;
    asm("sub.cc.u32 %0, 0, %3;\n\t"
        "subc.cc.u32 %1, %4, %5;\n\t"
        "subc.u32 %2, %6, 0;"
        : "=r"(z0), "=r"(z1), "=r"(z2)
        : "r"(v4), "r"(y0), "r"(v5), "r"(y1));
    uint32_t w0, w1, w2, w3, co;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, 0;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(co)
        : "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(z0), "r"(z1), "r"(z2));
    // one conditional add of 2^128 - p covers both the carry-out and w >= p
    uint32_t t0, t1, t2, t3, k;
    asm("add.cc.u32 %0, %5, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 %1, %6, 8;\n\t"
        "addc.cc.u32 %2, %7, 0;\n\t"
        "addc.cc.u32 %3, %8, 0;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(k)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    fp out;
    bool sel = (co | k) != 0;
    out.v[0] = sel ? t0 : w0; out.v[1] = sel ? t1 : w1; out.v[2] = sel ? t2 : w2; out.v[3] = sel ? t3 : w3;
    return out;
}

// 4x4 limb product, r = a * b (256 bits)
GS_D void d_mul_wide(const fp& a, const fp& b, uint32_t (&r)[8]) {
    // even[k] = limb k, odd[k] = limb k+1
    uint32_t e0, e1, e2, e3, e4, e5, e6, e7;
    uint32_t o0, o1, o2, o3, o4, o5, o6;
    const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3];
    const uint32_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3];
    // row b0
    asm("mul.lo.u32 %0, %4, %6;\n\t mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %6;\n\t mul.hi.u32 %3, %5, %6;"
        : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3) : "r"(a0), "r"(a2), "r"(b0));
    asm("mul.lo.u32 %0, %4, %6;\n\t mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %6;\n\t mul.hi.u32 %3, %5, %6;"
        : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(a1), "r"(a3), "r"(b0));
    // row b1: a0,a2 -> odd[0..3] (+carry odd[4]);  a1,a3 -> even[2..5]
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "=r"(o4) : "r"(a0), "r"(a2), "r"(b1));
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, 0;\n\t madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(e2), "+r"(e3), "=r"(e4), "=r"(e5) : "r"(a1), "r"(a3), "r"(b1));
    // row b2: a0,a2 -> even[2..5] (+carry even[6]);  a1,a3 -> odd[2..5]
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "=r"(e6) : "r"(a0), "r"(a2), "r"(b2));
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "=r"(o5) : "r"(a1), "r"(a3), "r"(b2));
    // row b3: a0,a2 -> odd[2..5] (+carry odd[6]);  a1,a3 -> even[4..7]
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "=r"(o6) : "r"(a0), "r"(a2), "r"(b3));
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t madc.hi.u32 %3, %5, %6, 0;"
        : "+r"(e4), "+r"(e5), "+r"(e6), "=r"(e7) : "r"(a1), "r"(a3), "r"(b3));
    // r = even + (odd << 32)
    r[0] = e0;
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7),
          "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6));
}

GS_D fp d_mul(const fp& a, const fp& b) {
    uint32_t r[8];
    d_mul_wide(a, b, r);
    return d_reduce256(r);
}

// a * a with 10 limb products instead of 16: the six cross products a_i a_j (i < j) once, doubled by a one-bit funnel
// shift, plus the four squares -- 10 IMAD.WIDE and 7 SHF where d_mul issues 16 IMAD.WIDE (the FMA-heavy pipe binds a
// multiplication, fp128.cuh header / scripts/pipe_probe.cu)
GS_D fp d_sqr(const fp& a) {
    const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3];
    uint32_t c1, c2, c3, c4, c5, c6, c7;
    asm("mul.lo.u32 %0, %4, %5;\n\t mul.hi.u32 %1, %4, %5;\n\t"
        "mul.lo.u32 %2, %4, %6;\n\t mul.hi.u32 %3, %4, %6;"
        : "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4) : "r"(a0), "r"(a1), "r"(a3));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;"
        : "+r"(c2), "+r"(c3), "+r"(c4), "=r"(c5) : "r"(a0), "r"(a2));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;"
        : "+r"(c3), "+r"(c4), "+r"(c5), "=r"(c6) : "r"(a1), "r"(a2));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c4), "+r"(c5), "+r"(c6) : "r"(a1), "r"(a3));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, 0, 0;"
        : "+r"(c5), "+r"(c6), "=r"(c7) : "r"(a2), "r"(a3));
    // 2 * cross (a^2 < 2^256, so nothing falls off the top)
    uint32_t r[8];
    r[7] = __funnelshift_l(c6, c7, 1); r[6] = __funnelshift_l(c5, c6, 1); r[5] = __funnelshift_l(c4, c5, 1);
    r[4] = __funnelshift_l(c3, c4, 1); r[3] = __funnelshift_l(c2, c3, 1); r[2] = __funnelshift_l(c1, c2, 1);
    r[1] = c1 << 1;
    // + squares at limbs 0, 2, 4, 6
    asm("mul.lo.u32 %0, %8, %8;\n\t"
        "mad.hi.cc.u32 %1, %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %9, %2;\n\t madc.hi.cc.u32 %3, %9, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %10, %4;\n\t madc.hi.cc.u32 %5, %10, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %11, %6;\n\t madc.hi.u32 %7, %11, %11, %7;"
        : "=r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
    return d_reduce256(r);
}

GS_D fp d_inv(const fp& a) {
    const uint32_t e[4] = {0xFFFFFFFFu, 0xFFFFFFF6u, 0xFFFFFFFFu, 0xFFFFFFFFu};
    fp r = fp_one();
    for (int w = 3; w >= 0; --w) {
        for (int bit = 31; bit >= 0; --bit) {
            r = d_sqr(r);
            if ((e[w] >> bit) & 1u) r = d_mul(r, a);
        }
    }
    return r;
}
)

// dispatch: PTX on the device, u128 on the host ------------------------------------------------------
GS_HD fp fp_add(const fp& a, const fp& b) {
#ifdef __CUDA_ARCH__
    return d_add(a, b);
#else
    return fp_from_u128(h_add(fp_to_u128(a), fp_to_u128(b)));
#endif
}
GS_HD fp fp_sub(const fp& a, const fp& b) {
#ifdef __CUDA_ARCH__
    return d_sub(a, b);
#else
    return fp_from_u128(h_sub(fp_to_u128(a), fp_to_u128(b)));
#endif
}
GS_HD fp fp_mul(const fp& a, const fp& b) {
#ifdef __CUDA_ARCH__
    return d_mul(a, b);
#else
    return fp_from_u128(h_mul(fp_to_u128(a), fp_to_u128(b)));
#endif
}
GS_HD fp fp_neg(const fp& a) { return fp_sub(fp_zero(), a); }
GS_HD fp fp_sqr(const fp& a) {
#ifdef __CUDA_ARCH__
    return d_sqr(a);
#else
    return fp_mul(a, a);
#endif
}

// shared between host and device -----------------------------------------------------------------
GS_HD fp fp_pow(fp b, uint64_t e) {
    fp r = fp_one();
    while (e) {
        if (e & 1) r = fp_mul(r, b);
        b = fp_sqr(b);
        e >>= 1;
    }
    return r;
}

// a^(p-2); inv(0) = 0.  p-2 = 0xFFFFFFFF_FFFFFFFF_FFFFFFF6_FFFFFFFF
GS_HD fp fp_inv(const fp& a) {
    const uint32_t e[4] = {0xFFFFFFFFu, 0xFFFFFFF6u, 0xFFFFFFFFu, 0xFFFFFFFFu};
    fp r = fp_one();
    for (int w = 3; w >= 0; --w) {
        for (int bit = 31; bit >= 0; --bit) {
            r = fp_sqr(r);
            if ((e[w] >> bit) & 1u) r = fp_mul(r, a);
        }
    }
    return r;
}

}  // namespace gs
