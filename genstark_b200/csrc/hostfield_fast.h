// Latency-tuned a*b+k and a*b*c+k for the compiled transition function (hostjit.h).  Text only: it is compiled with
// -march=native on the machine that runs it, so BMI2 mulx can be used where present; each primitive is a few asm
// blocks so that no intermediate goes through memory (GCC 13 spills _addcarry_u64 chains), and the final
// "wrapped past 2^128" correction is a never-taken branch instead of three dependent instructions.
//   f_mul3_add: the first product stays half reduced (W = T_lo + T_hi C9, 165 bits); W * c (293 bits) is folded once
//   with 2^128 = C9, 2^192 = C9 2^64, 2^256 = C9^2 = 80 2^64 + D0.  A MiMC step (x^3 + k) is one call:
//   ~44 cycles of dependent latency instead of ~58 for two separately reduced multiplications and an addition.
// tests/test_trace_jit.py checks the compiled traces against the interpreter and the oracle.
#pragma once
namespace gs {
static const char GS_HOSTFIELD_FAST_SRC[] = R"GSFAST(
#if defined(__x86_64__) && defined(__BMI2__) && !defined(GS_JIT_PORTABLE)
// ---- latency-tuned primitives (x86-64 BMI2), each a few asm blocks so no value goes through memory
static const u64_t W_D0 = 0xFFFFFFEE00000001ull;     // C9^2 = 80 * 2^64 + W_D0  (= 2^256 mod p)

struct w256 { u64_t t0, t1, t2, t3; };
// a * b, 256 bits
static inline w256 f_wide(w128 a, w128 b) {
    u64_t t0, t1, t2, t3, x0, x1, x2, x3;
    asm("movq %[a0], %%rdx\n\t"
        "mulx %[b0], %[t0], %[t1]\n\t"
        "mulx %[b1], %[x0], %[t2]\n\t"
        "movq %[a1], %%rdx\n\t"
        "mulx %[b0], %[x1], %[x2]\n\t"
        "mulx %[b1], %[x3], %[t3]\n\t"
        "addq %[x0], %[t1]\n\t"
        "adcq %[x3], %[t2]\n\t"
        "adcq $0, %[t3]\n\t"
        "addq %[x1], %[t1]\n\t"
        "adcq %[x2], %[t2]\n\t"
        "adcq $0, %[t3]"
        : [t0] "=&r"(t0), [t1] "=&r"(t1), [t2] "=&r"(t2), [t3] "=&r"(t3), [x0] "=&r"(x0), [x1] "=&r"(x1), [x2] "=&r"(x2), [x3] "=&r"(x3)
        : [a0] "r"(a.lo), [a1] "r"(a.hi), [b0] "r"(b.lo), [b1] "r"(b.hi)
        : "rdx", "cc");
    return {t0, t1, t2, t3};
}
// (z0 + z1 2^64 + z2 2^128 + z3 2^192 + z4 2^256) + (k0 + k1 2^64)  ->  weakly reduced;  z4 < 2^40
static inline w128 f_fold5(u64_t z0, u64_t z1, u64_t z2, u64_t z3, u64_t z4, u64_t k0, u64_t k1) {
    u64_t e0, e1, f0, f0h;
    unsigned char cf;
    asm("movabsq $0x8FFFFFFFF, %%rdx\n\t"
        "mulx %[z2], %[e0], %[z2]\n\t"         // z2 := hi(z2 * C9)
        "mulx %[z3], %[e1], %[z3]\n\t"         // z3 := hi(z3 * C9) < 2^36: becomes limb 2
        "movabsq $0xFFFFFFEE00000001, %%rdx\n\t"
        "mulx %[z4], %[f0], %[f0h]\n\t"
        "leaq (%[z4],%[z4],4), %[z4]\n\t"
        "shlq $4, %[z4]\n\t"                   // z4 := 80 * z4
        "addq %[k0], %[z0]\n\t"                // (z0, z1, z3) += k
        "adcq %[k1], %[z1]\n\t"
        "adcq $0, %[z3]\n\t"
        "addq %[z4], %[f0h]\n\t"               // < 2^46, no carry
        "addq %[e0], %[z0]\n\t"
        "adcq %[z2], %[z1]\n\t"
        "adcq $0, %[z3]\n\t"
        "addq %[f0], %[z0]\n\t"
        "adcq %[f0h], %[z1]\n\t"
        "adcq $0, %[z3]\n\t"
        "addq %[e1], %[z1]\n\t"
        "adcq $0, %[z3]\n\t"
        "movabsq $0x8FFFFFFFF, %%rdx\n\t"
        "mulx %[z3], %[e0], %[e1]\n\t"         // limb 2 < 2^37
        "addq %[e0], %[z0]\n\t"
        "adcq %[e1], %[z1]"
        : [z0] "+&r"(z0), [z1] "+&r"(z1), [z2] "+&r"(z2), [z3] "+&r"(z3), [z4] "+&r"(z4),
          [e0] "=&r"(e0), [e1] "=&r"(e1), [f0] "=&r"(f0), [f0h] "=&r"(f0h), "=@ccc"(cf)
        : [k0] "rm"(k0), [k1] "rm"(k1)
        : "rdx");
    if (__builtin_expect(cf, 0)) {       // wrapped past 2^128 (probability ~2^-50): the remainder is tiny, + C9 cannot carry
        asm("addq %2, %0\n\tadcq $0, %1" : "+&r"(z0), "+&r"(z1) : "r"(W_C9) : "cc");
    }
    return {z0, z1};
}
// a * b + k
static inline w128 f_mul_add(w128 a, w128 b, w128 k) {
    const w256 t = f_wide(a, b);
    return f_fold5(t.t0, t.t1, t.t2, t.t3, 0, k.lo, k.hi);
}
// a * b * c + k with the first product only half reduced: W = T_lo + T_hi C9 (165 bits), then W * c (293 bits) folded once
static inline w128 f_mul3_add(w128 a, w128 b, w128 c3, w128 k) {
    const w256 t = f_wide(a, b);
    u64_t w0 = t.t0, w1 = t.t1, w2, q0, qh0, q1;
    asm("movabsq $0x8FFFFFFFF, %%rdx\n\t"
        "mulx %[t2], %[q0], %[qh0]\n\t"
        "mulx %[t3], %[q1], %[w2]\n\t"
        "addq %[q0], %[w0]\n\t"
        "adcq %[qh0], %[w1]\n\t"
        "adcq $0, %[w2]\n\t"
        "addq %[q1], %[w1]\n\t"
        "adcq $0, %[w2]"
        : [w0] "+&r"(w0), [w1] "+&r"(w1), [w2] "=&r"(w2), [q0] "=&r"(q0), [qh0] "=&r"(qh0), [q1] "=&r"(q1)
        : [t2] "r"(t.t2), [t3] "r"(t.t3)
        : "rdx", "cc");
    u64_t z0, z1, z2, z3, z4, x0, x1, x2;
    asm("movq %[c0], %%rdx\n\t"
        "mulx %[w0], %[z0], %[z1]\n\t"
        "mulx %[w1], %[x0], %[z2]\n\t"
        "mulx %[w2], %[x1], %[z3]\n\t"
        "movq %[c1], %%rdx\n\t"
        "addq %[x0], %[z1]\n\t"
        "adcq %[x1], %[z2]\n\t"
        "adcq $0, %[z3]\n\t"
        "mulx %[w0], %[x0], %[x1]\n\t"
        "mulx %[w1], %[x2], %[w0]\n\t"         // w0, w1 are dead from here on: reused as temporaries
        "mulx %[w2], %[w1], %[z4]\n\t"
        "addq %[x0], %[z1]\n\t"
        "adcq %[x1], %[z2]\n\t"
        "adcq %[w0], %[z3]\n\t"
        "adcq $0, %[z4]\n\t"
        "addq %[x2], %[z2]\n\t"
        "adcq %[w1], %[z3]\n\t"
        "adcq $0, %[z4]"
        : [z0] "=&r"(z0), [z1] "=&r"(z1), [z2] "=&r"(z2), [z3] "=&r"(z3), [z4] "=&r"(z4),
          [x0] "=&r"(x0), [x1] "=&r"(x1), [x2] "=&r"(x2), [w0] "+&r"(w0), [w1] "+&r"(w1)
        : [w2] "r"(w2), [c0] "rm"(c3.lo), [c1] "rm"(c3.hi)
        : "rdx", "cc");
    return f_fold5(z0, z1, z2, z3, z4, k.lo, k.hi);
}
#else
static inline w128 f_mul_add(w128 a, w128 b, w128 k) { return w_add(w_mul(a, b), k); }
static inline w128 f_mul3_add(w128 a, w128 b, w128 c3, w128 k) { return w_add(w_mul(w_mul(a, b), c3), k); }
#endif
)GSFAST";
}  // namespace gs
