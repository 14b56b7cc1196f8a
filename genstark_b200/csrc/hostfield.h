// Latency-optimised host arithmetic for the one sequential stage of prove(): execution-trace generation
// (lib/Stark.ts:97).  MiMC is a chain of 2^20 dependent cubings, so what matters is the latency of one
// modular multiplication, not throughput.  Values are kept "weakly reduced" (any 128-bit representative
// of the residue); only the copy stored into the trace is canonicalised, off the dependency chain.
// 2^128 = c9 = 9*2^32 - 1 (mod p): the high half of a product is folded by one 64x36-bit multiply.
#pragma once
// The arithmetic below exists twice from one spelling: compiled into the library (interpreter, verifier) and
// as text (GS_HOSTFIELD_SRC) that hostjit.h prepends to the C++ it generates for an AIR's transition function.
#include <cstdint>
#include <x86intrin.h>

#define GS_DUAL_SOURCE(name, ...) __VA_ARGS__ static const char name[] = #__VA_ARGS__;

namespace gs {

GS_DUAL_SOURCE(GS_HOSTFIELD_SRC,

typedef unsigned long long u64_t;
struct w128 { u64_t lo, hi; };
static const u64_t W_C9 = (9ull << 32) - 1;

static inline u64_t w_mul64(u64_t a, u64_t b, u64_t* hi) { unsigned __int128 p = (unsigned __int128)a * b; *hi = (u64_t)(p >> 64); return (u64_t)p; }

static inline w128 w_mul(w128 a, w128 b) {
    u64_t h00, h01, h10, h11;
    const u64_t l00 = w_mul64(a.lo, b.lo, &h00), l01 = w_mul64(a.lo, b.hi, &h01);
    const u64_t l10 = w_mul64(a.hi, b.lo, &h10), l11 = w_mul64(a.hi, b.hi, &h11);
    u64_t t0 = l00, t1, t2, t3; unsigned char c;
    c = _addcarry_u64(0, h00, l01, &t1);
    c = _addcarry_u64(c, h01, l11, &t2);
    _addcarry_u64(c, h11, 0, &t3);
    c = _addcarry_u64(0, t1, l10, &t1);
    c = _addcarry_u64(c, t2, h10, &t2);
    _addcarry_u64(c, t3, 0, &t3);
    u64_t qh0, qh1;
    const u64_t q0 = w_mul64(t2, W_C9, &qh0), q1 = w_mul64(t3, W_C9, &qh1);
    u64_t r0, r1, r2;
    c = _addcarry_u64(0, t0, q0, &r0);
    c = _addcarry_u64(c, t1, qh0, &r1);
    _addcarry_u64(c, qh1, 0, &r2);
    c = _addcarry_u64(0, r1, q1, &r1);
    _addcarry_u64(c, r2, 0, &r2);
    u64_t sh; const u64_t s = w_mul64(r2, W_C9, &sh);        // r2 < 2^38
    c = _addcarry_u64(0, r0, s, &r0);
    c = _addcarry_u64(c, r1, sh, &r1);
    const u64_t m = (u64_t)0 - (u64_t)c;                     // a carry here leaves a tiny value: one more c9 cannot carry
    c = _addcarry_u64(0, r0, m & W_C9, &r0);
    _addcarry_u64(c, r1, 0, &r1);
    return {r0, r1};
}
static inline w128 w_add(w128 a, w128 b) {
    u64_t r0, r1; unsigned char c = _addcarry_u64(0, a.lo, b.lo, &r0); c = _addcarry_u64(c, a.hi, b.hi, &r1);
    const u64_t m = (u64_t)0 - (u64_t)c; c = _addcarry_u64(0, r0, m & W_C9, &r0); _addcarry_u64(c, r1, 0, &r1);
    return {r0, r1};
}
static inline w128 w_sub(w128 a, w128 b) {
    u64_t r0, r1; unsigned char bw = _subborrow_u64(0, a.lo, b.lo, &r0); bw = _subborrow_u64(bw, a.hi, b.hi, &r1);
    // a borrow means the true value is r - 2^128 == r - c9; that can borrow once more (r < c9), never twice
    u64_t m = (u64_t)0 - (u64_t)bw; bw = _subborrow_u64(0, r0, m & W_C9, &r0); bw = _subborrow_u64(bw, r1, 0, &r1);
    m = (u64_t)0 - (u64_t)bw; bw = _subborrow_u64(0, r0, m & W_C9, &r0); _subborrow_u64(bw, r1, 0, &r1);
    return {r0, r1};
}
static inline unsigned __int128 w_canon(w128 x) {
    const unsigned __int128 p = (((unsigned __int128)0xFFFFFFFFFFFFFFFFull) << 64) | 0xFFFFFFF700000001ull;
    unsigned __int128 v = ((unsigned __int128)x.hi << 64) | x.lo;
    return v >= p ? v - p : v;
}
static inline w128 w_from(unsigned __int128 v) { return {(u64_t)v, (u64_t)(v >> 64)}; }
static inline w128 w_pow(w128 b, u64_t elo, u64_t ehi) {
    w128 r = {1, 0};
    while (elo | ehi) {
        if (elo & 1) r = w_mul(r, b);
        b = w_mul(b, b);
        elo = (elo >> 1) | (ehi << 63); ehi >>= 1;
    }
    return r;
}
static inline w128 w_inv(w128 a) { return w_pow(a, 0xFFFFFFF6FFFFFFFFull, 0xFFFFFFFFFFFFFFFFull); }
)

}  // namespace gs
