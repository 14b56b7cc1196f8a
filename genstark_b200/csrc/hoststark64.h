// Host-side Stark.prove / Stark.verify for prime fields of at most 64 bits (BASELINE config 1: the Foo demo of
// /root/reference/README.md:22-50 over p = 2^32 - 3*2^25 + 1).  The B200 kernels are written for the 16-byte element of the
// 128-bit STARK field; the reference itself has no optimised backend for other fields either and falls back to JS bigint
// arithmetic (lib/Stark.ts:41-43, README.md:118).  This is that fallback's counterpart: plain C++ on the host, reached through the
// same instantiate() / prove() / verify() surface, never used for the 128-bit field (gs_host_stark_* refuse it: that field has a
// GPU path and no CPU one).  Protocol: lib/Stark.ts:81-248 and lib/components/*.ts; the composition polynomial is evaluated
// directly on the evaluation domain as in compose.cuh (identical field elements).
// Element wire size = max(8, ceil(bits / 8)) bytes, little-endian (galois elementSize for small fields).
#pragma once
#include <algorithm>
#include "hostair.h"
#include "hostcrypto.h"
#include "verifier.h"

namespace gs {
namespace small {

typedef uint64_t u64;

struct Field {
    u64 p = 0; int es = 8;
    u64 add(u64 a, u64 b) const { const u128 s = (u128)a + b; return (u64)(s >= p ? s - p : s); }
    u64 sub(u64 a, u64 b) const { return a >= b ? a - b : a + (p - b); }
    u64 mul(u64 a, u64 b) const { return (u64)(((u128)a * b) % p); }
    u64 pow(u64 b, u128 e) const { u64 r = 1 % p; while (e) { if (e & 1) r = mul(r, b); b = mul(b, b); e >>= 1; } return r; }
    u64 inv(u64 a) const { return a == 0 ? 0 : pow(a, (u128)p - 2); }                   // inv(0) = 0
    u64 neg(u64 a) const { return a ? p - a : 0; }
    // galois getRootOfUnity: smallest i >= 2 whose i^((p-1)/order) has exact order `order`
    u64 root_of_unity(u64 order) const {
        for (u64 i = 2; i < p; ++i) {
            const u64 g = pow(i, (u128)(p - 1) / order);
            if (pow(g, order) == 1 && (order == 1 || pow(g, order / 2) != 1)) return g;
        }
        return 0;
    }
    // Miller-Rabin with the first twelve primes as bases: deterministic below 3.3 * 10^24, so for every 64-bit modulus.
    // A composite "modulus" would not only make the arithmetic meaningless, it has no element of the required order and the
    // generator search above would walk the whole range (found by fuzzing the AIR blob: a one-bit change of p32).
    bool is_prime() const {
        if (p < 2) return false;
        for (u64 q : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) if (p % q == 0) return p == q;
        u64 d = p - 1; int sft = 0; while ((d & 1) == 0) { d >>= 1; ++sft; }
        for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
            u64 x = pow(a % p, d);
            if (x == 1 || x == p - 1) continue;
            bool comp = true;
            for (int r = 1; r < sft && comp; ++r) { x = mul(x, x); if (x == p - 1) comp = false; }
            if (comp) return false;
        }
        return true;
    }
    u64 digest_mod(const uint8_t d[32]) const { u128 r = 0; for (int k = 0; k < 32; ++k) r = ((r << 8) | d[k]) % p; return (u64)r; }
    u64 prng_one(const uint8_t* seed, size_t n) const { uint8_t d[32]; sha256_bytes(seed, n, d); return digest_mod(d); }
    std::vector<u64> prng_many(const uint8_t* seed, size_t n, int count) const {
        uint8_t st[32]; sha256_bytes(seed, n, st);
        const Be256 state = be_from_digest(st);
        std::vector<u64> out(count);
        for (int i = 0; i < count; ++i) { uint8_t d[32]; sha256_of_be(be_add_small(state, (uint64_t)i), d); out[i] = digest_mod(d); }
        return out;
    }
    void put(uint8_t* dst, u64 v) const { for (int i = 0; i < es; ++i) dst[i] = i < 8 ? (uint8_t)(v >> (8 * i)) : 0; }
    bool get(const uint8_t* src, u64* v) const {
        u64 x = 0; for (int i = 0; i < 8 && i < es; ++i) x |= (u64)src[i] << (8 * i);
        for (int i = 8; i < es; ++i) if (src[i]) return false;
        *v = x; return x < p;
    }
};
typedef std::vector<u64> Vec;

static inline Vec power_series(const Field& F, u64 base, size_t n) { Vec o(n); u64 a = 1 % F.p; for (size_t i = 0; i < n; ++i) { o[i] = a; a = F.mul(a, base); } return o; }
// in-place radix-2 DFT over the powers of `root` (order n = v.size()), natural order in and out
static inline void fft(const Field& F, Vec& v, u64 root) {
    const size_t n = v.size(); if (n <= 1) return;
    int lg = 0; while (((size_t)1 << lg) < n) ++lg;
    for (size_t i = 0; i < n; ++i) { size_t j = 0; for (int b = 0; b < lg; ++b) j |= ((i >> b) & 1) << (lg - 1 - b); if (j > i) std::swap(v[i], v[j]); }
    const Vec tw = power_series(F, root, n / 2);
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, stride = n / len;
        for (size_t blk = 0; blk < n; blk += len)
            for (size_t i = 0; i < half; ++i) { const u64 t = F.mul(v[blk + i + half], tw[i * stride]), u = v[blk + i]; v[blk + i] = F.add(u, t); v[blk + i + half] = F.sub(u, t); }
    }
}
static inline Vec interpolate_roots(const Field& F, Vec values, u64 root) {          // values at root^i -> coefficients
    fft(F, values, F.inv(root));
    const u64 ninv = F.inv((u64)(values.size() % F.p));
    for (auto& x : values) x = F.mul(x, ninv);
    return values;
}
static inline Vec eval_at_roots(const Field& F, const Vec& poly, size_t n, u64 root) { Vec v(n, 0); std::copy(poly.begin(), poly.end(), v.begin()); fft(F, v, root); return v; }
static inline u64 eval_poly(const Field& F, const Vec& poly, u64 x) { u64 acc = 0; for (size_t k = poly.size(); k-- > 0;) acc = F.add(F.mul(acc, x), poly[k]); return acc; }
static inline Vec batch_inverse(const Field& F, const Vec& v) {                          // zeros stay zero
    const size_t n = v.size(); Vec pre(n), out(n, 0); u64 acc = 1;
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (v[i]) acc = F.mul(acc, v[i]); }
    u64 inv = F.inv(acc);
    for (size_t i = n; i-- > 0;) { if (!v[i]) continue; out[i] = F.mul(inv, pre[i]); inv = F.mul(inv, v[i]); }
    return out;
}
// Lagrange interpolation, coefficients low -> high (BoundaryConstraints.ts:42, LowDegreeProver.ts:243)
static inline Vec lagrange(const Field& F, const Vec& xs, const Vec& ys) {
    const size_t n = xs.size();
    Vec root(n + 1, 0); root[0] = 1;
    for (size_t i = 0; i < n; ++i) { for (size_t k = i + 1; k > 0; --k) root[k] = F.sub(root[k - 1], F.mul(root[k], xs[i])); root[0] = F.sub(0, F.mul(root[0], xs[i])); }
    Vec out(n, 0), num(n);
    for (size_t i = 0; i < n; ++i) {
        u64 acc = 0;
        for (size_t k = n; k > 0; --k) { acc = F.add(root[k], F.mul(acc, xs[i])); num[k - 1] = acc; }
        const u64 f = F.mul(ys[i], F.inv(eval_poly(F, num, xs[i])));
        for (size_t k = 0; k < n; ++k) out[k] = F.add(out[k], F.mul(num[k], f));
    }
    return out;
}

// the AIR as the blob carries it (air.py: pack_air), any modulus below 2^64
struct Air {
    Field F; int R = 0, K = 0, log_t = 0, log_e = 0;
    struct Static { int kind; Vec values; };             // 0 cycle, 1 secret input, 2 public input
    std::vector<Static> statics; std::vector<int> degrees;
    struct Prog { std::vector<std::array<uint32_t, 4>> ins; Vec consts; std::vector<u128> wide; int n_slots = 0, n_out = 0; } transition, evaluation;
    int n_secret = 0, n_public = 0;
};
static inline std::string parse_air64(const uint8_t* blob, size_t len, Air& A, int* code) {
    *code = GS_E_ARG;
    size_t off = 0; bool ok = true;
    auto u32 = [&]() -> uint32_t { if (off + 4 > len) { ok = false; return 0; } uint32_t v; memcpy(&v, blob + off, 4); off += 4; return v; };
    auto wide = [&]() -> u128 { if (off + 16 > len) { ok = false; return 0; } u128 v = 0; for (int i = 15; i >= 0; --i) v = (v << 8) | blob[off + i]; off += 16; return v; };
    if (u32() != 0x52494147u) return "bad AIR blob magic";
    const u128 mod = wide();
    if (!ok || mod < 3 || (mod >> 64) != 0) { *code = GS_E_UNSUPPORTED; return "the host path takes prime fields of at most 64 bits"; }
    A.F.p = (u64)mod; int bits = 0; while (bits < 64 && (mod >> bits)) ++bits;
    A.F.es = std::max(8, (bits + 7) / 8);
    if (!A.F.is_prime()) { *code = GS_E_UNSUPPORTED; return "the modulus is not prime"; }
    A.R = (int)u32(); A.K = (int)u32(); A.log_t = (int)u32(); A.log_e = (int)u32();
    const uint32_t ns = u32();
    if (!ok || A.R < 1 || A.R > GS_MAX_COLS || A.K < 1 || A.K > GS_MAX_CONSTRAINTS || ns > GS_MAX_COLS) { *code = GS_E_UNSUPPORTED; return "AIR shape out of range"; }
    if (A.log_t < 2 || A.log_t > 20 || A.log_e < 1 || A.log_e > 5) return "trace length 4..2^20 and extension factor 2..32 required";
    if (((A.F.p - 1) >> (A.log_t + A.log_e)) << (A.log_t + A.log_e) != A.F.p - 1) { *code = GS_E_UNSUPPORTED; return "the field has no root of unity of the order of the evaluation domain"; }
    A.statics.resize(ns);
    for (auto& s : A.statics) {
        s.kind = (int)u32(); const uint32_t l = u32();
        if (!ok || l > (1u << 20) || s.kind < 0 || s.kind > 2 || (s.kind != 0 && l != 0)) return "bad static register";
        s.values.resize(l);
        for (auto& v : s.values) { const u128 w = wide(); if (w >= A.F.p) ok = false; v = (u64)w; }
        if (s.kind == 0 && (l == 0 || (l & (l - 1)) || l > (1u << A.log_t))) return "cycle length must be a power of two <= steps";
        if (s.kind == 1) A.n_secret++;
        if (s.kind == 2) A.n_public++;
    }
    A.degrees.resize(A.K);
    for (auto& d : A.degrees) d = (int)u32();
    for (Air::Prog* pr : {&A.transition, &A.evaluation}) {
        const uint32_t ni = u32(), nc = u32(); pr->n_slots = (int)u32(); pr->n_out = (int)u32();
        if (!ok || ni > (1u << 20) || nc > (1u << 20)) return "bad AIR program";
        pr->ins.resize(ni);
        for (auto& i : pr->ins) for (int k = 0; k < 4; ++k) i[k] = u32();
        pr->consts.resize(nc); pr->wide.resize(nc);
        for (uint32_t i = 0; i < nc; ++i) { pr->wide[i] = wide(); pr->consts[i] = (u64)(pr->wide[i] % A.F.p); }
    }
    if (!ok) return "truncated AIR blob";
    if (A.transition.n_out != A.R || A.evaluation.n_out != A.K) return "program outputs do not match the register / constraint counts";
    if (const char* bad = validate_instrs(A.transition.ins, A.transition.consts.size(), A.transition.n_slots, A.transition.n_out, A.R, (int)ns, true)) return std::string("transition ") + bad;
    if (const char* bad = validate_instrs(A.evaluation.ins, A.evaluation.consts.size(), A.evaluation.n_slots, A.evaluation.n_out, A.R, (int)ns, false)) return std::string("evaluation ") + bad;
    for (int d : A.degrees) if (d < 0 || d > 256) return "constraint degree out of range";
    *code = GS_OK;
    return "";
}
static inline bool run(const Field& F, const Air::Prog& pr, const u64* cur, const u64* nxt, const u64* st, Vec& slots, u64* out) {
    slots.assign(pr.n_slots + 1, 0);
    for (auto& i : pr.ins) {
        const uint32_t op = i[0], d = i[1], a = i[2], b = i[3];
        if (op != OP_OUT && d > (uint32_t)pr.n_slots) return false;
        switch (op) {
            case OP_CONST: if (a >= pr.consts.size()) return false; slots[d] = pr.consts[a]; break;
            case OP_CUR: slots[d] = cur[a]; break;
            case OP_NEXT: if (!nxt) return false; slots[d] = nxt[a]; break;
            case OP_STATIC: slots[d] = st[a]; break;
            case OP_ADD: slots[d] = F.add(slots[a], slots[b]); break;
            case OP_SUB: slots[d] = F.sub(slots[a], slots[b]); break;
            case OP_MUL: slots[d] = F.mul(slots[a], slots[b]); break;
            case OP_NEG: slots[d] = F.neg(slots[a]); break;
            case OP_INV: slots[d] = F.inv(slots[a]); break;
            case OP_EXP: if (b >= pr.wide.size()) return false; slots[d] = F.pow(slots[a], pr.wide[b]); break;
            case OP_OUT: if ((int)d >= pr.n_out) return false; out[d] = slots[a]; break;
            default: return false;
        }
    }
    return true;
}

struct Assertion64 { uint32_t reg, step; u64 value; };
typedef std::array<uint8_t, 32> Dg;

struct Tree { std::vector<Dg> nodes; size_t n = 0; const Dg& root() const { return nodes[1]; } };
static inline Tree make_tree(const HostHash& H, const std::vector<Dg>& leaves) {
    Tree t; t.n = leaves.size(); t.nodes.assign(2 * t.n, Dg{});
    std::copy(leaves.begin(), leaves.end(), t.nodes.begin() + t.n);
    for (size_t i = t.n - 1; i >= 1; --i) t.nodes[i] = H.merge(t.nodes[2 * i], t.nodes[2 * i + 1]);
    return t;
}
static inline int tree_proof(const Tree& t, const std::vector<uint32_t>& idx, BatchProof& bp, std::string& err) {
    if (merkle_prove_plan(idx, t.n, bp, err) != 0) return -1;
    bp.nodes.assign(bp.node_ids.size(), {});
    for (size_t c = 0; c < bp.node_ids.size(); ++c) for (uint32_t id : bp.node_ids[c]) bp.nodes[c].push_back(t.nodes[id]);
    return 0;
}
static inline std::vector<uint32_t> unique_in_order(const std::vector<uint32_t>& v) {
    std::vector<uint32_t> out; std::map<uint32_t, bool> seen;
    for (uint32_t x : v) if (!seen.count(x)) { seen[x] = true; out.push_back(x); }
    return out;
}

// Fiat-Shamir bookkeeping shared by prover and verifier (CompositionPolynomial.ts:29-61, LinearCombination.ts:21-34)
struct Plan {
    long long T, N, E; int log_n;
    long long comb_degree, comp_degree;
    std::vector<long long> group_deg; std::vector<std::vector<int>> group_idx;
    int d_count = 0, b_count = 0, n_lc = 0, lc_total = 0;
    std::vector<int> adj_idx;                 // per constraint: index of its second coefficient, -1 if none
    std::vector<long long> incr;              // per constraint: combination degree - its group's degree
    std::vector<uint32_t> b_regs; std::vector<Vec> b_xs, b_ys;
    u64 w_n = 0;
};
static inline std::string make_plan(const Air& A, const std::vector<Assertion64>& as, Plan& P) {
    const Field& F = A.F;
    P.T = 1ll << A.log_t; P.E = 1ll << A.log_e; P.N = P.T * P.E; P.log_n = A.log_t + A.log_e;
    P.w_n = F.root_of_unity((u64)P.N);
    if (!P.w_n) return "no root of unity";
    int max_deg = 1; for (int d : A.degrees) max_deg = std::max(max_deg, d);
    int lc = 0; while ((1 << lc) < max_deg) ++lc;
    P.comb_degree = P.T << lc; P.comp_degree = std::max(P.comb_degree - P.T, P.T);
    for (int k = 0; k < A.K; ++k) {
        const long long dg = (long long)A.degrees[k] * P.T;
        size_t g = 0; for (; g < P.group_deg.size(); ++g) if (P.group_deg[g] == dg) break;
        if (g == P.group_deg.size()) { P.group_deg.push_back(dg); P.group_idx.emplace_back(); }
        P.group_idx[g].push_back(k);
    }
    P.d_count = A.K; P.adj_idx.assign(A.K, -1); P.incr.assign(A.K, 0);
    int next = A.K;
    for (size_t g = 0; g < P.group_deg.size(); ++g) {
        if (P.group_deg[g] == P.comb_degree) continue;
        for (int k : P.group_idx[g]) { P.adj_idx[k] = next++; P.incr[k] = P.comb_degree - P.group_deg[g]; }
    }
    P.d_count = next;
    for (auto& a : as) {
        if ((int)a.reg >= A.R) return "Invalid assertion: register " + std::to_string(a.reg) + " is outside of register bank";
        if (a.step >= (uint64_t)P.T) return "Invalid assertion: step " + std::to_string(a.step) + " is outside of execution trace";
        size_t b = 0; for (; b < P.b_regs.size(); ++b) if (P.b_regs[b] == a.reg) break;
        if (b == P.b_regs.size()) { P.b_regs.push_back(a.reg); P.b_xs.emplace_back(); P.b_ys.emplace_back(); }
        const u64 x = F.pow(P.w_n, (u128)a.step * (u128)P.E);
        for (u64 seen : P.b_xs[b]) if (seen == x) return "repeated assertion for register " + std::to_string(a.reg);
        P.b_xs[b].push_back(x); P.b_ys[b].push_back(a.value);
    }
    P.b_count = (int)P.b_regs.size() * (P.comp_degree > P.T ? 2 : 1);
    P.n_lc = A.R + A.n_secret;
    P.lc_total = P.n_lc * (P.comp_degree > P.T ? 2 : 1);
    return "";
}
static inline Vec zpoly(const Field& F, const Vec& xs) {               // prod (x - X_k), low -> high
    Vec z{1};
    for (u64 x : xs) { Vec n(z.size() + 1, 0); for (size_t k = 0; k < z.size(); ++k) { n[k + 1] = F.add(n[k + 1], z[k]); n[k] = F.sub(n[k], F.mul(z[k], x)); } z.swap(n); }
    return z;
}
static inline std::vector<uint32_t> aug4(const std::vector<uint32_t>& p, long long column_length) {
    std::vector<uint32_t> m(p.size()); const uint32_t row = (uint32_t)(column_length >> 2);
    for (size_t i = 0; i < p.size(); ++i) m[i] = p[i] % row;
    return unique_in_order(m);
}
// verifyRemainder (LowDegreeProver.ts:223-252)
static inline std::string check_remainder(const Field& F, const Vec& rem, long long max_deg_p1, u64 rou, long long E) {
    const long long L = (long long)rem.size();
    std::vector<long long> pos; for (long long i = 0; i < L; ++i) if (i % E) pos.push_back(i);
    if (max_deg_p1 > (long long)pos.size()) return "remainder too short for degree " + std::to_string(max_deg_p1);
    const Vec dom = power_series(F, rou, (size_t)L);
    Vec xs(max_deg_p1), ys(max_deg_p1);
    for (long long i = 0; i < max_deg_p1; ++i) { xs[i] = dom[pos[i]]; ys[i] = rem[pos[i]]; }
    const Vec poly = lagrange(F, xs, ys);
    for (size_t i = (size_t)max_deg_p1; i < pos.size(); ++i)
        if (eval_poly(F, poly, dom[pos[i]]) != rem[pos[i]]) return "Remainder is not a valid degree " + std::to_string(max_deg_p1 - 1) + " polynomial";
    return "";
}
// one FRI fold (LowDegreeProver.ts:190-195): the cubic through (x_i iota^j, v[i + j L/4]) at x*, as fri.cuh computes it
static inline Vec fri_fold(const Field& F, const Vec& v, u64 w_layer /* root of order L */, u64 xs) {
    const size_t L = v.size(), Q = L / 4;
    const u64 iota_inv = F.inv(F.pow(w_layer, (u128)Q)), quarter = F.inv(4 % F.p), w_inv = F.inv(w_layer);
    Vec out(Q); u64 xinv = 1;
    for (size_t i = 0; i < Q; ++i) {
        const u64 y0 = v[i], y1 = v[i + Q], y2 = v[i + 2 * Q], y3 = v[i + 3 * Q];
        const u64 t = F.mul(xs, xinv);
        const u64 s02 = F.add(y0, y2), d02 = F.sub(y0, y2), s13 = F.add(y1, y3), d13 = F.mul(F.sub(y1, y3), iota_inv);
        const u64 c0 = F.add(s02, s13), c2 = F.sub(s02, s13), c1 = F.add(d02, d13), c3 = F.sub(d02, d13);
        u64 acc = F.add(F.mul(c3, t), c2); acc = F.add(F.mul(acc, t), c1); acc = F.add(F.mul(acc, t), c0);
        out[i] = F.mul(acc, quarter);
        xinv = F.mul(xinv, w_inv);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ prove
static inline std::string prove(const Air& A, int hash_alg, int exe_q, int fri_q, const std::vector<Assertion64>& as, const Vec& init,
                                const std::vector<Vec>& input_traces, const uint8_t* shapes, size_t shapes_len, std::vector<uint8_t>& out) {
    const Field& F = A.F; const HostHash H{hash_alg};
    if (as.empty()) return "At least one assertion must be provided";
    Plan P; { const std::string e = make_plan(A, as, P); if (!e.empty()) return "Failed to generate the execution trace: " + e; }
    const long long T = P.T, N = P.N, E = P.E;
    if (N < 128) return "Low degree proof failed: Invalid array length";              // getComponentCount quirk (LowDegreeProver.ts:287-291)
    const int R = A.R, K = A.K, n_static = (int)A.statics.size();
    if ((int)init.size() != R) return "Failed to generate the execution trace: initial state has the wrong width";
    if ((int)input_traces.size() != A.n_secret + A.n_public) return "input register traces required";
    // static registers over the trace
    std::vector<const Vec*> in_of(n_static, nullptr); { int ii = 0; for (int k = 0; k < n_static; ++k) if (A.statics[k].kind != 0) in_of[k] = &input_traces[ii++]; }
    auto static_at_step = [&](long long s, u64* sv) { for (int k = 0; k < n_static; ++k) sv[k] = in_of[k] ? (*in_of[k])[s] : A.statics[k].values[s & (A.statics[k].values.size() - 1)]; };
    // 1-2 execution trace
    std::vector<Vec> trace(R, Vec(T));
    { Vec st = init, nx(R), slots; std::vector<u64> sv(n_static + 1);
      for (long long s = 0; s < T; ++s) {
          for (int r = 0; r < R; ++r) trace[r][s] = st[r];
          if (s + 1 < T) { static_at_step(s, sv.data()); if (!run(F, A.transition, st.data(), nullptr, sv.data(), slots, nx.data())) return "Failed to generate the execution trace: bad transition program"; st = nx; }
      } }
    for (auto& a : as) if (trace[a.reg][a.step] != a.value)
        return "Failed to generate the execution trace: Assertion at step " + std::to_string(a.step) + ", register " + std::to_string(a.reg) + " conflicts with execution trace";
    // 3 P(x), low-degree extension
    const u64 w_t = F.pow(P.w_n, (u128)E);
    std::vector<Vec> pe(R);
    for (int r = 0; r < R; ++r) pe[r] = eval_at_roots(F, interpolate_roots(F, trace[r], w_t), (size_t)N, P.w_n);
    // static registers over the evaluation domain: cycles of length L repeat with period L * E
    std::vector<Vec> st_e(n_static); std::vector<Vec> secret_e;
    for (int k = 0; k < n_static; ++k) {
        if (A.statics[k].kind == 0) {
            const size_t L = A.statics[k].values.size();
            const u64 g = F.pow(P.w_n, (u128)(N / (long long)L));
            st_e[k] = eval_at_roots(F, interpolate_roots(F, A.statics[k].values, g), L * (size_t)E, F.pow(P.w_n, (u128)(T / (long long)L)));
        } else {
            st_e[k] = eval_at_roots(F, interpolate_roots(F, *in_of[k], w_t), (size_t)N, P.w_n);
            if (A.statics[k].kind == 1) secret_e.push_back(st_e[k]);
        }
    }
    // 4 leaves: H(P_0[i] || ... || S_0[i] || ...), tree
    std::vector<const Vec*> e_cols; for (auto& v : pe) e_cols.push_back(&v); for (auto& v : secret_e) e_cols.push_back(&v);
    const size_t leaf_bytes = e_cols.size() * F.es;
    auto leaf_value = [&](uint32_t i) { std::vector<uint8_t> b(leaf_bytes); for (size_t c = 0; c < e_cols.size(); ++c) F.put(b.data() + c * F.es, (*e_cols[c])[i]); return b; };
    std::vector<Dg> leaves((size_t)N);
    for (long long i = 0; i < N; ++i) { const auto b = leaf_value((uint32_t)i); leaves[i] = H.digest(b.data(), b.size()); }
    const Tree e_tree = make_tree(H, leaves);
    // 5 coefficients (one draw covers composition and linear combination: LinearCombination.ts:58-59)
    const Vec coef = F.prng_many(e_tree.root().data(), 32, P.d_count + P.b_count + P.lc_total);
    const long long delta = P.comp_degree - T;
    const Vec dom = power_series(F, P.w_n, (size_t)N);
    // transition part: D = (x - x_last) / (x^T - 1) * sum_k (d_k + d'_k x^incr_k) q_k
    Vec C((size_t)N, 0);
    { Vec num((size_t)E); const u64 w_e = F.pow(P.w_n, (u128)T); u64 a = 1; for (long long j = 0; j < E; ++j) { num[j] = F.sub(a, 1); a = F.mul(a, w_e); }
      const Vec inv_num = batch_inverse(F, num);
      const u64 x_last = F.pow(P.w_n, (u128)(T - 1) * (u128)E);
      Vec cur(R), nxt(R), q(K), slots; std::vector<u64> sv(n_static + 1);
      for (long long i = 0; i < N; ++i) {
          for (int r = 0; r < R; ++r) { cur[r] = pe[r][i]; nxt[r] = pe[r][(i + E) % N]; }
          for (int k = 0; k < n_static; ++k) sv[k] = st_e[k][(size_t)i % st_e[k].size()];
          if (!run(F, A.evaluation, cur.data(), nxt.data(), sv.data(), slots, q.data())) return "Failed to evaluate transition constraints: bad evaluation program";
          if (i % E == 0 && i / E < T - 1) for (int k = 0; k < K; ++k) if (q[k] != 0)
              return "Failed to evaluate transition constraints: Constraint " + std::to_string(k) + " didn't evaluate to 0 at step " + std::to_string(i / E);
          u64 acc = 0;
          for (int k = 0; k < K; ++k) {
              u64 c = coef[k];
              if (P.adj_idx[k] >= 0) c = F.add(c, F.mul(coef[P.adj_idx[k]], F.pow(dom[i], (u128)P.incr[k])));
              acc = F.add(acc, F.mul(c, q[k]));
          }
          C[i] = F.mul(F.mul(acc, F.sub(dom[i], x_last)), inv_num[i % E]);
      } }
    // boundary part
    const int nB = (int)P.b_regs.size();
    for (int b = 0; b < nB; ++b) {
        const Vec ip = lagrange(F, P.b_xs[b], P.b_ys[b]), zp = zpoly(F, P.b_xs[b]);
        Vec z((size_t)N); for (long long i = 0; i < N; ++i) z[i] = eval_poly(F, zp, dom[i]);
        const Vec zi = batch_inverse(F, z);
        for (long long i = 0; i < N; ++i) {
            const u64 bv = F.mul(F.sub(pe[P.b_regs[b]][i], eval_poly(F, ip, dom[i])), zi[i]);
            u64 c = coef[P.d_count + b];
            if (delta > 0) c = F.add(c, F.mul(coef[P.d_count + nB + b], F.pow(dom[i], (u128)delta)));
            C[i] = F.add(C[i], F.mul(c, bv));
        }
    }
    // 6 linear combination
    Vec Lv = C;
    for (int j = 0; j < P.n_lc; ++j)
        for (long long i = 0; i < N; ++i) {
            u64 c = coef[P.d_count + P.b_count + j];
            if (delta > 0) c = F.add(c, F.mul(coef[P.d_count + P.b_count + P.n_lc + j], F.pow(dom[i], (u128)delta)));
            Lv[i] = F.add(Lv[i], F.mul(c, (*e_cols[j])[i]));
        }
    // 7 FRI layers
    struct Layer { Vec v; Tree tree; };
    std::vector<Layer> layers;
    auto commit_layer = [&](Vec v) {
        const size_t Q = v.size() / 4; std::vector<Dg> rows(Q); std::vector<uint8_t> buf(4 * F.es);
        for (size_t i = 0; i < Q; ++i) { for (int j = 0; j < 4; ++j) F.put(buf.data() + j * F.es, v[i + j * Q]); rows[i] = H.digest(buf.data(), buf.size()); }
        Layer ly; ly.tree = make_tree(H, rows); ly.v.swap(v); layers.push_back(std::move(ly));
    };
    commit_layer(Lv);
    for (int depth = 0; (long long)layers.back().v.size() > 256; ++depth) {
        const u64 xs = F.prng_one(layers.back().tree.root().data(), 32);
        commit_layer(fri_fold(F, layers.back().v, F.pow(P.w_n, (u128)1 << (2 * depth)), xs));
    }
    const int n_layers = (int)layers.size(), last = n_layers - 1;
    { long long md = P.comp_degree; for (int d = 0; d < last; ++d) md /= 4;
      const std::string e = check_remainder(F, layers[last].v, md, F.pow(P.w_n, (u128)1 << (2 * last)), E);
      if (!e.empty()) return "Low degree proof failed: " + e; }
    // queries
    std::string err;
    auto rows_of = [&](const Layer& ly, const std::vector<uint32_t>& idx) {
        std::vector<std::vector<uint8_t>> vals; const size_t Q = ly.v.size() / 4;
        for (uint32_t i : idx) { std::vector<uint8_t> b(4 * F.es); for (int j = 0; j < 4; ++j) F.put(b.data() + j * F.es, ly.v[i + j * Q]); vals.push_back(b); }
        return vals;
    };
    std::vector<uint32_t> exe_pos;
    if (pseudorandom_indexes(layers[0].tree.root().data(), (int)std::min<long long>(exe_q, N - N / E), (uint64_t)N, (uint64_t)E, exe_pos, err) != 0) return "Low degree proof failed: " + err;
    BatchProof lc_proof, ev_proof;
    { const auto idx = aug4(exe_pos, N); if (tree_proof(layers[0].tree, idx, lc_proof, err) != 0) return "Low degree proof failed: " + err; lc_proof.values = rows_of(layers[0], idx); }
    { std::vector<uint32_t> m; for (uint32_t p : exe_pos) { m.push_back(p); m.push_back((uint32_t)((p + E) % N)); }
      const auto idx = unique_in_order(m); if (tree_proof(e_tree, idx, ev_proof, err) != 0) return err;
      for (uint32_t i : idx) ev_proof.values.push_back(leaf_value(i)); }
    struct Comp { BatchProof column, poly; };
    std::vector<Comp> comps(last);
    for (int d = 1; d < n_layers; ++d) {
        const Layer& pl = layers[d - 1]; const Layer& cl = layers[d];
        std::vector<uint32_t> pos;
        if (pseudorandom_indexes(cl.tree.root().data(), fri_q, (uint64_t)cl.v.size(), (uint64_t)E, pos, err) != 0) return "Low degree proof failed: " + err;
        const auto aug = aug4(pos, (long long)cl.v.size());
        if (tree_proof(cl.tree, aug, comps[d - 1].column, err) != 0 || tree_proof(pl.tree, pos, comps[d - 1].poly, err) != 0) return "Low degree proof failed: " + err;
        comps[d - 1].column.values = rows_of(cl, aug); comps[d - 1].poly.values = rows_of(pl, pos);
    }
    // serialize (Serializer.ts:35-79)
    out.clear();
    const size_t ld_leaf = 4 * F.es;
    for (const BatchProof* p : {&ev_proof, &lc_proof}) if (check_merkle_proof_limits(*p, err) != 0) return err;
    out.insert(out.end(), e_tree.root().begin(), e_tree.root().end());
    write_merkle_proof(out, ev_proof, leaf_bytes);
    out.insert(out.end(), layers[0].tree.root().begin(), layers[0].tree.root().end());
    write_merkle_proof(out, lc_proof, ld_leaf);
    out.push_back((uint8_t)comps.size());
    for (int d = 0; d < last; ++d) {
        if (check_merkle_proof_limits(comps[d].column, err) != 0 || check_merkle_proof_limits(comps[d].poly, err) != 0) return err;
        out.insert(out.end(), layers[d + 1].tree.root().begin(), layers[d + 1].tree.root().end());
        write_merkle_proof(out, comps[d].column, ld_leaf);
        write_merkle_proof(out, comps[d].poly, ld_leaf);
    }
    const Vec& rem = layers[last].v;
    out.push_back((uint8_t)(rem.size() == 256 ? 0 : rem.size()));
    for (u64 v : rem) { std::vector<uint8_t> b(F.es); F.put(b.data(), v); out.insert(out.end(), b.begin(), b.end()); }
    if (shapes && shapes_len) out.insert(out.end(), shapes, shapes + shapes_len); else out.push_back(0);
    return "";
}

// ------------------------------------------------------------------------------------------------ verify
static inline std::string verify(const Air& A, int hash_alg, int exe_q, int fri_q, const std::vector<Assertion64>& as,
                                 const uint8_t* proof, size_t proof_len, const std::vector<Vec>& public_traces) {
    const Field& F = A.F; const HostHash H{hash_alg};
    if (as.empty()) return "At least one assertion must be provided";
    Plan P; { const std::string e = make_plan(A, as, P); if (!e.empty()) return e; }
    const long long T = P.T, N = P.N, E = P.E;
    const int R = A.R, n_static = (int)A.statics.size();
    if ((int)public_traces.size() != A.n_public) return "public input traces required";
    const size_t leaf_bytes = (size_t)(R + A.n_secret) * F.es, ld_leaf = 4 * F.es;
    ProofReader rd{proof, proof_len};
    Digest e_root; rd.bytes(e_root.data(), 32);
    ParsedBatch ev; if (!rd.batch(ev, leaf_bytes)) return "malformed proof";
    Digest lc_root; rd.bytes(lc_root.data(), 32);
    ParsedBatch lcp; if (!rd.batch(lcp, ld_leaf)) return "malformed proof";
    const int n_comp = rd.u8();
    { int want = 0; for (long long L = N; L > 256; L >>= 2) ++want; if (n_comp != want) return "Verification of low degree failed: malformed proof"; }
    struct Comp { Digest root; ParsedBatch column, poly; };
    std::vector<Comp> comps(n_comp);
    for (auto& c : comps) { rd.bytes(c.root.data(), 32); if (!rd.batch(c.column, ld_leaf) || !rd.batch(c.poly, ld_leaf)) return "malformed proof"; }
    size_t rl = rd.u8(); if (rl == 0) rl = 256;
    Vec rem(rl); for (auto& v : rem) { std::vector<uint8_t> b(F.es); rd.bytes(b.data(), F.es); if (!F.get(b.data(), &v)) return "malformed proof (non-canonical remainder)"; }
    { const int n_shapes = rd.u8();                              // input shapes (Serializer.ts:66-76): parsed, and the proof must end there
      for (int i = 0; i < n_shapes && rd.ok; ++i) { const int rank = rd.u8(); for (int k = 0; k < rank && rd.ok; ++k) { uint32_t lv; rd.bytes(&lv, 4); } } }
    if (!rd.ok || rd.off != proof_len) return "malformed proof";
    // coefficients, positions
    const Vec coef = F.prng_many(e_root.data(), 32, P.d_count + P.b_count + P.lc_total);
    const long long delta = P.comp_degree - T;
    std::string err; std::vector<uint32_t> exe_pos;
    if (pseudorandom_indexes(lc_root.data(), (int)std::min<long long>(exe_q, N - N / E), (uint64_t)N, (uint64_t)E, exe_pos, err) != 0) return err;
    std::vector<uint32_t> m; for (uint32_t p : exe_pos) { m.push_back(p); m.push_back((uint32_t)((p + E) % N)); }
    const auto augmented = unique_in_order(m);
    if (ev.values.size() != augmented.size()) return "Verification of evaluation Merkle proof failed";
    std::map<uint32_t, Vec> pv, sv_map;
    std::vector<Digest> hashed;
    for (size_t i = 0; i < augmented.size(); ++i) {
        Vec p(R), s(A.n_secret);
        for (int r = 0; r < R; ++r) if (!F.get(ev.values[i].data() + r * F.es, &p[r])) return "malformed proof (non-canonical value)";
        for (int r = 0; r < A.n_secret; ++r) if (!F.get(ev.values[i].data() + (R + r) * F.es, &s[r])) return "malformed proof (non-canonical value)";
        pv[augmented[i]] = p; sv_map[augmented[i]] = s;
        hashed.push_back(H.digest(ev.values[i].data(), ev.values[i].size()));
    }
    if (!verify_batch(e_root, augmented, hashed, ev.nodes, ev.depth, H)) return "Verification of evaluation Merkle proof failed";
    // constraint evaluation at the queried points
    std::vector<Vec> cyc_poly(n_static), pub_poly(n_static);
    { int ip = 0; const u64 w_t = F.pow(P.w_n, (u128)E);
      for (int k = 0; k < n_static; ++k) {
          if (A.statics[k].kind == 0) { const size_t L = A.statics[k].values.size(); cyc_poly[k] = interpolate_roots(F, A.statics[k].values, F.pow(P.w_n, (u128)(N / (long long)L))); }
          else if (A.statics[k].kind == 2) { if ((long long)public_traces[ip].size() != T) return "public input trace of the wrong length"; pub_poly[k] = interpolate_roots(F, public_traces[ip++], w_t); }
      } }
    std::vector<Vec> b_ip, b_zp; for (size_t b = 0; b < P.b_regs.size(); ++b) { b_ip.push_back(lagrange(F, P.b_xs[b], P.b_ys[b])); b_zp.push_back(zpoly(F, P.b_xs[b])); }
    const int nB = (int)P.b_regs.size();
    const u64 x_last = F.pow(P.w_n, (u128)(T - 1) * (u128)E);
    Vec lc_values;
    for (uint32_t step : exe_pos) {
        const u64 x = F.pow(P.w_n, step);
        const Vec& p = pv[step]; const Vec& n = pv[(uint32_t)((step + E) % N)]; const Vec& s = sv_map[step];
        std::vector<u64> st(n_static + 1); int is = 0;
        for (int k = 0; k < n_static; ++k) {
            if (A.statics[k].kind == 0) st[k] = eval_poly(F, cyc_poly[k], F.pow(x, (u128)(T / (long long)A.statics[k].values.size())));
            else if (A.statics[k].kind == 1) st[k] = s[is++];
            else st[k] = eval_poly(F, pub_poly[k], x);
        }
        Vec q(A.K), slots;
        if (!run(F, A.evaluation, p.data(), n.data(), st.data(), slots, q.data())) return "bad evaluation program";
        u64 acc = 0;
        for (int k = 0; k < A.K; ++k) {
            u64 c = coef[k];
            if (P.adj_idx[k] >= 0) c = F.add(c, F.mul(coef[P.adj_idx[k]], F.pow(x, (u128)P.incr[k])));
            acc = F.add(acc, F.mul(c, q[k]));
        }
        // D = qc / Z(x), Z(x) = (x^T - 1) / (x - x_last)   (ZeroPolynomial.ts:28-34)
        const u64 zx = F.mul(F.sub(F.pow(x, (u128)T), 1), F.inv(F.sub(x, x_last)));
        u64 cval = F.mul(acc, F.inv(zx));
        for (int b = 0; b < nB; ++b) {
            const u64 bv = F.mul(F.sub(p[P.b_regs[b]], eval_poly(F, b_ip[b], x)), F.inv(eval_poly(F, b_zp[b], x)));
            u64 c = coef[P.d_count + b];
            if (delta > 0) c = F.add(c, F.mul(coef[P.d_count + nB + b], F.pow(x, (u128)delta)));
            cval = F.add(cval, F.mul(c, bv));
        }
        u64 lv = cval;
        for (int j = 0; j < P.n_lc; ++j) {
            u64 c = coef[P.d_count + P.b_count + j];
            if (delta > 0) c = F.add(c, F.mul(coef[P.d_count + P.b_count + P.n_lc + j], F.pow(x, (u128)delta)));
            lv = F.add(lv, F.mul(c, j < R ? p[j] : s[j - R]));
        }
        lc_values.push_back(lv);
    }
    // FRI verifier (LowDegreeProver.ts:70-172)
    auto column_values = [&](const ParsedBatch& b, const std::vector<uint32_t>& positions, const std::vector<uint32_t>& aug, long long column_length, Vec& outv) -> bool {
        const uint32_t row = (uint32_t)(column_length >> 2);
        for (uint32_t p : positions) {
            size_t idx = 0; for (; idx < aug.size(); ++idx) if (aug[idx] == p % row) break;
            if (idx >= b.values.size()) return false;
            u64 v; if (!F.get(b.values[idx].data() + (p / row) * F.es, &v)) return false;
            outv.push_back(v);
        }
        return true;
    };
    auto rehash = [&](const ParsedBatch& b) { std::vector<Digest> h; for (auto& v : b.values) h.push_back(H.digest(v.data(), v.size())); return h; };
    long long column_length = N;
    { const auto lc_pos = aug4(exe_pos, column_length); Vec checks;
      if (lcp.values.size() != lc_pos.size() || !column_values(lcp, exe_pos, lc_pos, column_length, checks)) return "Verification of linear combination Merkle proof failed";
      if (!verify_batch(lc_root, lc_pos, rehash(lcp), lcp.nodes, lcp.depth, H)) return "Verification of linear combination Merkle proof failed";
      for (size_t i = 0; i < checks.size(); ++i) if (checks[i] != lc_values[i]) return "Verification of linear combination correctness failed"; }
    Digest p_root = lc_root; u64 rou = P.w_n; long long md = P.comp_degree;
    column_length >>= 2;
    for (int depth = 0; depth < n_comp; ++depth) {
        const Comp& c = comps[depth];
        if (column_length < 8) return "Verification of low degree failed: malformed proof (column too short)";
        std::vector<uint32_t> positions;
        if (pseudorandom_indexes(c.root.data(), fri_q, (uint64_t)column_length, (uint64_t)E, positions, err) != 0) return "Verification of low degree failed: " + err;
        const auto aug = aug4(positions, column_length);
        Vec col_vals;
        if (c.column.values.size() != aug.size() || !column_values(c.column, positions, aug, column_length, col_vals)) return "Verification of column Merkle proof failed at depth " + std::to_string(depth);
        if (!verify_batch(c.root, aug, rehash(c.column), c.column.nodes, c.column.depth, H)) return "Verification of column Merkle proof failed at depth " + std::to_string(depth);
        if (c.poly.values.size() != positions.size() || !verify_batch(p_root, positions, rehash(c.poly), c.poly.nodes, c.poly.depth, H))
            return "Verification of polynomial Merkle proof failed at depth " + std::to_string(depth);
        const u64 xs = F.prng_one(p_root.data(), 32);
        const long long L = column_length << 2;
        const u64 iota = F.pow(rou, (u128)(L / 4));
        for (size_t i = 0; i < positions.size(); ++i) {
            Vec xq(4), yq(4); const u64 xe = F.pow(rou, positions[i]); u64 f = 1;
            for (int j = 0; j < 4; ++j) { xq[j] = F.mul(f, xe); f = F.mul(f, iota); if (!F.get(c.poly.values[i].data() + j * F.es, &yq[j])) return "malformed proof (non-canonical value)"; }
            if (eval_poly(F, lagrange(F, xq, yq), xs) != col_vals[i]) return "Degree 4 polynomial didn't evaluate to column value at depth " + std::to_string(depth);
        }
        p_root = c.root; rou = F.pow(rou, 4); md /= 4; column_length >>= 2;
    }
    if (md > (long long)rem.size()) return "Remainder degree is greater than number of remainder values";
    { const size_t Q = rem.size() / 4; if (Q < 1) return "malformed proof (remainder)";
      std::vector<Dg> rows(Q); std::vector<uint8_t> buf(4 * F.es);
      for (size_t i = 0; i < Q; ++i) { for (int j = 0; j < 4; ++j) F.put(buf.data() + j * F.es, rem[i + j * Q]); rows[i] = H.digest(buf.data(), buf.size()); }
      if (Q & (Q - 1)) return "malformed proof (remainder)";
      const Tree t = make_tree(H, rows);
      if (!(t.root() == p_root)) return "Remainder values do not match Merkle root of the last column"; }
    const std::string e = check_remainder(F, rem, md, rou, E);
    if (!e.empty()) return e;
    return "";
}

}  // namespace small
}  // namespace gs
