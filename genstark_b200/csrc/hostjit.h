// Specialised execution-trace generation: the AIR's transition function is emitted as straight-line C++ over
// the weakly reduced host arithmetic (hostfield.h), compiled once per distinct program with the host compiler
// and loaded with dlopen -- the native counterpart of what the reference does at instantiate() time, where
// air-assembly generates JavaScript source for the transition function and evaluates it (called from
// lib/Stark.ts:97 as context.generateExecutionTrace()).  The compiled loop keeps the state in registers across
// steps, so a MiMC chain runs at the latency of its two dependent modular multiplications instead of at
// interpreter speed.  No compiler, GS_TRACE_JIT=0, or any failure => the interpreter in hostair.h (same results;
// tests/test_trace_jit.py compares them).
#pragma once
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <condition_variable>
#include <deque>
#include <functional>
#include "hostair.h"
#include "hostcrypto.h"
#include "hostfield_fast.h"
#include "jitcache.h"

namespace gs {

struct JitArgs {
    w128* state;                    // R values, in/out (weakly reduced)
    const w128* const* stat;        // per static register: table of values (cycle values or a T-length input column)
    const u64_t* stat_mask;         // index = step & mask
    w128* trace;                    // R x T canonical residues
    long long T, s0, s1;            // visits rows s0..s1-1; applies the transition after each of them except step T-1
    long long w0;                   // rows below w0 are visited but not written (warm-up step of a parallel chunk)
};
typedef void (*JitTraceFn)(const JitArgs*);

struct TraceJit {
    void* dl = nullptr;
    JitTraceFn fn = nullptr;
    std::string status;             // "jit <hash>" or the reason the interpreter is used
    ~TraceJit() { /* handles stay loaded for the life of the process (shared by every Stark with this program) */ }
};

static inline std::string jit_hex_u64(u64_t v) { char b[32]; snprintf(b, sizeof b, "0x%llxull", v); return b; }

// ---- mask specialisation -------------------------------------------------------------------------------------------
// Multi-instance AIRs select between sub-computations with 0/1 cyclic registers: next = m * A + (1 - m) * B (the full /
// partial round of Poseidon, "first step of a segment" masks).  Evaluated as written, every step pays for both A and B.
// The masks are static registers whose cycle values are all 0 or 1 -- known when the code is generated -- so the transition
// is specialised once per combination of mask values that occurs: the mask becomes a constant, x * 0 / x * 1 / x + 0 fold
// away and what only fed the dead branch is dropped; the step loop switches on the mask values of the step.  Same values
// (the folds are exact field identities), ~2.5x fewer multiplications per Poseidon step.
struct JitVariant { unsigned combo; HostProgram prog; };

// pr with static register mask_regs[j] fixed to bit j of `bits`: constants folded, identities aliased, dead code dropped.
// Result in SSA form (a fresh slot per value).  false: a shape this pass does not handle (the generic body is used).
static inline bool jit_specialise(const HostProgram& pr, const std::vector<int>& mask_regs, unsigned bits, HostProgram& out) {
    const size_t n = pr.instrs.size();
    struct Val { int kind = -1; u128 k = 0; int node = -1; };            // kind 0: node (instruction index), 1: known constant
    std::vector<Val> val(n);
    std::vector<int> def_of_slot(pr.n_slots + 1, -1);
    struct Node { uint32_t op; Val a, b; uint32_t aux; };                   // aux: register / static index, or the exponent's constant index
    std::vector<Node> node(n);
    std::vector<Val> outs(pr.n_out);
    auto known = [](const Val& v, u128 x) { return v.kind == 1 && v.k == x; };
    auto K = [](u128 x) { Val v; v.kind = 1; v.k = x; return v; };
    for (size_t i = 0; i < n; ++i) {
        const uint32_t op = pr.instrs[i][0], d = pr.instrs[i][1], x = pr.instrs[i][2], y = pr.instrs[i][3];
        Val A, B;
        const bool bin = (op == OP_ADD || op == OP_SUB || op == OP_MUL), un = (op == OP_NEG || op == OP_INV || op == OP_EXP || op == OP_OUT);
        if (bin || un) { if (x > (uint32_t)pr.n_slots || def_of_slot[x] < 0) return false; A = val[def_of_slot[x]]; }
        if (bin) { if (y > (uint32_t)pr.n_slots || def_of_slot[y] < 0) return false; B = val[def_of_slot[y]]; }
        Val r; r.kind = 0; r.node = (int)i;
        node[i] = Node{op, A, B, 0};
        switch (op) {
            case OP_CONST: if (x >= pr.consts.size()) return false; r = K(pr.consts[x]); break;
            case OP_CUR: node[i].aux = x; break;
            case OP_STATIC: {
                node[i].aux = x;
                for (size_t j = 0; j < mask_regs.size(); ++j) if ((uint32_t)mask_regs[j] == x) r = K((bits >> j) & 1u);
                break;
            }
            case OP_ADD:
                if (A.kind == 1 && B.kind == 1) r = K(h_add(A.k, B.k));
                else if (known(A, 0)) r = B;
                else if (known(B, 0)) r = A;
                break;
            case OP_SUB:
                if (A.kind == 1 && B.kind == 1) r = K(h_sub(A.k, B.k));
                else if (known(B, 0)) r = A;
                break;
            case OP_MUL:
                if (known(A, 0) || known(B, 0)) r = K(0);
                else if (A.kind == 1 && B.kind == 1) r = K(h_mul(A.k, B.k));
                else if (known(A, 1)) r = B;
                else if (known(B, 1)) r = A;
                break;
            case OP_NEG: if (A.kind == 1) r = K(h_sub(0, A.k)); break;
            case OP_INV: if (A.kind == 1) r = K(h_inv(A.k)); break;
            case OP_EXP: if (y >= pr.consts.size()) return false; node[i].aux = y; if (A.kind == 1) r = K(h_pow(A.k, pr.consts[y])); break;
            case OP_OUT: if (d >= (uint32_t)pr.n_out) return false; outs[d] = A; continue;
            default: return false;
        }
        if (d > (uint32_t)pr.n_slots) return false;
        val[i] = r;
        def_of_slot[d] = (int)i;
    }
    for (const Val& v : outs) if (v.kind < 0) return false;
    // liveness from the outputs
    std::vector<char> live(n, 0);
    std::vector<int> stack;
    for (const Val& v : outs) if (v.kind == 0 && !live[v.node]) { live[v.node] = 1; stack.push_back(v.node); }
    while (!stack.empty()) {
        const int i = stack.back(); stack.pop_back();
        for (const Val* v : {&node[i].a, &node[i].b}) if (v->kind == 0 && !live[v->node]) { live[v->node] = 1; stack.push_back(v->node); }
    }
    // re-emit in SSA form
    out = HostProgram(); out.n_out = pr.n_out;
    std::map<u128, uint32_t> const_idx;
    auto cidx = [&](u128 x) { auto it = const_idx.find(x); if (it != const_idx.end()) return it->second; const uint32_t k = (uint32_t)out.consts.size(); out.consts.push_back(x); const_idx[x] = k; return k; };
    std::vector<int> slot_of(n, -1);
    uint32_t next_slot = 0;
    auto operand = [&](const Val& v) -> uint32_t {                       // slot holding the operand (constants get a slot of their own)
        if (v.kind == 0) return (uint32_t)slot_of[v.node];
        const uint32_t sl = next_slot++;
        out.instrs.push_back({(uint32_t)OP_CONST, sl, cidx(v.k), 0u});
        return sl;
    };
    for (size_t i = 0; i < n; ++i) {
        if (!live[i]) continue;
        const Node& nd = node[i];
        uint32_t a = 0, b = 0;
        switch (nd.op) {
            case OP_CUR: case OP_STATIC: a = nd.aux; break;
            case OP_ADD: case OP_SUB: case OP_MUL: a = operand(nd.a); b = operand(nd.b); break;
            case OP_NEG: case OP_INV: a = operand(nd.a); break;
            case OP_EXP: a = operand(nd.a); b = cidx(pr.consts[nd.aux]); break;
            default: return false;
        }
        slot_of[i] = (int)next_slot++;
        out.instrs.push_back({nd.op, (uint32_t)slot_of[i], a, b});
    }
    for (int r = 0; r < pr.n_out; ++r) out.instrs.push_back({(uint32_t)OP_OUT, (uint32_t)r, operand(outs[r]), 0u});
    out.n_slots = (int)std::max<uint32_t>(next_slot, 1u);
    return true;
}

// the static registers that are 0/1 cycles (at most four are used), and the combinations of their values that occur
static inline void jit_find_masks(const AirHost* S, std::vector<int>& mask_regs, std::vector<unsigned>& combos, std::vector<size_t>* counts = nullptr) {
    mask_regs.clear(); combos.clear(); if (counts) counts->clear();
    if (const char* e = getenv("GS_TRACE_SPECIALISE")) if (e[0] == '0') return;
    size_t period = 1;
    for (size_t k = 0; k < S->statics.size() && mask_regs.size() < 4; ++k) {
        const StaticReg& sr = S->statics[k];
        if (sr.kind != 0 || sr.values.empty() || (sr.values.size() & (sr.values.size() - 1))) continue;
        bool binary = true;
        for (u128 v : sr.values) if (v > 1) { binary = false; break; }
        if (!binary) continue;
        bool used = false;
        for (const auto& ins : S->transition.instrs) if (ins[0] == OP_STATIC && ins[2] == (uint32_t)k) { used = true; break; }
        if (!used) continue;
        mask_regs.push_back((int)k);
        period = std::max(period, sr.values.size());
    }
    if (mask_regs.empty()) return;
    std::vector<int> seen(1u << mask_regs.size(), -1);
    for (size_t s = 0; s < period; ++s) {
        unsigned c = 0;
        for (size_t j = 0; j < mask_regs.size(); ++j) { const auto& v = S->statics[mask_regs[j]].values; c |= (unsigned)(v[s & (v.size() - 1)] & 1) << j; }
        if (seen[c] < 0) { seen[c] = (int)combos.size(); combos.push_back(c); if (counts) counts->push_back(0); }
        if (counts) (*counts)[seen[c]]++;
    }
}
// rough cost of one step of a program in multiplications (an exponentiation or an inversion is a long chain of them)
static inline double jit_program_cost(const HostProgram& pr) {
    double c = 0;
    for (const auto& ins : pr.instrs) {
        if (ins[0] == OP_MUL) c += 1;
        else if (ins[0] == OP_EXP) { u128 e = ins[3] < pr.consts.size() ? pr.consts[ins[3]] : 0; int bits = 0; while (e) { ++bits; e >>= 1; } c += 1.5 * bits; }
        else if (ins[0] == OP_INV) c += 190;
        else if (ins[0] == OP_ADD || ins[0] == OP_SUB || ins[0] == OP_NEG) c += 0.1;
    }
    return c;
}

static inline void jit_emit_consts(std::ostringstream& o, const HostProgram& pr, const std::string& kp) {
    for (size_t i = 0; i < pr.consts.size(); ++i)
        o << "static const w128 " << kp << i << " = {" << jit_hex_u64((u64_t)pr.consts[i]) << ", " << jit_hex_u64((u64_t)(pr.consts[i] >> 64)) << "};\n";
}

// One step of a transition program as straight-line SSA (state s<r> -> s<r>); constants are named <kp><index>.  Products
// that feed exactly one further multiplication or addition are fused with it (f_mul3_add / f_mul_add,
// hostfield_fast.h), which removes a modular reduction from the dependency chain of S-boxes such as x^3 + k.
static inline bool jit_emit_body(std::ostringstream& o, const HostProgram& pr, int R, int n_static, const std::string& kp, const std::string& ind) {
    const size_t n = pr.instrs.size();
    // definitions: which instruction defines each operand, and how often each definition is read
    std::vector<int> def_of_slot(pr.n_slots + 1, -1), da(n, -1), db(n, -1), uses(n, 0);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t op = pr.instrs[i][0];
        const bool bin = (op == OP_ADD || op == OP_SUB || op == OP_MUL), un = (op == OP_NEG || op == OP_INV || op == OP_EXP || op == OP_OUT);
        if (bin) { da[i] = def_of_slot[pr.instrs[i][2]]; db[i] = def_of_slot[pr.instrs[i][3]]; if (da[i] < 0 || db[i] < 0) return false; uses[da[i]]++; uses[db[i]]++; }
        if (un) { da[i] = def_of_slot[pr.instrs[i][2]]; if (da[i] < 0) return false; uses[da[i]]++; }
        if (op == OP_NEXT || op > OP_OUT) return false;
        if (op != OP_OUT) def_of_slot[pr.instrs[i][1]] = (int)i;
    }
    // fusion: child[j] = the product folded into instruction j; that product is not emitted on its own
    std::vector<int> child(n, -1), other(n, -1); std::vector<char> deferred(n, 0);
    auto is_mul = [&](int d) { return d >= 0 && pr.instrs[d][0] == OP_MUL; };
    for (size_t j = 0; j < n; ++j) {
        const uint32_t op = pr.instrs[j][0];
        if (op != OP_MUL && op != OP_ADD) continue;
        if (da[j] == db[j]) continue;                                       // x*x, x+x: nothing to fold
        for (int side = 0; side < 2 && child[j] < 0; ++side) {
            const int d = side ? db[j] : da[j], o2 = side ? da[j] : db[j];
            if (!is_mul(d) || uses[d] != 1 || deferred[d]) continue;
            if (op == OP_MUL && child[d] >= 0) continue;                     // a product of at most three factors
            child[j] = d; other[j] = o2; deferred[d] = 1;
        }
    }
    std::vector<std::string> name(n);            // expression naming each definition's value
    std::vector<std::string> nxt(R);
    auto operands = [&](int d) { return name[da[d]] + ", " + name[db[d]]; };      // of a product
    for (size_t i = 0; i < n; ++i) {
        const uint32_t op = pr.instrs[i][0], d = pr.instrs[i][1], x = pr.instrs[i][2], y = pr.instrs[i][3];
        const std::string v = "v" + std::to_string(i);
        if (deferred[i]) continue;
        switch (op) {
            case OP_CONST: if (x >= pr.consts.size()) return false; name[i] = kp + std::to_string(x); break;
            case OP_CUR: if ((int)x >= R) return false; name[i] = "s" + std::to_string(x); break;
            case OP_STATIC: if ((int)x >= n_static) return false; o << ind << "const w128 " << v << " = st" << x << "[(u64_t)s & m" << x << "];\n"; name[i] = v; break;
            case OP_ADD:
                if (child[i] >= 0) {
                    const int m = child[i];
                    if (child[m] >= 0) o << ind << "const w128 " << v << " = f_mul3_add(" << operands(child[m]) << ", " << name[other[m]] << ", " << name[other[i]] << ");\n";
                    else o << ind << "const w128 " << v << " = f_mul_add(" << operands(m) << ", " << name[other[i]] << ");\n";
                } else o << ind << "const w128 " << v << " = w_add(" << name[da[i]] << ", " << name[db[i]] << ");\n";
                name[i] = v; break;
            case OP_MUL:
                if (child[i] >= 0) o << ind << "const w128 " << v << " = f_mul3_add(" << operands(child[i]) << ", " << name[other[i]] << ", k_zero);\n";
                else o << ind << "const w128 " << v << " = f_mul_add(" << name[da[i]] << ", " << name[db[i]] << ", k_zero);\n";
                name[i] = v; break;
            case OP_SUB: o << ind << "const w128 " << v << " = w_sub(" << name[da[i]] << ", " << name[db[i]] << ");\n"; name[i] = v; break;
            case OP_NEG: o << ind << "const w128 " << v << " = w_sub(k_zero, " << name[da[i]] << ");\n"; name[i] = v; break;
            case OP_INV: o << ind << "const w128 " << v << " = w_inv(" << name[da[i]] << ");\n"; name[i] = v; break;
            case OP_EXP: {
                if (y >= pr.consts.size()) return false;
                const u128 e = pr.consts[y];
                o << ind << "const w128 " << v << " = w_pow(" << name[da[i]] << ", " << jit_hex_u64((u64_t)e) << ", " << jit_hex_u64((u64_t)(e >> 64)) << ");\n";
                name[i] = v; break;
            }
            case OP_OUT: if ((int)d < R) nxt[d] = name[da[i]]; break;
            default: return false;
        }
    }
    for (int r = 0; r < R; ++r) if (nxt[r].empty()) return false;
    // the new state is assigned after every output is computed (outputs may read the old state)
    for (int r = 0; r < R; ++r) o << ind << "const w128 n" << r << " = " << nxt[r] << ";\n";
    for (int r = 0; r < R; ++r) o << ind << "s" << r << " = n" << r << ";\n";
    return true;
}

// C++ source of the block function for this transition program: the generic step, and one specialised step per combination
// of mask values (variants; mask_regs[j] = static register whose value is bit j of a combination)
static inline std::string jit_emit_source(const HostProgram& pr, int R, int n_static, const std::vector<int>& mask_regs = {},
                                          const std::vector<JitVariant>& variants = {}) {
    std::ostringstream o;
    o << "#include <x86intrin.h>\n" << GS_HOSTFIELD_SRC << "\n" << GS_HOSTFIELD_FAST_SRC << "\n";
    o << "struct JitArgs { w128* state; const w128* const* stat; const u64_t* stat_mask; w128* trace; long long T, s0, s1, w0; };\n";
    o << "static const w128 k_zero = {0, 0};\n";
    jit_emit_consts(o, pr, "k");
    for (size_t v = 0; v < variants.size(); ++v) jit_emit_consts(o, variants[v].prog, "q" + std::to_string(v) + "_");
    o << "extern \"C\" void gs_trace_block(const JitArgs* a) {\n";
    o << "  const long long T = a->T, w0 = a->w0; w128* const tr = a->trace;\n";
    for (int r = 0; r < R; ++r) o << "  w128 s" << r << " = a->state[" << r << "];\n";
    for (int k = 0; k < n_static; ++k) o << "  const w128* const st" << k << " = a->stat[" << k << "]; const u64_t m" << k << " = a->stat_mask[" << k << "];\n";
    o << "  for (long long s = a->s0; s < a->s1; ++s) {\n";
    o << "    if (__builtin_expect(s >= w0, 1)) {\n";
    for (int r = 0; r < R; ++r) o << "      tr[" << r << " * T + s] = w_from(w_canon(s" << r << "));\n";
    o << "    }\n";
    o << "    if (s + 1 == T) break;\n";
    if (variants.empty()) {
        if (!jit_emit_body(o, pr, R, n_static, "k", "    ")) return "";
    } else {
        // the mask values of this step select the specialised step; any other value (a mask that is not 0 / 1 after all)
        // takes the generic one
        o << "    unsigned combo = 0, exact = 1;\n";
        for (size_t j = 0; j < mask_regs.size(); ++j) {
            const int k = mask_regs[j];
            if (k < 0 || k >= n_static) return "";
            o << "    { const w128 mv = st" << k << "[(u64_t)s & m" << k << "]; combo |= (unsigned)(mv.lo & 1) << " << j << "; exact &= (unsigned)(mv.hi == 0 && mv.lo <= 1); }\n";
        }
        o << "    switch (exact ? combo : ~0u) {\n";
        for (size_t v = 0; v < variants.size(); ++v) {
            o << "      case " << variants[v].combo << "u: {\n";
            if (!jit_emit_body(o, variants[v].prog, R, n_static, "q" + std::to_string(v) + "_", "        ")) return "";
            o << "      } break;\n";
        }
        o << "      default: {\n";
        if (!jit_emit_body(o, pr, R, n_static, "k", "        ")) return "";
        o << "      } break;\n";
        o << "    }\n";
    }
    o << "  }\n";
    for (int r = 0; r < R; ++r) o << "  a->state[" << r << "] = s" << r << ";\n";
    o << "}\n";
    return o.str();
}

// compile (or reuse) the block function of a transition program; never throws, never fails the prove
std::shared_ptr<TraceJit> jit_get(const AirHost* S);
#ifdef GS_HOSTAIR_IMPL
std::shared_ptr<TraceJit> jit_get(const AirHost* S) {
    const HostProgram& pr = S->transition; const int R = S->R, n_static = (int)S->statics.size();
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<TraceJit>> cache;
    // the fallback is never silent: one line on stderr per program (GS_TRACE_JIT=0 is a request, not a failure), and
    // gs_trace_backend() / bench.py's `backends.trace` say which generator ran
    auto interp = [](const std::string& why) {
        auto j = std::make_shared<TraceJit>(); j->status = "interpreter (" + why + ")";
        if (why != "GS_TRACE_JIT=0") {
            static std::mutex wmu; static std::map<std::string, bool> warned;
            std::lock_guard<std::mutex> wl(wmu);
            if (!warned[why]) {
                warned[why] = true;
                fprintf(stderr, "genstark_b200: execution-trace JIT unavailable (%s); using the interpreter (about 1.5x slower per step)\n", why.c_str());
            }
        }
        return j;
    };
    if (const char* e = getenv("GS_TRACE_JIT")) if (e[0] == '0') return interp("GS_TRACE_JIT=0");
    std::vector<int> mask_regs; std::vector<unsigned> combos; std::vector<size_t> counts; std::vector<JitVariant> variants;
    jit_find_masks(S, mask_regs, combos, &counts);
    double weighted = 0, total = 0;
    for (size_t ci = 0; ci < combos.size(); ++ci) {
        JitVariant v; v.combo = combos[ci];
        if (!jit_specialise(pr, mask_regs, combos[ci], v.prog) || validate_program(v.prog, R, n_static, true) != nullptr) { variants.clear(); break; }
        weighted += (double)counts[ci] * jit_program_cost(v.prog); total += (double)counts[ci];
        variants.push_back(std::move(v));
    }
    // worth it only when the average step gets markedly cheaper (Poseidon: 0.4x); a select that guards a few multiplications
    // next to two 128-bit exponentiations (Rescue) only adds a switch and code
    const char* force = getenv("GS_TRACE_SPECIALISE");                       // "2": specialise whatever the estimate says (tests)
    if (!variants.empty() && !(force && force[0] == '2') && weighted / total > 0.8 * jit_program_cost(pr)) variants.clear();
    if (variants.empty()) mask_regs.clear();
    const std::string src = jit_emit_source(pr, R, n_static, mask_regs, variants);
    if (src.empty()) return interp("program not supported by the code generator");
    if (const char* dump = getenv("GS_JIT_DUMP")) { if (FILE* f = fopen(dump, "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); } }   // the generated source, for inspection
    const char* cxx = getenv("GS_JIT_CXX"); if (!cxx) cxx = getenv("CXX"); if (!cxx) cxx = "g++";
    // -march=native: the code runs on the machine that compiles it (mulx / adx shorten the carry chains); compiler, flags
    // and CPU model are part of the key, so a cache on a shared home never serves another machine's object
    const std::vector<std::string> flags = {"-O3", "-march=native", "-std=c++17", "-fPIC", "-shared"};
    std::string keyed = src + "\n//cxx " + cxx;
    for (const std::string& f : flags) keyed += " " + f;
    keyed += "\n//cpu " + jit_cpu_tag();
    uint8_t dg[32]; sha256_bytes((const uint8_t*)keyed.data(), keyed.size(), dg);
    char hex[33]; for (int i = 0; i < 16; ++i) snprintf(hex + 2 * i, 3, "%02x", dg[i]);
    const std::string key(hex);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto done = [&](std::shared_ptr<TraceJit> j) { cache[key] = j; return j; };
    std::string why;
    const std::string dir = jit_cache_dir(&why);
    if (dir.empty()) return done(interp(why));
    const std::string so = dir + "/trace_" + key + ".so";
    struct stat st;
    if (lstat(so.c_str(), &st) == 0 && !jit_file_trusted(so)) return done(interp(so + " exists but is not a private regular file of this user"));
    if (!jit_file_trusted(so)) {
        const std::string tag = dir + "/trace_" + key + "." + std::to_string((long)getpid());
        const std::string cpp = tag + ".cpp", tmp = tag + ".so.tmp", log = tag + ".log";
        FILE* f = fopen(cpp.c_str(), "w");
        if (!f) return done(interp("cannot write " + cpp));
        fwrite(src.data(), 1, src.size(), f); fclose(f);
        std::vector<std::string> argv = {cxx};
        argv.insert(argv.end(), flags.begin(), flags.end());
        argv.insert(argv.end(), {"-o", tmp, cpp});
        const int rc = jit_spawn(argv, log);
        if (rc != 0 || rename(tmp.c_str(), so.c_str()) != 0) { unlink(tmp.c_str()); return done(interp(std::string("host compiler failed: ") + cxx + " (log: " + log + ")")); }
        unlink(cpp.c_str()); unlink(log.c_str());
    }
    auto j = std::make_shared<TraceJit>();
    j->dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!j->dl) return done(interp(std::string("dlopen: ") + dlerror()));
    j->fn = (JitTraceFn)dlsym(j->dl, "gs_trace_block");
    if (!j->fn) return done(interp("gs_trace_block missing in " + so));
    j->status = "jit " + key;
    return done(j);
}
#endif

#ifdef GS_HOSTAIR_IMPL
// Workers for the chunks of a multi-instance trace, created once and kept: no thread creation per prove (15 of them for 16
// chunks) -- Poseidon Merkle-proof trace, 8 chunks in this container: 20.9 -> 18.4 ms steady state.  (Either way the first
// ~0.8 s of multi-threaded work of a process runs at 2-4x the steady time here: the virtual CPUs have to wake up; not something
// the library can fix.)  The pool is leaked on purpose (nothing to join at exit) and rebuilt in a forked child, where the
// threads do not exist.
class TracePool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::pair<std::function<void()>, int*>> q;          // task, the counter of the call it belongs to
    size_t n_workers = 0;
    void loop() {
        for (;;) {
            std::pair<std::function<void()>, int*> t;
            { std::unique_lock<std::mutex> l(mu); cv_work.wait(l, [&] { return !q.empty(); }); t = std::move(q.front()); q.pop_front(); }
            t.first();
            { std::lock_guard<std::mutex> l(mu); if (--*t.second == 0) cv_done.notify_all(); }
        }
    }
public:
    static TracePool& get() {
        static std::mutex gmu; static TracePool* pool = nullptr; static pid_t owner = 0;
        std::lock_guard<std::mutex> l(gmu);
        if (!pool || owner != getpid()) { pool = new TracePool(); owner = getpid(); }
        return *pool;
    }
    // runs every task; the last one on the calling thread, the others on workers (or here, when no worker can be had)
    void run(std::vector<std::function<void()>>& tasks) {
        if (tasks.empty()) return;
        const size_t want = tasks.size() - 1;
        {
            std::lock_guard<std::mutex> l(mu);
            while (n_workers < want) {
                try { std::thread([this] { loop(); }).detach(); ++n_workers; }
                catch (...) { break; }
            }
        }
        int left = 0;
        size_t queued = 0;
        {
            std::lock_guard<std::mutex> l(mu);
            if (n_workers > 0) { for (; queued < want; ++queued) { q.emplace_back(tasks[queued], &left); ++left; } }
        }
        if (queued) cv_work.notify_all();
        for (size_t k = queued; k < tasks.size(); ++k) tasks[k]();          // the caller's share (everything, without workers)
        if (queued) { std::unique_lock<std::mutex> l(mu); cv_done.wait(l, [&] { return left == 0; }); }
    }
};

static thread_local std::string g_trace_backend = "not run";
const char* trace_backend_status() { return g_trace_backend.c_str(); }

void trace_prepare(const AirHost* S) { g_trace_backend = jit_get(S)->status; }

void generate_trace(const AirHost* S, const u128* init_state, const fp* input_traces, fp* tr, const TraceChunkFn* on_chunk) {
    const int R = S->R; const long long T = 1ll << S->log_t;
    const size_t n_stat = S->statics.size();
    const std::shared_ptr<TraceJit> jit = jit_get(S);
    g_trace_backend = jit->status;
    if (jit->fn) {
        std::vector<w128> state(R);
        for (int r = 0; r < R; ++r) state[r] = w_from(init_state[r]);
        std::vector<const w128*> stat(n_stat + 1, nullptr);
        std::vector<u64_t> mask(n_stat + 1, 0);
        int ii = 0;
        for (size_t k = 0; k < n_stat; ++k) {
            const StaticReg& sr = S->statics[k];
            // u128 and fp are both 16 little-endian bytes: the tables are read in place
            if (sr.kind == 0) { stat[k] = reinterpret_cast<const w128*>(sr.values.data()); mask[k] = sr.values.size() - 1; }
            else { stat[k] = reinterpret_cast<const w128*>(input_traces + (size_t)(ii++) * T); mask[k] = ~0ull; }
        }
        JitArgs a{state.data(), stat.data(), mask.data(), reinterpret_cast<w128*>(tr), T, 0, 0, 0};
        // Segments in parallel.  A step whose transition does not read the current state (the last step of a `for each`
        // segment: every output is mask * f(inputs) + (1 - mask) * g(state) with mask = 1) cuts the chain: the rows behind it
        // follow from the static registers alone.  Such steps sit at the end of power-of-two cycles, so step b - 1 is tried for
        // every chunk boundary b (and a few steps after it); "does not read the state" is tested on the compiled function
        // itself with two random states -- equal outputs mean a constant polynomial in the state except with probability
        // ~2^-120.  No such step (MiMC: one chain) => the sequential loop below.  GS_TRACE_THREADS=1 turns this off.
        int threads = (int)std::thread::hardware_concurrency(); if (threads > 16) threads = 16;
        if (const char* e = getenv("GS_TRACE_THREADS")) threads = atoi(e);
        std::vector<long long> cut;                     // cut[k]: first row of chunk k
        // chunks of at least 256 steps: a chunk is handed to a waiting worker in a few microseconds (TracePool), and 256 steps
        // of a hash round are >= 100 us of work (2^12 steps of the Rescue chain: 16 chunks instead of 4)
        if (threads >= 2 && T >= 1024 && R <= 64) {
            int P = 1; while (2 * P <= threads && T / (2 * P) >= 256) P *= 2;
            const long long chunk = T / P;
            auto state_free = [&](long long s) {        // transition at step s ignores the current state?
                w128 sa[64], sb[64];
                u64_t seed = 0x9E3779B97F4A7C15ull ^ (u64_t)s;
                auto next = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
                for (int r = 0; r < R; ++r) { sa[r] = w128{next(), next() >> 1}; sb[r] = w128{next(), next() >> 1}; }
                JitArgs t = a; t.s0 = s; t.s1 = s + 1; t.w0 = T;
                t.state = sa; jit->fn(&t);
                t.state = sb; jit->fn(&t);
                for (int r = 0; r < R; ++r) if (sa[r].lo != sb[r].lo || sa[r].hi != sb[r].hi) return false;
                return true;
            };
            cut.push_back(0);
            for (int k = 1; k < P && !cut.empty(); ++k) {
                long long found = -1;
                for (long long s = k * chunk - 1; s < k * chunk - 1 + 64 && s + 1 < T; ++s) if (state_free(s)) { found = s + 1; break; }
                if (found < 0 || found <= cut.back()) cut.clear(); else cut.push_back(found);
            }
        }
        if (cut.size() >= 2) {
            const int P = (int)cut.size();
            std::vector<std::vector<w128>> st(P, std::vector<w128>(R, w128{0, 0}));
            st[0] = state;
            std::vector<std::function<void()>> tasks;
            for (int k = 0; k < P; ++k) {
                JitArgs ak = a;
                ak.state = st[k].data();
                ak.s0 = k == 0 ? 0 : cut[k] - 1;        // chunk k > 0 starts one step early: that step ignores the state
                ak.w0 = cut[k];
                ak.s1 = k + 1 < P ? cut[k + 1] : T;
                tasks.emplace_back([fn = jit->fn, ak]() { fn(&ak); });
            }
            TracePool::get().run(tasks);                  // chunks are independent; the caller's thread takes the last one
            g_trace_backend = jit->status + " x" + std::to_string(P) + " threads";
            if (on_chunk) for (long long s0 = 0; s0 < T; s0 += 0x10000) (*on_chunk)(s0, s0 + 0x10000 < T ? s0 + 0x10000 : T);
            return;
        }
        for (long long s0 = 0; s0 < T; s0 += 0x10000) {
            a.s0 = s0; a.s1 = s0 + 0x10000 < T ? s0 + 0x10000 : T;
            jit->fn(&a);
            if (on_chunk) (*on_chunk)(a.s0, a.s1);
        }
        return;
    }
    TransitionRunner run; run.init(S->transition, R, (int)n_stat);
    for (int r = 0; r < R; ++r) run.buf[0][r] = w_from(init_state[r]);
    int p = 0;
    for (long long s = 0; s < T; ++s, p ^= 1) {
        const std::vector<w128>& cur = run.buf[p];
        for (int r = 0; r < R; ++r) tr[(size_t)r * T + s] = fp_from_u128(w_canon(cur[r]));
        if (s + 1 < T) {
            int ii = 0;
            for (size_t k = 0; k < n_stat; ++k) {
                const StaticReg& sr = S->statics[k];
                if (sr.kind == 0) run.stat[k] = w_from(sr.values[s & (sr.values.size() - 1)]);
                else run.stat[k] = w_from(fp_to_u128(input_traces[(size_t)(ii++) * T + s]));
            }
            run.step(p);
        }
        if (on_chunk && ((s + 1) & 0xFFFF) == 0) (*on_chunk)(s + 1 - 0x10000, s + 1);
    }
    if (on_chunk && (T & 0xFFFF)) (*on_chunk)(T & ~0xFFFFll, T);
}
#endif

}  // namespace gs
