// Host-only AIR pieces (no CUDA): the flattened AirModule (air.py: pack_air), its parser, and the
// execution-trace generator.  Compiled by the host compiler in host.cpp (GS_HOSTAIR_IMPL) so the one
// sequential stage of prove() gets plain g++ -O3 code generation.
#pragma once
#include <array>
#include <functional>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "fp128.cuh"
#include "hostfield.h"
#include "../../include/genstark_b200.h"

namespace gs {

enum { OP_CONST = 0, OP_CUR = 1, OP_NEXT = 2, OP_STATIC = 3, OP_ADD = 4, OP_SUB = 5, OP_MUL = 6, OP_NEG = 7,
       OP_INV = 8, OP_EXP = 9, OP_OUT = 10 };
#define GS_MAX_COLS 64          // trace + static registers visible to a program
#define GS_MAX_CONSTRAINTS 64

struct HostProgram {
    std::vector<std::array<uint32_t, 4>> instrs;
    std::vector<u128> consts;
    int n_slots = 0, n_out = 0;
};

struct StaticReg {
    int kind = 0;                 // 0 cycle, 1 secret input, 2 public input
    std::vector<u128> values;     // cycle values
};

struct AirHost {
    int R = 0, K = 0, log_t = 0, log_e = 0;
    std::vector<StaticReg> statics;
    std::vector<int> degrees;
    HostProgram transition, evaluation;
    int n_secret = 0, n_public = 0;
};

// ------------------------------------------------------------------------------ AIR blob (air.py)
struct BlobReader {
    const uint8_t* p; size_t n, off = 0; bool ok = true;
    uint32_t u32() { if (off + 4 > n) { ok = false; return 0; } uint32_t v; memcpy(&v, p + off, 4); off += 4; return v; }
    u128 elem() { if (off + 16 > n) { ok = false; return 0; } fp f; memcpy(&f, p + off, 16); off += 16; return fp_to_u128(f); }
};

static inline bool read_program(BlobReader& r, HostProgram& pr) {
    uint32_t ni = r.u32(), nc = r.u32(); pr.n_slots = (int)r.u32(); pr.n_out = (int)r.u32();
    if (!r.ok || ni > (1u << 20) || nc > (1u << 20)) return false;
    pr.instrs.resize(ni);
    for (auto& i : pr.instrs) for (int k = 0; k < 4; ++k) i[k] = r.u32();
    pr.consts.resize(nc);
    for (auto& c : pr.consts) c = r.elem();
    return r.ok;
}

// Every index an instruction carries is checked once, here, so that neither the host interpreter / code generator nor the
// device evaluator (interpreting kernel, NVRTC source) ever indexes with a value taken from the blob unchecked: value slots
// 0 .. n_slots-1 (reused, so "defined before it is read" is tracked per slot), constants, trace registers, static registers,
// outputs (each written at least once).  `next` values exist in the constraint evaluator only.
static inline const char* validate_instrs(const std::vector<std::array<uint32_t, 4>>& instrs, size_t n_consts, int n_slots, int n_out,
                                          int R, int n_static, bool is_transition) {
    if (n_slots < 1 || n_slots > (1 << 20)) return "program: value-slot count out of range";
    if (n_out < 1 || n_out > GS_MAX_CONSTRAINTS) return "program: output count out of range";
    std::vector<char> defined((size_t)n_slots, 0), written((size_t)n_out, 0);
    const uint32_t ns = (uint32_t)n_slots, nc = (uint32_t)n_consts;
    for (const auto& ins : instrs) {
        const uint32_t op = ins[0], d = ins[1], x = ins[2], y = ins[3];
        auto readable = [&](uint32_t slot) { return slot < ns && defined[slot]; };
        switch (op) {
            case OP_CONST: if (x >= nc) return "program: constant index out of range"; break;
            case OP_NEXT: if (is_transition) return "program: a transition function cannot read the next state";   /* fall through */
            case OP_CUR: if (x >= (uint32_t)R) return "program: trace register index out of range"; break;
            case OP_STATIC: if (x >= (uint32_t)n_static) return "program: static register index out of range"; break;
            case OP_ADD: case OP_SUB: case OP_MUL: if (!readable(x) || !readable(y)) return "program: operand slot read before it is written"; break;
            case OP_NEG: case OP_INV: if (!readable(x)) return "program: operand slot read before it is written"; break;
            case OP_EXP: if (!readable(x)) return "program: operand slot read before it is written"; if (y >= nc) return "program: exponent index out of range"; break;
            case OP_OUT:
                if (d >= (uint32_t)n_out) return "program: output index out of range";
                if (!readable(x)) return "program: output reads a slot that was never written";
                written[d] = 1;
                continue;
            default: return "program: unknown opcode";
        }
        if (d >= ns) return "program: destination slot out of range";
        defined[d] = 1;
    }
    for (char w : written) if (!w) return "program: an output is never written";
    return nullptr;
}
static inline const char* validate_program(const HostProgram& pr, int R, int n_static, bool is_transition) {
    return validate_instrs(pr.instrs, pr.consts.size(), pr.n_slots, pr.n_out, R, n_static, is_transition);
}

// parse the AIR blob (air.py: pack_air) into the host part of a Stark; returns "" or an error message
std::string parse_air(const uint8_t* air_blob, size_t blob_len, AirHost* S, int* code);
#ifdef GS_HOSTAIR_IMPL
std::string parse_air(const uint8_t* air_blob, size_t blob_len, AirHost* S, int* code) {
    BlobReader r{air_blob, blob_len};
    *code = GS_E_ARG;
    if (r.u32() != 0x52494147u) return "bad AIR blob magic";
    fp mod;
    for (int i = 0; i < 4; ++i) mod.v[i] = r.u32();
    if (!(mod.v[0] == P0 && mod.v[1] == P1 && mod.v[2] == P2 && mod.v[3] == P3)) { *code = GS_E_UNSUPPORTED; return "no native backend for this modulus (isOptimized = false)"; }
    S->R = (int)r.u32(); S->K = (int)r.u32(); S->log_t = (int)r.u32(); S->log_e = (int)r.u32();
    const uint32_t n_static = r.u32();
    if (!r.ok || S->R < 1 || S->R > GS_MAX_COLS || S->K < 1 || S->K > GS_MAX_CONSTRAINTS || n_static > GS_MAX_COLS) { *code = GS_E_UNSUPPORTED; return "AIR shape out of range"; }
    if (S->log_t < 2 || S->log_t > 30 || S->log_e < 1 || S->log_e > 5) return "trace length >= 4 and extension factor 2..32 required";
    S->statics.resize(n_static);
    S->n_secret = S->n_public = 0;
    for (auto& sr : S->statics) {
        sr.kind = (int)r.u32();
        const uint32_t len = r.u32();
        if (!r.ok || len > (1u << 24) || sr.kind < 0 || sr.kind > 2 || (sr.kind != 0 && len != 0)) return "bad static register";
        sr.values.resize(len);
        for (auto& v : sr.values) v = r.elem();
        if (sr.kind == 0 && (len == 0 || (len & (len - 1)) || len > (1u << S->log_t))) return "cycle length must be a power of two <= steps";
        if (sr.kind == 1) S->n_secret++;
        if (sr.kind == 2) S->n_public++;
    }
    S->degrees.resize(S->K);
    for (auto& d : S->degrees) d = (int)r.u32();
    if (!read_program(r, S->transition) || !read_program(r, S->evaluation)) return "bad AIR program";
    if (S->transition.n_out != S->R || S->evaluation.n_out != S->K) return "program outputs do not match the register / constraint counts";
    if (const char* bad = validate_program(S->transition, S->R, (int)n_static, true)) return std::string("transition ") + bad;
    if (const char* bad = validate_program(S->evaluation, S->R, (int)n_static, false)) return std::string("evaluation ") + bad;
    for (int d : S->degrees) if (d < 0 || d > 256) return "constraint degree out of range";
    for (auto& ins : S->evaluation.instrs) if (ins[0] == OP_EXP) { *code = GS_E_UNSUPPORTED; return "exp with a large exponent in a constraint"; }
    *code = GS_OK;
    return "";
}
#endif

// host interpreter (trace generation): state -> next state, on weakly reduced values (hostfield.h).
// The flat program is re-encoded in accumulator form: loads of trace / static / constant values become
// operand pointers (no copies), and a value consumed by the very next operation stays in a register
// instead of going through memory -- store-to-load forwarding on the dependency chain is what an
// interpreter costs on a chain of dependent cubings.
enum { A_LOAD = 0, A_STORE, A_ADD, A_SUB, A_RSUB, A_MUL, A_SQR, A_NEG, A_INV, A_EXP, A_MOVE, A_DBL };
struct AccIns { uint32_t op; w128* dst; const w128* src; u128 e; };
struct TransitionRunner {
    std::vector<w128> slots, consts, stat, buf[2];
    std::vector<AccIns> code[2];        // code[p]: reads state from buf[p], writes the next state to buf[1-p]
    void init(const HostProgram& pr, int R, int n_static) {
        slots.assign(pr.n_slots + 1, w128{0, 0}); stat.assign(n_static > 0 ? n_static : 1, w128{0, 0});
        buf[0].assign(R, w128{0, 0}); buf[1].assign(R, w128{0, 0}); consts.resize(pr.consts.size());
        for (size_t i = 0; i < pr.consts.size(); ++i) consts[i] = w_from(pr.consts[i]);
        const size_t n = pr.instrs.size();
        // SSA-level use information: def id of each slot at each point, and for each def its later readers
        std::vector<int> def_of_slot(pr.n_slots + 1, -1);
        std::vector<int> da(n, -1), db(n, -1);               // defining instr of operands
        std::vector<std::vector<int>> users(n);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t op = pr.instrs[i][0];
            const bool bin = (op == OP_ADD || op == OP_SUB || op == OP_MUL), un = (op == OP_NEG || op == OP_INV || op == OP_EXP || op == OP_OUT);
            if (bin) { da[i] = def_of_slot[pr.instrs[i][2]]; db[i] = def_of_slot[pr.instrs[i][3]]; }
            if (un) da[i] = def_of_slot[pr.instrs[i][2]];
            if (da[i] >= 0) users[da[i]].push_back((int)i);
            if (db[i] >= 0 && db[i] != da[i]) users[db[i]].push_back((int)i);
            if (op != OP_OUT) def_of_slot[pr.instrs[i][1]] = (int)i;
        }
        for (int p = 0; p < 2; ++p) {
            std::vector<w128>& cur = buf[p]; std::vector<w128>& nxt = buf[1 - p];
            std::vector<AccIns>& out = code[p]; out.clear();
            std::vector<const w128*> where(n, nullptr);   // memory location of each def's value (null = only in acc)
            int acc_def = -1;                             // def currently held in the accumulator
            auto is_leaf = [&](uint32_t op) { return op == OP_CONST || op == OP_CUR || op == OP_STATIC; };
            auto next_arith = [&](size_t i) { for (size_t j = i + 1; j < n; ++j) if (!is_leaf(pr.instrs[j][0])) return (int)j; return -1; };
            for (size_t i = 0; i < n; ++i) {
                const uint32_t op = pr.instrs[i][0], d = pr.instrs[i][1], a = pr.instrs[i][2], b = pr.instrs[i][3];
                if (op == OP_CONST) { where[i] = &consts[a]; continue; }
                if (op == OP_CUR) { where[i] = &cur[a]; continue; }
                if (op == OP_STATIC) { where[i] = &stat[a]; continue; }
                if (op == OP_OUT) {
                    if (da[i] == acc_def && acc_def >= 0) out.push_back({A_STORE, &nxt[d], nullptr, 0});
                    else out.push_back({A_MOVE, &nxt[d], where[da[i]], 0});
                    continue;
                }
                // arithmetic: get the first operand into the accumulator
                const bool bin = (op == OP_ADD || op == OP_SUB || op == OP_MUL);
                int other = -1; bool reversed = false;
                if (bin && da[i] == acc_def && acc_def >= 0) other = db[i];
                else if (bin && db[i] == acc_def && acc_def >= 0) { other = da[i]; reversed = true; }
                else if (!bin && da[i] == acc_def && acc_def >= 0) { /* unary on acc */ }
                else { out.push_back({A_LOAD, nullptr, where[da[i]], 0}); if (bin) other = db[i]; }
                switch (op) {
                    case OP_ADD:
                        if (da[i] == db[i]) out.push_back({A_DBL, nullptr, nullptr, 0});
                        else out.push_back({A_ADD, nullptr, where[other], 0});
                        break;
                    case OP_MUL:
                        if (da[i] == db[i]) out.push_back({A_SQR, nullptr, nullptr, 0});
                        else out.push_back({A_MUL, nullptr, where[other], 0});
                        break;
                    case OP_SUB:
                        if (da[i] == db[i]) { out.push_back({A_LOAD, nullptr, &slots[pr.n_slots], 0}); }   // x - x = 0 (spare slot is zero)
                        else out.push_back({reversed ? (uint32_t)A_RSUB : (uint32_t)A_SUB, nullptr, where[other], 0});
                        break;
                    case OP_NEG: out.push_back({A_NEG, nullptr, nullptr, 0}); break;
                    case OP_INV: out.push_back({A_INV, nullptr, nullptr, 0}); break;
                    case OP_EXP: out.push_back({A_EXP, nullptr, nullptr, pr.consts[b]}); break;
                    default: break;
                }
                acc_def = (int)i;
                // keep a memory copy unless the only reader is the next arithmetic op (via the accumulator) or OUTs right after
                bool need_store = false;
                const int na = next_arith(i);
                for (int u : users[i]) {
                    if (u == na && pr.instrs[u][0] != OP_OUT) {
                        // the next op takes it from the accumulator; but x*x style double use is fine, and a binary op
                        // whose BOTH operands are defs other than this one cannot happen here
                        continue;
                    }
                    if (pr.instrs[u][0] == OP_OUT && u == na) continue;     // stored straight from the accumulator
                    need_store = true;
                }
                // an OUT that is not immediately next still needs the value in memory
                if (need_store) { out.push_back({A_STORE, &slots[d], nullptr, 0}); where[i] = &slots[d]; }
                else where[i] = nullptr;
            }
        }
    }
    // Threaded dispatch (one indirect branch per opcode site)
    // so the branch predictor sees a separate history for every position in the program.
    inline void step(int p) {
        w128 acc{0, 0};
        const AccIns* i = code[p].data();
        const AccIns* const end = i + code[p].size();
#if defined(__GNUC__)
        static const void* const tbl[] = {&&L_LOAD, &&L_STORE, &&L_ADD, &&L_SUB, &&L_RSUB, &&L_MUL, &&L_SQR, &&L_NEG, &&L_INV, &&L_EXP, &&L_MOVE, &&L_DBL};
#define GS_NEXT if (++i == end) return; goto *(void*)tbl[i->op]
        if (i == end) return;
        goto *(void*)tbl[i->op];
        L_LOAD: acc = *i->src; GS_NEXT;
        L_STORE: *i->dst = acc; GS_NEXT;
        L_MOVE: *i->dst = *i->src; GS_NEXT;
        L_ADD: acc = w_add(acc, *i->src); GS_NEXT;
        L_DBL: acc = w_add(acc, acc); GS_NEXT;
        L_SUB: acc = w_sub(acc, *i->src); GS_NEXT;
        L_RSUB: acc = w_sub(*i->src, acc); GS_NEXT;
        L_MUL: acc = w_mul(acc, *i->src); GS_NEXT;
        L_SQR: acc = w_mul(acc, acc); GS_NEXT;
        L_NEG: acc = w_sub(w128{0, 0}, acc); GS_NEXT;
        L_INV: acc = w_from(h_inv(w_canon(acc))); GS_NEXT;
        L_EXP: acc = w_from(h_pow(w_canon(acc), i->e)); GS_NEXT;
#undef GS_NEXT
#else
        for (; i != end; ++i) {
            switch (i->op) {
                case A_LOAD: acc = *i->src; break;
                case A_STORE: *i->dst = acc; break;
                case A_MOVE: *i->dst = *i->src; break;
                case A_ADD: acc = w_add(acc, *i->src); break;
                case A_DBL: acc = w_add(acc, acc); break;
                case A_SUB: acc = w_sub(acc, *i->src); break;
                case A_RSUB: acc = w_sub(*i->src, acc); break;
                case A_MUL: acc = w_mul(acc, *i->src); break;
                case A_SQR: acc = w_mul(acc, acc); break;
                case A_NEG: acc = w_sub(w128{0, 0}, acc); break;
                case A_INV: acc = w_from(h_inv(w_canon(acc))); break;
                case A_EXP: acc = w_from(h_pow(w_canon(acc), i->e)); break;
                default: break;
            }
        }
#endif
    }
};


// generateExecutionTrace (lib/Stark.ts:97): R x T, row = register; canonical residues
// on_chunk(first_step, end_step) is called whenever another block of steps is final (the prover starts their
// host->device copy while the next block is being generated)
typedef std::function<void(long long, long long)> TraceChunkFn;
void generate_trace(const AirHost* S, const u128* init_state, const fp* input_traces, fp* tr, const TraceChunkFn* on_chunk = nullptr);
// (implementation: hostjit.h -- compiled transition function, or the interpreter above)
const char* trace_backend_status();
// compile (or load from the cache) the native transition function ahead of the first prove
void trace_prepare(const AirHost* S);

}  // namespace gs
