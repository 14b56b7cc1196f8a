// K2: constraint evaluation + composition polynomial C(x) + random linear combination L(x), fused.
//
// Reference data flow (lib/components/CompositionPolynomial.ts:71-146, LinearCombination.ts:36-64):
//   Q = constraints over the composition domain (M points) -> degree adjust -> sum d_i Q_i
//     -> iNTT(M) -> NTT(N) -> * Z^-1 -> + boundary part -> C ;  L = C + sum kappa_j V_j (+ x^D copies)
// Every term of the combined constraint polynomial has degree < M (term i: d_i (T-1) + (M - d_i T)),
// so interpolating it on M points and evaluating on N points returns exactly its values on the N
// points; those values are obtained here by evaluating the constraints directly at the evaluation
// points (next state = position i + E, same coset).  Exact arithmetic => identical field elements; the
// oracle keeps the literal detour and tests/ compare the two.
//
// One thread per evaluation point, 16-byte coalesced loads of the trace / secret columns; the AIR
// evaluation function is a flat register-machine program (genstark_b200/air.py) interpreted with the
// instruction stream read uniformly by the whole warp.
#pragma once
#include "core.cuh"
#include "ntt.cuh"
#include "hostair.h"

namespace gs {

#define GS_MAX_POWERS 8

struct ComposeParams {
    long long n;                 // evaluation domain size N (global)
    int log_n, log_e;            // N = 2^log_n, E = 2^log_e
    // coset sharding: this launch covers the n_loc positions of cosets [j0, j0 + 2^log_el); local index
    // i_loc = q * 2^log_el + (j - j0) for global position i = q * E + j.  Single GPU: log_el = log_e, j0 = 0.
    long long n_loc; int log_el, j0;
    // program
    const uint4* instrs; int n_instr; const fp* consts; int n_slots;
    // trace columns (LDE over N)
    int n_trace; const fp* trace[GS_MAX_COLS];
    // static registers: kind 0 = cyclic table (index i & mask), kind 1 = full column over N
    int n_static; const fp* stat[GS_MAX_COLS]; unsigned stat_mask[GS_MAX_COLS];
    // transition part: q_k * (dk[k] + dk_adj[k] * x^incr[pow_idx[k]])
    int n_constraints;
    const fp* dk; const fp* dk_adj; const int* pow_idx;      // device arrays [K]
    // every degree increment is a multiple of T (comb = cT, group degrees d*T, delta = compDeg - T), so
    // x^incr = w^(i*incr) depends on i mod E only: tables of E entries, pow_tab[g*E + (i mod E)]
    int n_powers; const fp* pow_tab; const fp* delta_tab;
    // zero polynomial: D = qc * (x - x_last) * inv_num[i mod E]
    fp x_last; const fp* inv_num;
    // boundary part: for asserted register slot b: (P_reg - I(x)) / Z_b(x) * (bk[b] + bk_adj[b] * x^delta).
    // 1/Z_b(x_i) by partial fractions over the per-context table u[j] = 1/(w^j - 1):
    //   1/prod_k (x - X_k) = sum_k c_k / (x - X_k),  1/(w^i - w^s) = w^-s * u[(i - s) mod N],
    // so each assertion costs one coalesced table read and one modmul; at x = X_k the reference's
    // inv(0) = 0 convention makes the whole term zero.
    int n_boundary; const int* b_reg; const int* b_ipoly_off; const int* b_ipoly_len; const fp* b_ipoly;
    const int* b_pf_off; const int* b_pf_len; const fp* b_pf_coef; const unsigned* b_pf_shift;   // shift = step * E
    const fp* u_table;           // [N]
    const fp* bk; const fp* bk_adj;
    // linear combination: columns V = trace then secret; L = C + sum V_j (lk[j] + lk_adj[j] * x^delta)
    int n_lc; const fp* lc_col[GS_MAX_COLS]; const fp* lk; const fp* lk_adj;
    unsigned long long delta;    // compositionDegree - T (0 => no adjusted copies)
    // roots
    const fp* tw_lo; const fp* tw_hi; int log_g, log_lo;
    fp* out;                     // L(x) over N
    fp* c_out;                   // optional: C(x) over N (stage-level parity tests); may be null
    int* fail_flag;              // min over violations of (step << 6 | constraint); INT_MAX when the trace satisfies the AIR
};

GS_D fp root_pow(const ComposeParams& P, unsigned long long e_n) {
    // w_N^e for e < N, from the w_G tables
    const unsigned e = (unsigned)(e_n << (P.log_g - P.log_n));
    fp lo = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
    if (P.log_g <= P.log_lo) return lo;
    return fp_mul(lo, ldg_fp(P.tw_hi + (e >> P.log_lo)));
}

template <int NSLOT>
__global__ void __launch_bounds__(256) compose_kernel(const ComposeParams* __restrict__ Pp) {
    const ComposeParams& P = *Pp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned long long lmask = (unsigned long long)P.n_loc - 1ull;
    const unsigned E = 1u << P.log_e, EL = 1u << P.log_el;
    for (long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x; il < P.n_loc; il += stride) {
        // all per-position data (trace, inputs, u table, outputs) is indexed locally; the domain point is global
        const long long i = ((il >> P.log_el) << P.log_e) + P.j0 + (il & (EL - 1));
        const long long inext = (il + EL) & (long long)lmask;
        fp slot[NSLOT];
        // powers of x used by this point
        const fp x = root_pow(P, (unsigned long long)i);
        const unsigned ie = (unsigned)i & (E - 1);
        fp xpow[GS_MAX_POWERS];
#pragma unroll 1
        for (int g = 0; g < P.n_powers; ++g) xpow[g] = ldg_fp(P.pow_tab + (g << P.log_e) + ie);
        fp xdelta = fp_one();
        if (P.delta) xdelta = ldg_fp(P.delta_tab + ie);

        // ---- transition constraints, combined on the fly
        fp acc = fp_zero();
#pragma unroll 1
        for (int pc = 0; pc < P.n_instr; ++pc) {
            const uint4 ins = __ldg(P.instrs + pc);
            const unsigned op = ins.x, d = ins.y, a = ins.z, b = ins.w;
            switch (op) {
                case OP_CONST: slot[d] = ldg_fp(P.consts + a); break;
                case OP_CUR: slot[d] = ld_fp(P.trace[a] + il); break;
                case OP_NEXT: slot[d] = ld_fp(P.trace[a] + inext); break;
                case OP_STATIC: slot[d] = ld_fp(P.stat[a] + (P.stat_mask[a] == 0xFFFFFFFFu ? (unsigned long long)il : ((unsigned long long)i & P.stat_mask[a]))); break;
                case OP_ADD: slot[d] = fp_add(slot[a], slot[b]); break;
                case OP_SUB: slot[d] = fp_sub(slot[a], slot[b]); break;
                case OP_MUL: slot[d] = fp_mul(slot[a], slot[b]); break;
                case OP_NEG: slot[d] = fp_neg(slot[a]); break;
                case OP_INV: slot[d] = fp_inv(slot[a]); break;
                case OP_OUT: {
                    const fp qv = slot[a];
                    // the reference (air-assembly) refuses a trace that violates a constraint: positions that
                    // are trace steps (i % E == 0) other than the last step must evaluate to zero
                    // (the first violating step, then the lowest constraint index, as the sequential reference reports)
                    if (((unsigned)i & (E - 1)) == 0 && i < P.n - E && !fp_is_zero(qv))
                        atomicMin(P.fail_flag, (int)(((unsigned)(i >> P.log_e) << 6) | d));
                    fp coef = ldg_fp(P.dk + d);
                    const int pi = __ldg(P.pow_idx + d);
                    if (pi >= 0) coef = fp_add(coef, fp_mul(ldg_fp(P.dk_adj + d), xpow[pi]));
                    acc = fp_add(acc, fp_mul(qv, coef));
                    break;
                }
                default: break;
            }
        }
        // ---- D(x) = Q(x) / Z(x) = Q * (x - x_last) * inv(x^T - 1)
        fp c = fp_mul(fp_mul(acc, fp_sub(x, P.x_last)), ldg_fp(P.inv_num + ((unsigned)i & (E - 1))));
        // ---- boundary constraints
#pragma unroll 1
        for (int bi = 0; bi < P.n_boundary; ++bi) {
            const int off = P.b_ipoly_off[bi], len = P.b_ipoly_len[bi];
            fp iv = ldg_fp(P.b_ipoly + off + len - 1);
            for (int k = len - 2; k >= 0; --k) iv = fp_add(fp_mul(iv, x), ldg_fp(P.b_ipoly + off + k));
            const fp pv = ld_fp(P.trace[P.b_reg[bi]] + il);
            fp zinv = fp_zero();
            bool at_root = false;
            const int po = P.b_pf_off[bi], pl = P.b_pf_len[bi];
            for (int k = 0; k < pl; ++k) {
                const unsigned sh = P.b_pf_shift[po + k];
                at_root |= ((unsigned)i == sh);
                // (i - sh) mod N stays in the same coset: locally it is a shift by (sh / E) rows
                const unsigned long long j = ((unsigned long long)il - ((unsigned long long)(sh >> P.log_e) << P.log_el)) & lmask;
                zinv = fp_add(zinv, fp_mul(ldg_fp(P.b_pf_coef + po + k), ld_fp(P.u_table + j)));
            }
            if (at_root) zinv = fp_zero();
            const fp bv = fp_mul(fp_sub(pv, iv), zinv);
            fp coef = ldg_fp(P.bk + bi);
            if (P.delta) coef = fp_add(coef, fp_mul(ldg_fp(P.bk_adj + bi), xdelta));
            c = fp_add(c, fp_mul(bv, coef));
        }
        if (P.c_out) st_fp(P.c_out + il, c);
        // ---- linear combination with P(x) and S(x)
        fp l = c;
#pragma unroll 1
        for (int j = 0; j < P.n_lc; ++j) {
            fp coef = ldg_fp(P.lk + j);
            if (P.delta) coef = fp_add(coef, fp_mul(ldg_fp(P.lk_adj + j), xdelta));
            l = fp_add(l, fp_mul(ld_fp(P.lc_col[j] + il), coef));
        }
        st_fp(P.out + il, l);
    }
}

// u[j] = w_N^j - 1 (inverted afterwards by K3; u[0] stays 0)
struct UTableParams { long long n_loc; int log_n, log_e, log_el, j0; const fp* tw_lo; const fp* tw_hi; int log_g, log_lo; fp* out; };
__global__ void __launch_bounds__(256) u_table_kernel(const UTableParams P) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x; il < P.n_loc; il += stride) {
        const long long i = ((il >> P.log_el) << P.log_e) + P.j0 + (il & ((1ll << P.log_el) - 1));
        const unsigned e = (unsigned)((unsigned long long)i << (P.log_g - P.log_n));
        fp x = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
        if (P.log_g > P.log_lo) x = fp_mul(x, ldg_fp(P.tw_hi + (e >> P.log_lo)));
        st_fp(P.out + il, fp_sub(x, fp_one()));
    }
}

}  // namespace gs
