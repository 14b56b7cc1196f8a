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


// ComposeParams and the per-point pieces below exist twice from one spelling (fp128.cuh: GS_DUAL_SOURCE): compiled
// here for the interpreting kernel, and as text for the kernel devjit.cuh generates for an AIR's constraints.
GS_DUAL_SOURCE(GS_COMPOSE_PARAMS_SRC,
struct ComposeParams {
    long long n;
    int log_n; int log_e;
    long long n_loc; int log_el; int j0;
    const uint4* instrs; int n_instr; const fp* consts; int n_slots;
    int n_trace; const fp* trace[GS_MAX_COLS];
    int n_static; const fp* stat[GS_MAX_COLS]; unsigned stat_mask[GS_MAX_COLS];
    int n_constraints; const fp* cd_tab;
    fp x_last;
    int n_boundary; const int* b_reg; const int* b_ipoly_off; const int* b_ipoly_len; const fp* b_ipoly;
    const int* b_pf_off; const int* b_pf_len; const fp* pf_tab; const unsigned* b_pf_shift;
    const fp* u_table;
    int n_lc; const fp* lc_col[GS_MAX_COLS]; const fp* lk_tab;
    const fp* tw_lo; const fp* tw_hi; int log_g; int log_lo;
    fp* out;
    fp* c_out;
    int* fail_flag;
};
)
// Fields of ComposeParams:
//   n, log_n, log_e        evaluation domain size N = 2^log_n (global), E = 2^log_e
//   n_loc, log_el, j0      coset sharding: this launch covers the n_loc positions of cosets [j0, j0 + 2^log_el); local
//                          index i_loc = q * 2^log_el + (j - j0) for global position i = q * E + j (1 GPU: log_el = log_e)
//   instrs, consts         the AIR evaluation function as a flat register-machine program (genstark_b200/air.py)
//   trace[]                trace columns (LDE over N)
//   stat[], stat_mask[]    static registers: cyclic table (index i & mask) or a full column over N (mask = ~0)
//   Every degree increment is a multiple of T (comb = cT, group degrees d*T, delta = compDeg - T), so
//   x^incr = w^(i*incr) depends on i mod E only, and so does 1/(x^T - 1).  The host folds the random coefficients
//   with those E-periodic factors into E-entry tables once per proof (prover.cuh):
//   cd_tab    transition:  D(x) = (x - x_last) * sum_k q_k * cd_tab[k*E + (i mod E)],
//             cd_tab = (d_k + d'_k * x^incr_k) * inv(x^T - 1)                    (CompositionPolynomial.ts:84-120)
//   pf_tab    boundary:  sum_b (P_reg - I_b(x)) * sum_a pf_tab[(off_b + a)*E + (i mod E)] * u[(i - s_a E) mod N],
//             pf_tab = c_a * X_a^-1 * (b_b + b'_b * x^delta); 1/Z_b by partial fractions over the per-context table
//             u[j] = 1/(w^j - 1):  1/prod_a (x - X_a) = sum_a c_a / (x - X_a),  1/(w^i - w^s) = w^-s * u[(i - s) mod N];
//             at x = X_a the reference's inv(0) = 0 convention makes the whole term zero
//             (BoundaryConstraints.ts:71-95); b_pf_shift = step * E
//   lk_tab    combination: L = C + sum_j V_j * lk_tab[j*E + (i mod E)],  lk_tab = kappa_j + kappa'_j * x^delta
//             (LinearCombination.ts:36-64); lc_col = trace then secret columns
//   out, c_out             L(x) over N; optional C(x) over N (stage-level parity tests)
//   fail_flag              min over violations of (step << 6 | constraint); INT_MAX when the trace satisfies the AIR

GS_DEVICE_DUAL_SOURCE(GS_COMPOSE_DEVICE_SRC,
// w_N^e for e < N, from the w_G tables
GS_D fp root_pow(const ComposeParams& P, unsigned long long e_n) {
    const unsigned e = (unsigned)(e_n << (P.log_g - P.log_n));
    fp lo = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
    if (P.log_g <= P.log_lo) return lo;
    return d_mul(lo, ldg_fp(P.tw_hi + (e >> P.log_lo)));
}

// constraint k evaluated to qv at global position i: the violation check of the reference and the running sum
GS_D fp compose_out(const ComposeParams& P, long long i, unsigned ie, unsigned k, const fp& qv, const fp& acc) {
    // the reference (air-assembly) refuses a trace that violates a constraint: positions that are trace steps
    // (i % E == 0) other than the last step must evaluate to zero (the first violating step, then the lowest
    // constraint index, as the sequential reference reports)
    if (ie == 0 && i < P.n - (1ll << P.log_e) && !fp_is_zero(qv))
        atomicMin(P.fail_flag, (int)(((unsigned)(i >> P.log_e) << 6) | k));
    return d_add(acc, d_mul(qv, ldg_fp(P.cd_tab + (k << P.log_e) + ie)));
}

// everything after the constraints: D(x), the boundary part, C(x), and the linear combination L(x)
GS_D void compose_tail(const ComposeParams& P, long long il, long long i, unsigned ie, const fp& x, const fp& acc) {
    const unsigned long long lmask = (unsigned long long)P.n_loc - 1ull;
    // D(x) = Q(x) / Z(x) = Q * (x - x_last) * inv(x^T - 1), the last factor already inside cd_tab
    fp c = d_mul(acc, d_sub(x, P.x_last));
    _Pragma("unroll 1")
    for (int bi = 0; bi < P.n_boundary; ++bi) {
        const int off = P.b_ipoly_off[bi]; const int len = P.b_ipoly_len[bi];
        fp iv = ldg_fp(P.b_ipoly + off + len - 1);
        for (int k = len - 2; k >= 0; --k) iv = d_add(d_mul(iv, x), ldg_fp(P.b_ipoly + off + k));
        const fp pv = ld_fp(P.trace[P.b_reg[bi]] + il);
        fp zinv = fp_zero();
        bool at_root = false;
        const int po = P.b_pf_off[bi]; const int pl = P.b_pf_len[bi];
        for (int k = 0; k < pl; ++k) {
            const unsigned sh = P.b_pf_shift[po + k];
            at_root |= ((unsigned)i == sh);
            // (i - sh) mod N stays in the same coset: locally it is a shift by (sh / E) rows
            const unsigned long long j = ((unsigned long long)il - ((unsigned long long)(sh >> P.log_e) << P.log_el)) & lmask;
            zinv = d_add(zinv, d_mul(ldg_fp(P.pf_tab + ((po + k) << P.log_e) + ie), ld_fp(P.u_table + j)));
        }
        if (!at_root) c = d_add(c, d_mul(d_sub(pv, iv), zinv));
    }
    if (P.c_out) st_fp(P.c_out + il, c);
    fp l = c;
    _Pragma("unroll 1")
    for (int j = 0; j < P.n_lc; ++j)
        l = d_add(l, d_mul(ld_fp(P.lc_col[j] + il), ldg_fp(P.lk_tab + (j << P.log_e) + ie)));
    st_fp(P.out + il, l);
}
)

template <int NSLOT>
__global__ void __launch_bounds__(256) compose_kernel(const ComposeParams* __restrict__ Pp) {
#ifdef __CUDA_ARCH__
    const ComposeParams& P = *Pp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned long long lmask = (unsigned long long)P.n_loc - 1ull;
    const unsigned E = 1u << P.log_e, EL = 1u << P.log_el;
    for (long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x; il < P.n_loc; il += stride) {
        // all per-position data (trace, inputs, u table, outputs) is indexed locally; the domain point is global
        const long long i = ((il >> P.log_el) << P.log_e) + P.j0 + (il & (EL - 1));
        const long long inext = (il + EL) & (long long)lmask;
        fp slot[NSLOT];
        const fp x = root_pow(P, (unsigned long long)i);
        const unsigned ie = (unsigned)i & (E - 1);
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
                case OP_OUT: acc = compose_out(P, i, ie, d, slot[a], acc); break;
                default: break;
            }
        }
        compose_tail(P, il, i, ie, x, acc);
    }
#endif
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
