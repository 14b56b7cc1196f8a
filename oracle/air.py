"""ORACLE (test infrastructure) -- AIR runtime: proving / verification contexts.

Restates what genSTARK uses from ``@guildofweavers/air-assembly 0.3.6`` (pinned at
/root/reference/package-lock.json:12-20, NOT vendored): ``initProvingContext`` (Stark.ts:90),
``generateExecutionTrace`` (Stark.ts:97), ``evaluateTransitionConstraints``
(CompositionPolynomial.ts:76), ``initVerificationContext`` (Stark.ts:177) and
``evaluateConstraintsAt`` (CompositionPolynomial.ts:153).  Internals are the package's published
behaviour [RECALLED, SURVEY §3.2 / App. A.1 / App. C]; the in-tree anchors are the domain
conventions the verifier relies on (Stark.ts:222-232, BoundaryConstraints.ts:23,
ZeroPolynomial.ts:21-23, CompositionPolynomial.ts:84-85).  PARITY UNPINNED.

The AIR itself comes in as the flat IR of ``genstark_b200/air.py`` (pure data); this file has its
own interpreter for it.
"""
from __future__ import annotations

from typing import List, Optional

from genstark_b200 import air as ir
from .field import PrimeField, Vector, Matrix


def run_program(f: PrimeField, prog: ir.Program, cur, nxt, statics) -> List[int]:
    p = f.modulus
    s = [0] * prog.n_slots
    out = [0] * prog.n_out
    for op, d, a, b in prog.instrs:
        if op == ir.OP_CONST: s[d] = prog.consts[a]
        elif op == ir.OP_CUR: s[d] = cur[a]
        elif op == ir.OP_NEXT: s[d] = nxt[a]
        elif op == ir.OP_STATIC: s[d] = statics[a]
        elif op == ir.OP_ADD: s[d] = (s[a] + s[b]) % p
        elif op == ir.OP_SUB: s[d] = (s[a] - s[b]) % p
        elif op == ir.OP_MUL: s[d] = (s[a] * s[b]) % p
        elif op == ir.OP_NEG: s[d] = (-s[a]) % p
        elif op == ir.OP_INV: s[d] = f.inv(s[a])
        elif op == ir.OP_EXP: s[d] = pow(s[a], prog.consts[b], p)
        elif op == ir.OP_OUT: out[d] = s[a]
        else: raise ValueError(op)
    return out


class ConstraintDescriptor:
    def __init__(self, degree: int):
        self.degree = degree


class AirContext:
    """Common part of the proving and verification contexts."""

    def __init__(self, air: ir.AirModule):
        self.air = air
        self.field = PrimeField(air.modulus)
        self.trace_length = air.trace_length
        self.extension_factor = air.extension_factor
        self.constraints = [ConstraintDescriptor(d) for d in air.constraint_degrees]
        n = self.trace_length * self.extension_factor
        self.root_of_unity = self.field.get_root_of_unity(n)
        max_deg = max(air.constraint_degrees)
        self.composition_factor = 1
        while self.composition_factor < max_deg:
            self.composition_factor *= 2
        # cyclic registers as polynomials over the subgroup of order len(values)
        f = self.field
        self.cycle_polys = {}
        for k, reg in enumerate(air.static_registers):
            if reg.kind == 'cycle':
                L = len(reg.values)
                g = f.exp(self.root_of_unity, n // L)
                self.cycle_polys[k] = f.interpolate_roots(f.get_power_series(g, L), list(reg.values))


class ProvingContext(AirContext):
    def __init__(self, air: ir.AirModule, inputs, seed):
        super().__init__(air)
        f = self.field
        T, E = self.trace_length, self.extension_factor
        N = T * E
        self.inputs, self.seed = inputs or [], seed or []
        self.evaluation_domain = f.get_power_series(self.root_of_unity, N)
        self.execution_domain = f.pluck_vector(self.evaluation_domain, E, T)
        M = T * self.composition_factor
        self.composition_domain = f.pluck_vector(self.evaluation_domain, N // M, M)
        self.input_shapes = air.input_shapes(self.inputs)
        # input registers: T-length traces -> polynomials; secret ones are also extended over the
        # evaluation domain and committed next to P(x) (Stark.ts:113-115)
        self.input_traces = air.expand_inputs(self.inputs)
        self.input_polys = [f.interpolate_roots(self.execution_domain, list(t)) for t in self.input_traces]
        self.secret_register_traces: List[Vector] = []
        it = iter(self.input_polys)
        self._static_source = []          # per static register: ('cycle', k) | ('input', poly)
        for k, reg in enumerate(air.static_registers):
            if reg.kind == 'cycle':
                self._static_source.append(('cycle', k))
            else:
                poly = next(it)
                self._static_source.append(('input', poly))
                if reg.secret:
                    self.secret_register_traces.append(f.eval_poly_at_roots(poly, self.evaluation_domain))

    def static_trace_values(self, step: int) -> List[int]:
        vals = []
        it = iter(self.input_traces)
        for reg in self.air.static_registers:
            if reg.kind == 'cycle':
                vals.append(reg.values[step % len(reg.values)])
            else:
                vals.append(next(it)[step])
        return vals

    def generate_execution_trace(self) -> Matrix:
        """R x T matrix, row = register (Stark.ts:97,106)."""
        f, air = self.field, self.air
        T, R = self.trace_length, air.trace_register_count
        state = list(air.init(self.inputs, self.seed))
        assert len(state) == R
        trace = [[0] * T for _ in range(R)]
        for step in range(T):
            for r in range(R):
                trace[r][step] = state[r]
            if step + 1 < T:
                state = run_program(f, air.transition, state, None, self.static_trace_values(step))
        return trace

    def generate_static_trace(self) -> Matrix:
        T = self.trace_length
        cols = [self.static_trace_values(s) for s in range(T)]
        return [list(r) for r in zip(*cols)] if cols and cols[0] else []

    def evaluate_transition_constraints(self, p_polys: Matrix) -> Matrix:
        """K x M matrix of constraint evaluations over the composition domain; throws when a
        constraint is non-zero on a trace step."""
        f, air = self.field, self.air
        M = len(self.composition_domain)
        c = M // self.trace_length
        t_evals = f.eval_polys_at_roots(p_polys, self.composition_domain)
        statics = []
        for kind, src in self._static_source:
            if kind == 'cycle':
                L = len(air.static_registers[src].values)
                g = f.exp(self.root_of_unity, (self.trace_length * self.extension_factor) // (L * c))
                ev = f.eval_poly_at_roots(self.cycle_polys[src], f.get_power_series(g, L * c))
                statics.append(ev)
            else:
                statics.append(f.eval_poly_at_roots(src, self.composition_domain))
        K = air.constraint_count
        R = air.trace_register_count
        out = [[0] * M for _ in range(K)]
        nf_steps = M - c
        for pos in range(M):
            cur = [t_evals[r][pos] for r in range(R)]
            nxt = [t_evals[r][(pos + c) % M] for r in range(R)]
            sv = [s[pos % len(s)] for s in statics]
            q = run_program(f, air.evaluation, cur, nxt, sv)
            if pos % c == 0 and pos < nf_steps:
                for k in range(K):
                    if q[k] != 0:
                        raise ValueError(f"Constraint {k} didn't evaluate to 0 at step {pos // c}")
            for k in range(K):
                out[k][pos] = q[k]
        return out


class VerificationContext(AirContext):
    def __init__(self, air: ir.AirModule, input_shapes, public_inputs):
        super().__init__(air)
        f = self.field
        self.input_shapes = input_shapes
        self.public_inputs = public_inputs or []
        T, E = self.trace_length, self.extension_factor
        self._exe_root = f.exp(self.root_of_unity, E)
        # public input registers are recomputed by the verifier
        self.public_polys = {}
        pub_regs = [k for k, r in enumerate(air.static_registers) if r.kind == 'input' and not r.secret]
        if pub_regs:
            traces = air.expand_public_inputs(self.public_inputs)
            dom = f.get_power_series(self._exe_root, T)
            for k, t in zip(pub_regs, traces):
                self.public_polys[k] = f.interpolate_roots(dom, list(t))

    def evaluate_constraints_at(self, x: int, p_values, n_values, h_values) -> List[int]:
        f, air = self.field, self.air
        T = self.trace_length
        sv = []
        h = iter(h_values)
        for k, reg in enumerate(air.static_registers):
            if reg.kind == 'cycle':
                L = len(reg.values)
                sv.append(f.eval_poly_at(self.cycle_polys[k], f.exp(x, T // L)))
            elif reg.secret:
                sv.append(next(h))
            else:
                sv.append(f.eval_poly_at(self.public_polys[k], x))
        return run_program(f, air.evaluation, p_values, n_values, sv)
