"""Hand-built AIRs for the workloads BASELINE.json names (SURVEY.md §8 config legend).

The reference compiles these from AirScript / AirAssembly text with packages that are not in the
reference tree; the front-end is out of scope, so each AIR is written directly in
the IR of ``air.py`` with the reference source it restates cited beside it.
"""
from __future__ import annotations

from typing import List

from .air import (AirModule, ProgramBuilder, StaticRegister, P128, P32, prng_sha256)

MIMC_SEED = bytes.fromhex('4d694d43')       # examples/mimc/mimc128.ts:15,36


def mimc_round_constants(count: int = 64, modulus: int = P128) -> List[int]:
    return prng_sha256(MIMC_SEED, count, modulus)


def mimc128(steps: int = 2**13, constant_count: int = 64, extension_factor: int = None) -> AirModule:
    """MiMC over p128: x_{s+1} = x_s^3 + k[s mod 64]; 1 register, 1 constraint of degree 3.
    Restates the AirAssembly module at examples/mimc/mimc128Assembly.ts:28-51 (no input registers:
    the start value comes in through ``seed``, mimc128Assembly.ts:66)."""
    p = P128
    t = ProgramBuilder(p)
    t.out(0, t.exp(t.cur(0), 3) + t.static(0))
    e = ProgramBuilder(p)
    e.out(0, e.nxt(0) - (e.exp(e.cur(0), 3) + e.static(0)))
    return AirModule(
        name='mimc', modulus=p, trace_register_count=1, trace_length=steps,
        transition=t.build(), evaluation=e.build(),
        static_registers=[StaticRegister('cycle', mimc_round_constants(constant_count, p))],
        extension_factor=extension_factor,
        init=lambda inputs, seed: [int(seed[0]) % p])


def run_mimc(steps: int, round_constants: List[int], seed: int, modulus: int = P128) -> List[int]:
    """Control values, examples/mimc/utils.ts:7-14."""
    out = [seed % modulus]
    for i in range(steps - 1):
        out.append((pow(out[i], 3, modulus) + round_constants[i % len(round_constants)]) % modulus)
    return out


def foo(steps: int = 64, extension_factor: int = None) -> AirModule:
    """README.md:22-39: x_{n+1} = x_n + 2 over p32 with one secret input register (startValue held
    for the whole segment).  CPU-only plumbing config (BASELINE.json configs[0])."""
    p = P32
    t = ProgramBuilder(p)
    t.out(0, t.cur(0) + 2)
    e = ProgramBuilder(p)
    e.out(0, e.nxt(0) - (e.cur(0) + 2))
    return AirModule(
        name='foo', modulus=p, trace_register_count=1, trace_length=steps,
        transition=t.build(), evaluation=e.build(),
        static_registers=[StaticRegister('input', secret=True)],
        extension_factor=extension_factor,
        init=lambda inputs, seed: [int(inputs[0][0]) % p],
        expand_inputs=lambda inputs: [[int(inputs[0][0]) % p] * steps],
        input_shapes=lambda inputs: [[1]])


def fibonacci(steps: int = 64, modulus: int = P32, extension_factor: int = None) -> AirModule:
    """examples/demo/fibonacci.ts:13-34: 2 registers, a0' = a0+a1, a1' = a1 + (a0+a1)."""
    p = modulus
    t = ProgramBuilder(p)
    a2 = t.cur(0) + t.cur(1)
    t.out(0, a2)
    t.out(1, t.cur(1) + a2)
    e = ProgramBuilder(p)
    a2 = e.cur(0) + e.cur(1)
    e.out(0, e.nxt(0) - a2)
    e.out(1, e.nxt(1) - (e.cur(1) + a2))
    return AirModule(
        name='fibonacci', modulus=p, trace_register_count=2, trace_length=steps,
        transition=t.build(), evaluation=e.build(), extension_factor=extension_factor,
        init=lambda inputs, seed: [int(seed[0]) % p, int(seed[1]) % p])
