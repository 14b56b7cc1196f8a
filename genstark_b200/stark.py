"""Host-side mirror of genSTARK's public API on top of libgenstark_b200.so.

``Stark`` keeps the interface of /root/reference/lib/Stark.ts (genstark.d.ts:85-124): ``prove`` /
``verify`` / ``serialize`` / ``parse`` / ``sizeOf`` / ``securityLevel``; ``instantiate`` mirrors
index.ts:18-33 with the AIR given as an ``AirModule`` (the AirScript / AirAssembly compilers are out
of scope).  The prover body runs on the GPU through ``gs_stark_prove``; there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import struct
from typing import Dict, List, Optional, Sequence

from . import _native
from .air import AirModule, input_blob, input_cbuf, pack_air, public_blob
from .field import Context

DEFAULT_EXE_QUERY_COUNT, DEFAULT_FRI_QUERY_COUNT = 80, 40          # Stark.ts:13-14
MAX_EXE_QUERY_COUNT, MAX_FRI_QUERY_COUNT = 128, 64                 # Stark.ts:16-17
HASH_ALGORITHMS = ['sha256', 'blake2s256']                         # Stark.ts:19
DEFAULT_HASH_ALGORITHM = 'sha256'
MAX_ARRAY_LENGTH, MAX_MATRIX_COLUMN_LENGTH = 256, 127              # lib/utils/sizeof.ts:7-8


class StarkError(Exception):
    """lib/StarkError.ts:3-13"""


class BatchMerkleProof:
    """@guildofweavers/merkle BatchMerkleProof {values, nodes, depth} (lib/utils/serialization.ts:31-35)"""

    def __init__(self, values: List[bytes], nodes: List[List[bytes]], depth: int):
        self.values, self.nodes, self.depth = values, nodes, depth


def _pow_log2(base: float, exponent: int) -> float:                # lib/utils/index.ts:23-30
    twos = 0
    while exponent % 2 == 0:
        twos += 1
        exponent //= 2
    return (2 ** twos) * math.log2(base ** exponent)


class Stark:
    def __init__(self, air: AirModule, options: Optional[dict] = None, logger=None, context: Optional[Context] = None):
        options = options or {}
        self.air = air.with_options(options.get('extensionFactor'))
        # buildSecurityOptions (Stark.ts:318-344)
        exe = options.get('exeQueryCount') or DEFAULT_EXE_QUERY_COUNT
        if exe < 1 or exe > MAX_EXE_QUERY_COUNT or int(exe) != exe:
            raise TypeError(f'Execution sample size must be an integer between 1 and {MAX_EXE_QUERY_COUNT}')
        fri = options.get('friQueryCount') or DEFAULT_FRI_QUERY_COUNT
        if fri < 1 or fri > MAX_FRI_QUERY_COUNT or int(fri) != fri:
            raise TypeError(f'FRI sample size must be an integer between 1 and {MAX_FRI_QUERY_COUNT}')
        alg = options.get('hashAlgorithm') or DEFAULT_HASH_ALGORITHM
        if alg not in HASH_ALGORITHMS:
            raise TypeError(f'Hash algorithm {alg} is not supported')
        if not self.air.extension_factor:
            raise TypeError('Extension factor is undefined')
        self.exeQueryCount, self.friQueryCount, self.hashAlgorithm = int(exe), int(fri), alg
        self.digestSize = 32
        self.elementSize = self.air.element_size
        self.logger = logger
        self._lib = _native.lib()
        self.context = context or Context(options.get('device', 0))
        h = C.c_void_p()
        blob = pack_air(self.air)
        rc = self._lib.gs_stark_create(self.context.handle, blob, len(blob), HASH_ALGORITHMS.index(alg),
                                       self.exeQueryCount, self.friQueryCount, C.byref(h))
        self.context.check(rc)
        self._handle = h

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.gs_stark_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ACCESSORS -------------------------------------------------------------------------------------
    @property
    def securityLevel(self) -> int:                                # Stark.ts:62-77
        e = self.air.extension_factor
        es = _pow_log2(e / self.air.max_constraint_degree, self.exeQueryCount)
        fs = math.log2(e) * self.friQueryCount
        hs = self.digestSize * 4
        return math.floor(min(es, fs, hs))

    # PROVER ----------------------------------------------------------------------------------------
    def prove_bytes(self, assertions: Sequence[dict], inputs=None, seed=None, _reuse_resident_trace: bool = False) -> bytes:
        """Stark.prove + serialize in one crossing: the serialized proof as produced on the device path."""
        if not isinstance(assertions, (list, tuple)):
            raise TypeError('Assertions parameter must be an array')
        if len(assertions) == 0:
            raise TypeError('At least one assertion must be provided')
        air = self.air
        p = air.modulus
        init = [int(v) % p for v in air.init(inputs or [], seed or [])]
        if len(init) != air.trace_register_count:
            raise StarkError('Failed to generate the execution trace: initial state has the wrong width')
        a_blob = b''.join(struct.pack('<II', int(a['register']), int(a['step'])) + (int(a['value']) % p).to_bytes(16, 'little')
                          for a in assertions)
        init_blob = b''.join(v.to_bytes(16, 'little') for v in init)
        in_blob = input_cbuf(air, inputs)          # zero-copy columns of the input registers
        shapes = air.input_shapes(inputs or [])
        s_blob = bytes([len(shapes)]) + b''.join(bytes([len(s)]) + b''.join(struct.pack('<I', x) for x in s) for s in shapes)
        out_p, out_n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._lib.gs_stark_prove_ex(self._handle, a_blob, len(assertions), init_blob, in_blob, s_blob, len(s_blob),
                                         1 if _reuse_resident_trace else 0, C.byref(out_p), C.byref(out_n))
        if rc == -4:
            raise StarkError(self._lib.gs_last_error(self.context.handle).decode())
        self.context.check(rc)
        return C.string_at(out_p, out_n.value)

    def prove(self, assertions: Sequence[dict], inputs=None, seed=None) -> dict:
        """lib/Stark.ts:81-163 -> StarkProof {evRoot, evProof, ldProof, iShapes}"""
        return self.parse(self.prove_bytes(assertions, inputs, seed))

    def generateExecutionTrace(self, inputs=None, seed=None) -> List[List[int]]:
        """lib/Stark.ts:252-257: the execution trace as register rows (host only)."""
        return generate_execution_trace(self.air, inputs, seed)

    def last_timing(self):
        """(CUDA-event ms of the device part, host wall-clock ms) of the last prove"""
        d, h = C.c_float(), C.c_double()
        self._lib.gs_stark_last_timing(self._handle, C.byref(d), C.byref(h))
        return d.value, h.value

    # VERIFIER ----------------------------------------------------------------------------------------
    def verify(self, assertions: Sequence[dict], proof, publicInputs=None) -> bool:
        """lib/Stark.ts:167-248 (native host code in the library; throws StarkError like the reference)."""
        buf = proof if isinstance(proof, (bytes, bytearray)) else self.serialize(proof)
        return verify_proof(self.air, dict(hashAlgorithm=self.hashAlgorithm, exeQueryCount=self.exeQueryCount,
                                           friQueryCount=self.friQueryCount), assertions, buf, publicInputs)

    def compose_backend(self) -> str:
        """'nvrtc <hash>' when the constraint evaluator was compiled for this AIR, else 'interpreter (<reason>)'"""
        return (self._lib.gs_stark_compose_backend(self._handle) or b'').decode()

    def stage_times(self) -> List:
        return json.loads(self._lib.gs_stark_stage_times(self._handle).decode() or '[]')

    # test hooks
    def _set_debug(self, on: bool = True):
        self.context.check(self._lib.gs_stark_set_debug(self._handle, 1 if on else 0))

    def _read_intermediate(self, which: int) -> List[int]:
        n = self.air.trace_length * self.air.extension_factor
        count = {0: self.air.trace_register_count * n, 1: n, 2: n, 3: self.air.trace_register_count * self.air.trace_length}[which]
        buf = C.create_string_buffer(count * 16)
        self.context.check(self._lib.gs_stark_read_intermediate(self._handle, which, buf, count * 16))
        raw = buf.raw
        return [int.from_bytes(raw[i:i + 16], 'little') for i in range(0, len(raw), 16)]

    # WIRE FORMAT (lib/Serializer.ts, lib/utils/serialization.ts, lib/utils/sizeof.ts) ------------------
    def _leaf_sizes(self):
        es = self.elementSize
        return (self.air.trace_register_count + self.air.secret_input_count) * es, es * 4

    def serialize(self, proof: dict) -> bytes:                      # Serializer.ts:35-79
        ev_leaf, ld_leaf = self._leaf_sizes()
        es = self.elementSize
        out = bytearray(proof['evRoot'])
        out += _write_merkle_proof(proof['evProof'], ev_leaf)
        ld = proof['ldProof']
        out += ld['lcRoot']
        out += _write_merkle_proof(ld['lcProof'], ld_leaf)
        out.append(len(ld['components']))
        for comp in ld['components']:
            out += comp['columnRoot']
            out += _write_merkle_proof(comp['columnProof'], ld_leaf)
            out += _write_merkle_proof(comp['polyProof'], ld_leaf)
        rl = len(ld['remainder'])
        out.append(0 if rl == 256 else rl)                           # zero means 256, :59-63
        for v in ld['remainder']:
            out += int(v).to_bytes(es, 'little')
        out.append(len(proof['iShapes']))
        for shape in proof['iShapes']:
            out.append(len(shape))
            for level in shape:
                out += struct.pack('<I', level)
        return bytes(out)

    def parse(self, buf: bytes) -> dict:                            # Serializer.ts:83-144
        ev_leaf, ld_leaf = self._leaf_sizes()
        return parse_proof(buf, ev_leaf, ld_leaf, self.elementSize, self.digestSize)

    def sizeOf(self, proof: dict) -> int:                           # sizeof.ts:12-53
        es, ds = self.elementSize, self.digestSize
        size = ds + _size_of_merkle_proof(proof['evProof'])
        ld = proof['ldProof']
        size += 1 + _size_of_merkle_proof(ld['lcProof']) + ds
        for comp in ld['components']:
            size += ds + _size_of_merkle_proof(comp['columnProof']) + _size_of_merkle_proof(comp['polyProof'])
        size += len(ld['remainder']) * es + 1
        size += 1 + sum(1 + 4 * len(s) for s in proof['iShapes'])
        return size


def parse_proof(buf: bytes, ev_leaf: int, ld_leaf: int, es: int, ds: int) -> dict:
    """Serializer.parseProof (lib/Serializer.ts:83-144): only the leaf sizes, the element size and the digest size enter the
    wire format -- not the trace length -- so a proof can be read before the instance for its input shapes exists"""
    ev_root = bytes(buf[:ds])
    ev_proof, off = _read_merkle_proof(buf, ds, ev_leaf, ds)
    lc_root = bytes(buf[off:off + ds]); off += ds
    lc_proof, off = _read_merkle_proof(buf, off, ld_leaf, ds)
    count = buf[off]; off += 1
    comps = []
    for _ in range(count):
        column_root = bytes(buf[off:off + ds]); off += ds
        column_proof, off = _read_merkle_proof(buf, off, ld_leaf, ds)
        poly_proof, off = _read_merkle_proof(buf, off, ld_leaf, ds)
        comps.append({'columnRoot': column_root, 'columnProof': column_proof, 'polyProof': poly_proof})
    rl = buf[off] or MAX_ARRAY_LENGTH; off += 1
    remainder = []
    for _ in range(rl):
        remainder.append(int.from_bytes(buf[off:off + es], 'little')); off += es
    n_inputs = buf[off]; off += 1
    shapes = []
    for _ in range(n_inputs):
        rank = buf[off]; off += 1
        shape = []
        for _ in range(rank):
            shape.append(struct.unpack_from('<I', buf, off)[0]); off += 4
        shapes.append(shape)
    return {'evRoot': ev_root, 'evProof': ev_proof,
            'ldProof': {'lcRoot': lc_root, 'lcProof': lc_proof, 'components': comps, 'remainder': remainder},
            'iShapes': shapes}


def _size_of_merkle_proof(p: BatchMerkleProof) -> int:              # sizeof.ts:55-99
    if len(p.values) == 0:
        raise ValueError('Array cannot be zero-length')
    if len(p.values) > MAX_ARRAY_LENGTH:
        raise ValueError(f'Array length ({len(p.values)}) cannot exceed {MAX_ARRAY_LENGTH}')
    if len(p.nodes) > MAX_ARRAY_LENGTH:
        raise ValueError(f'Matrix column count ({len(p.nodes)}) cannot exceed {MAX_ARRAY_LENGTH}')
    size = 1 + sum(len(v) for v in p.values) + 1 + len(p.nodes)
    for col in p.nodes:
        if len(col) >= MAX_MATRIX_COLUMN_LENGTH:
            raise ValueError(f'Matrix column length ({len(col)}) cannot exceed {MAX_MATRIX_COLUMN_LENGTH}')
        size += sum(len(x) for x in col)
    return size + 1


def _write_merkle_proof(p: BatchMerkleProof, leaf_size: int) -> bytes:   # serialization.ts:18-96
    out = bytearray([0 if len(p.values) == MAX_ARRAY_LENGTH else len(p.values)])
    for v in p.values:
        out += v
    out.append(0 if len(p.nodes) == MAX_ARRAY_LENGTH else len(p.nodes))
    for col in p.nodes:
        t = 1 if (len(col) > 0 and len(col[0]) == leaf_size) else 0
        out.append(((len(col) << 1) | t) & 0xFF)
    for col in p.nodes:
        for x in col:
            out += x
    out.append(p.depth)
    return bytes(out)


def _read_merkle_proof(buf: bytes, off: int, leaf_size: int, node_size: int):   # serialization.ts:25-124
    n = buf[off] or MAX_ARRAY_LENGTH; off += 1
    values = []
    for _ in range(n):
        values.append(bytes(buf[off:off + leaf_size])); off += leaf_size
    cols = buf[off] or MAX_ARRAY_LENGTH; off += 1
    lens, types = [], []
    for _ in range(cols):
        lt = buf[off]; off += 1
        lens.append(lt >> 1); types.append(lt & 1)
    nodes = []
    for i in range(cols):
        col = []
        for j in range(lens[i]):
            sz = (leaf_size if types[i] == 1 else node_size) if j == 0 else node_size
            col.append(bytes(buf[off:off + sz])); off += sz
        nodes.append(col)
    depth = buf[off]; off += 1
    return BatchMerkleProof(values, nodes, depth), off


class HostStark(Stark):
    """Stark over a prime field of at most 64 bits (BASELINE config 1: the Foo demo over 2^32 - 3*2^25 + 1, README.md:22-50):
    prove() and verify() run on the HOST inside libgenstark_b200.so (csrc/hoststark64.h) -- the counterpart of the reference's own
    fallback to unoptimised arithmetic for fields without a WASM backend (lib/Stark.ts:41-43).  Same surface as ``Stark``; needs no
    GPU.  The 128-bit field never comes here (it has the device path and no CPU one)."""

    def __init__(self, air: AirModule, options: Optional[dict] = None, logger=None, context=None):
        options = options or {}
        self.air = air.with_options(options.get('extensionFactor'))
        if self.air.modulus.bit_length() > 64:
            raise StarkError(f'field modulus {self.air.modulus} is not supported by the host path (at most 64 bits)')
        exe = options.get('exeQueryCount') or DEFAULT_EXE_QUERY_COUNT
        if exe < 1 or exe > MAX_EXE_QUERY_COUNT or int(exe) != exe:
            raise TypeError(f'Execution sample size must be an integer between 1 and {MAX_EXE_QUERY_COUNT}')
        fri = options.get('friQueryCount') or DEFAULT_FRI_QUERY_COUNT
        if fri < 1 or fri > MAX_FRI_QUERY_COUNT or int(fri) != fri:
            raise TypeError(f'FRI sample size must be an integer between 1 and {MAX_FRI_QUERY_COUNT}')
        alg = options.get('hashAlgorithm') or DEFAULT_HASH_ALGORITHM
        if alg not in HASH_ALGORITHMS:
            raise TypeError(f'Hash algorithm {alg} is not supported')
        if not self.air.extension_factor:
            raise TypeError('Extension factor is undefined')
        self.exeQueryCount, self.friQueryCount, self.hashAlgorithm = int(exe), int(fri), alg
        self.digestSize = 32
        self.elementSize = self.air.element_size
        self.logger = logger
        self._lib = _native.lib()
        self.context = None
        self._handle = None

    def close(self):
        pass

    def prove_bytes(self, assertions: Sequence[dict], inputs=None, seed=None, _reuse_resident_trace: bool = False) -> bytes:
        if not isinstance(assertions, (list, tuple)):
            raise TypeError('Assertions parameter must be an array')
        if len(assertions) == 0:
            raise TypeError('At least one assertion must be provided')
        air = self.air
        p = air.modulus
        init = [int(v) % p for v in air.init(inputs or [], seed or [])]
        if len(init) != air.trace_register_count:
            raise StarkError('Failed to generate the execution trace: initial state has the wrong width')
        a_blob = b''.join(struct.pack('<II', int(a['register']), int(a['step'])) + (int(a['value']) % p).to_bytes(16, 'little')
                          for a in assertions)
        in_blob = input_cbuf(air, inputs)          # zero-copy columns of the input registers
        shapes = air.input_shapes(inputs or [])
        s_blob = bytes([len(shapes)]) + b''.join(bytes([len(s)]) + b''.join(struct.pack('<I', x) for x in s) for s in shapes)
        blob = pack_air(air)
        out_p, out_n, err = C.POINTER(C.c_uint8)(), C.c_size_t(), C.create_string_buffer(512)
        rc = self._lib.gs_host_stark_prove(blob, len(blob), HASH_ALGORITHMS.index(self.hashAlgorithm), self.exeQueryCount, self.friQueryCount,
                                           a_blob, len(assertions), b''.join(v.to_bytes(16, 'little') for v in init), in_blob,
                                           s_blob, len(s_blob), C.byref(out_p), C.byref(out_n), err, 512)
        if rc == -4:
            raise StarkError(err.value.decode())
        if rc != 0:
            raise _native.NativeError(rc, err.value.decode())
        return C.string_at(out_p, out_n.value)

    def verify(self, assertions: Sequence[dict], proof, publicInputs=None) -> bool:
        if len(assertions) < 1:
            raise TypeError('At least one assertion must be provided')
        buf = proof if isinstance(proof, (bytes, bytearray)) else self.serialize(proof)
        air = self.air
        p = air.modulus
        a_blob = b''.join(struct.pack('<II', int(a['register']), int(a['step'])) + (int(a['value']) % p).to_bytes(16, 'little')
                          for a in assertions)
        # the proof carries the input shapes it was generated for (Serializer.ts:66-76): they have to be this instance's
        es = self.elementSize
        try:
            claimed = parse_proof(bytes(buf), (air.trace_register_count + air.secret_input_count) * es, 4 * es, es, 32)['iShapes']
        except (IndexError, struct.error):
            raise StarkError('Verification failed: malformed proof')
        try:
            expected = [list(x) for x in air.input_shapes(None)]
        except Exception:
            expected = None
        if expected is not None and [list(x) for x in claimed] != expected:
            raise StarkError(f'Verification failed: the proof was generated for input shapes {claimed}, this instance is built for {expected}')
        pub = air.expand_public_inputs(publicInputs or []) if air.expand_public_inputs else []
        pub_blob = b''.join((int(v) % p).to_bytes(16, 'little') for t in pub for v in t) if pub else None
        blob = pack_air(air)
        err = C.create_string_buffer(512)
        rc = self._lib.gs_host_stark_verify(blob, len(blob), HASH_ALGORITHMS.index(self.hashAlgorithm), self.exeQueryCount, self.friQueryCount,
                                            a_blob, len(assertions), bytes(buf), len(buf), pub_blob, err, 512)
        if rc != 0:
            raise StarkError(err.value.decode() or f'verification failed (status {rc})')
        return True

    def generateExecutionTrace(self, inputs=None, seed=None):
        raise StarkError('generateExecutionTrace is not provided on the host path for small fields')

    def compose_backend(self) -> str:
        return 'host (field of at most 64 bits)'

    def stage_times(self):
        return []

    def last_timing(self):
        return 0.0, 0.0


def verify_proof(air: AirModule, options: dict, assertions: Sequence[dict], proof_bytes: bytes, publicInputs=None) -> bool:
    """Stark.verify without a device: O(queries log N) host work inside libgenstark_b200.so (gs_stark_verify)."""
    if len(assertions) < 1:
        raise TypeError('At least one assertion must be provided')
    lib = _native.lib()
    air = air.with_options(options.get('extensionFactor'))
    p = air.modulus
    alg = options.get('hashAlgorithm') or DEFAULT_HASH_ALGORITHM
    if alg not in HASH_ALGORITHMS:
        raise TypeError(f'Hash algorithm {alg} is not supported')
    a_blob = b''.join(struct.pack('<II', int(a['register']), int(a['step'])) + (int(a['value']) % p).to_bytes(16, 'little')
                      for a in assertions)
    # the proof carries the input shapes it was generated for (Serializer.ts:66-76); the reference builds its verification
    # context from them (Stark.ts:177), here the instance already has a trace length, so they have to agree
    es = max(8, (p.bit_length() + 7) // 8)
    try:
        claimed = parse_proof(bytes(proof_bytes), (air.trace_register_count + air.secret_input_count) * es, 4 * es, es, 32)['iShapes']
    except (IndexError, struct.error):
        raise StarkError('Verification failed: malformed proof')
    try:
        expected = [list(x) for x in air.input_shapes(None)]
    except Exception:                       # an AIR whose shapes depend on the inputs: nothing to compare with
        expected = None
    if expected is not None and [list(x) for x in claimed] != expected:
        raise StarkError(f'Verification failed: the proof was generated for input shapes {claimed}, this instance is built for {expected}')
    pub_blob = public_blob(air, publicInputs)
    blob = pack_air(air)
    err = C.create_string_buffer(512)
    rc = lib.gs_stark_verify(blob, len(blob), HASH_ALGORITHMS.index(alg), int(options.get('exeQueryCount') or DEFAULT_EXE_QUERY_COUNT),
                             int(options.get('friQueryCount') or DEFAULT_FRI_QUERY_COUNT), a_blob, len(assertions),
                             bytes(proof_bytes), len(proof_bytes), pub_blob, err, 512)
    if rc != 0:
        raise StarkError(err.value.decode() or f'verification failed (status {rc})')
    return True


def generate_execution_trace(air: AirModule, inputs=None, seed=None) -> List[List[int]]:
    """context.generateExecutionTrace() (lib/Stark.ts:97,252-257) without a device: gs_air_generate_trace runs the
    AIR's transition function compiled to native code (or interpreted: gs_trace_backend says which)."""
    lib = _native.lib()
    p = air.modulus
    init = [int(v) % p for v in air.init(inputs or [], seed or [])]
    if len(init) != air.trace_register_count:
        raise StarkError('Failed to generate the execution trace: initial state has the wrong width')
    in_blob = input_cbuf(air, inputs)
    blob = pack_air(air)
    r, t = air.trace_register_count, air.trace_length
    out = C.create_string_buffer(16 * r * t)
    rc = lib.gs_air_generate_trace(blob, len(blob), b''.join(v.to_bytes(16, 'little') for v in init), in_blob, out)
    if rc != 0:
        raise StarkError(f'Failed to generate the execution trace (status {rc})')
    raw = out.raw
    return [[int.from_bytes(raw[16 * (k * t + i):16 * (k * t + i) + 16], 'little') for i in range(t)] for k in range(r)]


def trace_backend() -> str:
    """'jit <hash>' or 'interpreter (<reason>)' for the calling thread's last trace generation"""
    return (_native.lib().gs_trace_backend() or b'').decode()


class ScriptStark:
    """``Stark`` for an AirAssembly component whose trace length follows from the inputs (``for each`` nesting):
    the reference builds a proving context per call (lib/Stark.ts:90); here one device instance is kept per
    distinct input shape.  Same public surface as ``Stark``."""

    def __init__(self, component, options: Optional[dict] = None, logger=None, context: Optional[Context] = None):
        self.component, self.options, self.logger, self.context = component, dict(options or {}), logger, context
        self._by_shape: Dict[tuple, Stark] = {}
        self._proto: Optional[Stark] = None

    def _stark_for_shapes(self, shapes) -> Stark:
        key = tuple(tuple(s) for s in shapes)
        st = self._by_shape.get(key)
        if st is None:
            air = self.component.module(shapes, self.options.get('extensionFactor'))
            st = Stark(air, self.options, self.logger, context=self.context or (self._proto.context if self._proto else None))
            self._by_shape[key] = st
            self._proto = self._proto or st
        return st

    def _unit_module(self) -> AirModule:
        """the component with one value per input level: register counts, degrees and the extension factor do not depend
        on the input shapes, only the trace length does"""
        shapes = [[1] * r for r in self.component._rank]
        return self.component.module(shapes, self.options.get('extensionFactor'))

    @property
    def air(self):
        return self._proto.air if self._proto is not None else self._unit_module()

    @property
    def securityLevel(self) -> int:                                # Stark.ts:62-77 (no device instance needed)
        air = self.air
        e = air.extension_factor
        exe = self.options.get('exeQueryCount') or DEFAULT_EXE_QUERY_COUNT
        fri = self.options.get('friQueryCount') or DEFAULT_FRI_QUERY_COUNT
        return math.floor(min(_pow_log2(e / air.max_constraint_degree, int(exe)), math.log2(e) * int(fri), 32 * 4))

    def prove_bytes(self, assertions, inputs=None, seed=None) -> bytes:
        return self._stark_for_shapes(self.component.input_shapes(inputs or [])).prove_bytes(assertions, inputs, seed)

    def prove(self, assertions, inputs=None, seed=None) -> dict:
        st = self._stark_for_shapes(self.component.input_shapes(inputs or []))
        return st.parse(st.prove_bytes(assertions, inputs, seed))

    def generateExecutionTrace(self, inputs=None, seed=None):
        return generate_execution_trace(self.component.module_for(inputs or [], self.options.get('extensionFactor')), inputs, seed)

    def _shapes_of(self, proof) -> List[List[int]]:
        if isinstance(proof, (bytes, bytearray)):
            # iShapes close the wire format (Serializer.ts:66-76): walk back from the end is ambiguous, so parse with
            # the shape-independent reader (leaf sizes do not depend on the trace length)
            es = max(8, (self.component.modulus.bit_length() + 7) // 8)
            proof = parse_proof(proof, (self.component.trace_register_count + self.component.secret_input_count) * es, 4 * es, es, 32)
        return [list(s) for s in proof['iShapes']]

    def verify(self, assertions, proof, publicInputs=None) -> bool:
        return self._stark_for_shapes(self._shapes_of(proof)).verify(assertions, proof, publicInputs)

    def serialize(self, proof: dict) -> bytes:
        return self._stark_for_shapes(proof['iShapes']).serialize(proof)

    def parse(self, buf: bytes) -> dict:
        return self._stark_for_shapes(self._shapes_of(buf)).parse(buf)

    def sizeOf(self, proof: dict) -> int:
        return self._stark_for_shapes(proof['iShapes']).sizeOf(proof)


def instantiate(source, component: str = 'default', options: Optional[dict] = None, logger=None, context: Optional[Context] = None):
    """index.ts:18-33.  ``source``: an ``AirModule`` (already lowered), AirAssembly text (bytes / str) or the path of
    an ``.aa`` file; ``component``: the export to prove.  For backward compatibility ``instantiate(air, options)``
    with an ``AirModule`` accepts the options as second argument."""
    if isinstance(source, AirModule):
        if isinstance(component, dict) and options is None:
            component, options = 'default', component
        if source.modulus.bit_length() <= 64:
            return HostStark(source, options, logger)            # fields of at most 64 bits: host path (no WASM backend in the reference either)
        return Stark(source, options, logger, context=context)
    from . import assembly
    comp = assembly.compile(source).component(component)
    if comp.modulus.bit_length() <= 64 and comp.input_count == 0:
        return HostStark(comp.module([], (options or {}).get('extensionFactor')), options, logger)
    if comp.modulus.bit_length() > 128 or comp.modulus != (2**128 - 9 * 2**32 + 1):
        # the reference falls back to JS bigint arithmetic for such fields (README.md:118,224); here fields of at most 64 bits
        # have a host path (HostStark) and the device path is p128 only
        raise StarkError(f'field modulus {comp.modulus} is not supported by the B200 backend (p128 on the device, at most 64 bits on the host)')
    if comp.input_count == 0:
        return Stark(comp.module([], (options or {}).get('extensionFactor')), options, logger, context=context)
    return ScriptStark(comp, options, logger, context=context)
