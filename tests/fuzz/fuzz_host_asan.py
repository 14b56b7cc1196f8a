"""Mutation fuzz of the host-side parsers against an ASan/UBSan build of host.cpp (LD_PRELOAD=libasan)."""
import sys, random, struct, ctypes as C, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['GS_TRACE_JIT'] = '0'
import cases
from genstark_b200.air import pack_air, input_blob
from genstark_b200.stark import HASH_ALGORITHMS
from oracle.stark import Stark as OracleStark
L = C.CDLL(os.environ.get('GS_FUZZ_LIB', '/tmp/genstark_fuzz/libhostasan.so'))
mode = sys.argv[1]; name = sys.argv[2]; seed = int(sys.argv[3]); N = int(sys.argv[4])
air, opts, a, inputs, sd = {'mimc': lambda: cases.mimc(64, 8), 'poseidon': lambda: cases.poseidon(2, 1, e=16), 'rescue': lambda: cases.rescue(2)}[name]()
blob = pack_air(air); p = air.modulus
r = random.Random(seed)
def mutate(src, lenbytes=False):
    t = bytearray(src)
    for _ in range(r.choice([1, 1, 2, 3])):
        k = r.choice(['flip', 'byte', 'word', 'trunc', 'insert', 'delete', 'splice', 'len'])
        if len(t) < 8: break
        if k == 'flip': q = r.randrange(len(t)); t[q] ^= 1 << r.randrange(8)
        elif k == 'byte': t[r.randrange(len(t))] = r.randrange(256)
        elif k == 'word':
            q = r.randrange(0, len(t) - 4) & ~3
            t[q:q + 4] = struct.pack('<I', r.choice([0, 1, 2, 3, 31, 32, 33, 63, 64, 65, 255, 256, 1023, 65535, 65536, 2**24, 2**31 - 1, 2**31, 2**32 - 1]))
        elif k == 'trunc': t = t[:r.randrange(len(t))]
        elif k == 'insert': q = r.randrange(len(t)); t[q:q] = bytes(r.randrange(256) for _ in range(r.randrange(1, 17)))
        elif k == 'splice':
            i = r.randrange(len(t) - 1); j = min(len(t), i + 1 + r.randrange(200)); kk = r.randrange(len(t) - (j - i) + 1); t[kk:kk + j - i] = t[i:j]
        elif k == 'len': t[r.randrange(len(t))] = r.choice([0, 1, 2, 15, 16, 17, 127, 128, 129, 254, 255])
        else: q = r.randrange(len(t)); del t[q:q + r.randrange(1, 17)]
    return bytes(t)
if mode == 'air':
    init = b''.join((int(v) % p).to_bytes(16, 'little') for v in air.init(inputs or [], sd or []))
    ib = input_blob(air, inputs)
    out = C.create_string_buffer(16 * air.trace_register_count * air.trace_length * 4 + 4096)
    L.asan_generate_trace.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_void_p]
    assert L.asan_generate_trace(blob, len(blob), init, ib, out) == 0
    for it in range(N):
        t = mutate(blob)
        if len(t) >= 40:
            lt, le = struct.unpack_from('<2I', t, 28); rr = struct.unpack_from('<I', t, 20)[0]
            if lt > 8 or le > 5 or rr > 4 * air.trace_register_count: continue
        # exact-size heap copies so ASan sees any read past the end
        tb = C.create_string_buffer(t, len(t))
        L.asan_generate_trace(tb, len(t), init, ib, out)
else:
    ora = OracleStark(air, opts)
    proof = ora.serialize(ora.prove(a, inputs, sd))
    a_blob = b''.join(struct.pack('<II', int(x['register']), int(x['step'])) + (int(x['value']) % p).to_bytes(16, 'little') for x in a)
    pub = air.expand_public_inputs(inputs[4:]) if (name == 'poseidon' and air.expand_public_inputs) else []
    pub_blob = b''.join((int(v) % p).to_bytes(16, 'little') for tt in pub for v in tt) if pub else None
    alg = HASH_ALGORITHMS.index(opts['hashAlgorithm'])
    err = C.create_string_buffer(512)
    L.gs_stark_verify.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
    def verify(b, pr):
        bb = C.create_string_buffer(b, len(b)); pp = C.create_string_buffer(pr, len(pr))
        return L.gs_stark_verify(bb, len(b), alg, int(opts['exeQueryCount']), int(opts['friQueryCount']), a_blob, len(a), pp, len(pr), pub_blob, err, 512)
    assert verify(blob, proof) == 0, err.value
    for it in range(N):
        if mode == 'proof': verify(blob, mutate(proof))
        else:
            t = mutate(blob)
            if len(t) >= 40:
                lt, le = struct.unpack_from('<2I', t, 28)
                if lt > 12 or le > 5: continue
            verify(t, proof)
print(mode, name, seed, 'done', N, flush=True)
