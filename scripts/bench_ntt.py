"""K1 alone: NTT / iNTT / LDE timings (CUDA events on the context stream), median of 7 after 3 warm-ups."""
import ctypes as C
import os
import random
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genstark_b200.field import GpuField

f = GpuField()
L, ctx = f._lib, f.ctx
r = random.Random(0xB200)
cases = [('fwd 2^23', 23, 23, 0, 1), ('fwd 2^20', 20, 20, 0, 1), ('lde 2^20->2^23', 20, 23, 0, 1), ('inv 2^20', 20, 20, 1, 1),
         ('fwd 2^16 x12', 16, 16, 0, 12), ('lde 2^16->2^21 x12', 16, 21, 0, 12), ('fwd 2^24', 24, 24, 0, 1)]
for name, log_t, log_n, inv, rows in cases:
    t_, n_ = 1 << log_t, 1 << log_n
    chunk = r.randbytes(1 << 20)
    ba = bytearray((chunk * ((16 * t_ * rows >> 20) + 1))[:16 * t_ * rows])
    ba[15::16] = bytes(x & 0x7F for x in ba[15::16])
    src = f._from_bytes(bytes(ba), rows, t_)
    dst, work = C.c_void_p(), C.c_void_p()
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, n_, C.byref(dst)))
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, n_, C.byref(work)))
    ms = C.c_float()
    ts = []
    for i in range(10):
        L.gs_timer_begin(ctx.handle)
        ctx.check(L.gs_ntt_into(ctx.handle, src.handle, dst, work, inv))
        L.gs_timer_end(ctx.handle, C.byref(ms))
        if i >= 3:
            ts.append(ms.value)
    med = sorted(ts)[len(ts) // 2]
    print(f'{name:22s} {med:8.4f} ms  {rows * n_ / med / 1e6:9.2f} G elem/s  alg {16 * rows * (t_ + n_) / med / 1e6:8.1f} GB/s', flush=True)
    L.gs_mat_free(dst); L.gs_mat_free(work); src.free()
