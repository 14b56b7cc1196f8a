"""A few LDEs / NTTs through K1 alone (for ncu captures of the NTT kernels).  Usage: python scripts/lde_once.py [reps]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genstark_b200.field import GpuField

f = GpuField()
L, ctx = f._lib, f.ctx
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for log_t, log_n, inv, rows in ((20, 23, 0, 1), (20, 20, 1, 1), (23, 23, 0, 1), (16, 21, 0, 12)):
    src, dst, work = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_t, C.byref(src)))
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_n, C.byref(dst)))
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_n, C.byref(work)))
    ctx.check(L.gs_mat_fill_random(ctx.handle, src, 0xB200))
    for i in range(reps):
        ctx.check(L.gs_ntt_into(ctx.handle, src, dst, work, inv))
    ctx.sync()
    for h in (src, dst, work):
        L.gs_mat_free(h)
print('done')
