"""Small transforms through every K1 kernel shape plus one small prove -- meant to run under
`compute-sanitizer --tool racecheck` / `--tool memcheck` (shared-memory hazards in the tile exchanges, out-of-bounds)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genstark_b200.field import GpuField

f = GpuField()
L, ctx = f._lib, f.ctx
for log_t, log_n, inv, rows in ((16, 17, 0, 2), (17, 17, 1, 1), (18, 19, 0, 1), (19, 19, 0, 1), (20, 20, 0, 1), (12, 14, 0, 2), (14, 14, 1, 1)):
    src, dst, work = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_t, C.byref(src)))
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_n, C.byref(dst)))
    ctx.check(L.gs_mat_alloc(ctx.handle, rows, 1 << log_n, C.byref(work)))
    ctx.check(L.gs_mat_fill_random(ctx.handle, src, 0xB200))
    ctx.check(L.gs_ntt_into(ctx.handle, src, dst, work, inv))
    ctx.sync()
    for h in (src, dst, work):
        L.gs_mat_free(h)
    print('transform', log_t, log_n, inv, rows, 'done', flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'prove':
    from genstark_b200 import workloads
    from genstark_b200.stark import Stark
    air, opts, a, inputs, seed = workloads.mimc(1 << 10, 8)
    st = Stark(air, opts)
    print('proof bytes', len(st.prove_bytes(a, inputs, seed)))
