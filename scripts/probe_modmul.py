import sys, ctypes as C
sys.path.insert(0, '.')
from genstark_b200.field import GpuField
f = GpuField()
L = f._lib
ms = C.c_float()
for blocks in (148*4, 148*8):
    f.ctx.check(L.gs_debug_modmul_probe(f.ctx.handle, blocks, 2000, C.byref(ms)))
    n = blocks*256*4*2000
    print(f'modmul probe blocks={blocks}: {ms.value:.3f} ms -> {n/ms.value/1e6:.1f} G modmul/s')
