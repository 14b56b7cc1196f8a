import random

from genstark_b200.air import P128


def rand_elems(n, seed):
    r = random.Random(seed)
    special = [0, 1, 2, P128 - 1, P128 - 2, 2**64, 2**96, 2**127, P128 - 2**32, 9 * 2**32 - 1, 9 * 2**32]
    out = [r.randrange(P128) for _ in range(n)]
    for i in range(min(n // 4, len(special))):
        out[(i * 7) % n] = special[i]
    return out


_ctx = None


def gpu_field():
    """one GpuField per test session"""
    global _ctx
    from genstark_b200.field import GpuField
    if _ctx is None:
        _ctx = GpuField()
    return _ctx
