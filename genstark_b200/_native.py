"""ctypes binding of libgenstark_b200.so (include/genstark_b200.h).  There is no CPU fallback: if the
library is missing it is built with nvcc, and if no CUDA device is present every compute call fails
loudly (NativeError)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'libgenstark_b200: {message} (status {code})')
        self.code = code
        self.message = message


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()
    # development: A/B another build of the same library (scripts/ build variants next to it); never a different implementation
    alt = os.environ.get('GS_NATIVE_LIB')
    if alt:
        path = alt if os.path.isabs(alt) else os.path.join(os.path.dirname(_build.LIB), alt)
    L = C.CDLL(path)
    vp, i32, i64, u64, cp = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_char_p
    P = C.POINTER
    sigs = {
        'gs_ctx_create': (i32, [i32, P(vp)]),
        'gs_ctx_destroy': (None, [vp]),
        'gs_last_error': (cp, [vp]),
        'gs_ctx_sync': (i32, [vp]),
        'gs_comm_unique_id': (i32, [C.c_char_p]),
        'gs_ctx_comm_init': (i32, [vp, i32, i32, cp]),
        'gs_shard_map': (i32, [i32, i32, i32, i64, i32, P(i64), P(i32)]),
        'gs_ctx_launch_count': (u64, [vp]),
        'gs_field_supported': (i32, [cp, C.c_size_t]),
        'gs_field_root_of_unity': (i32, [i32, C.c_char_p]),
        'gs_field_scalar_op': (i32, [i32, cp, cp, C.c_char_p]),
        'gs_mat_alloc': (i32, [vp, i64, i64, P(vp)]),
        'gs_mat_from_bytes': (i32, [vp, cp, i64, i64, P(vp)]),
        'gs_mat_to_bytes': (i32, [vp, vp, vp]),
        'gs_mat_shape': (i32, [vp, P(i64), P(i64)]),
        'gs_mat_device_ptr': (vp, [vp]),
        'gs_mat_free': (None, [vp]),
        'gs_interpolate_roots': (i32, [vp, vp, P(vp)]),
        'gs_eval_polys_at_roots': (i32, [vp, vp, i32, P(vp)]),
        'gs_vec_binary': (i32, [vp, i32, vp, vp, cp, P(vp)]),
        'gs_vec_div': (i32, [vp, vp, vp, P(vp)]),
        'gs_vec_combine_many': (i32, [vp, P(vp), i32, cp, P(vp)]),
        'gs_power_series': (i32, [vp, cp, i64, P(vp)]),
        'gs_pluck_vector': (i32, [vp, vp, i64, i64, P(vp)]),
        'gs_transpose_vector': (i32, [vp, vp, i32, i64, P(vp)]),
        'gs_fri_fold': (i32, [vp, vp, i32, i32, cp, P(vp)]),
        'gs_hash_merge_vector_rows': (i32, [vp, i32, P(vp), i32, P(vp)]),
        'gs_hash_digest_values': (i32, [vp, i32, vp, P(vp)]),
        'gs_digests_to_bytes': (i32, [vp, vp, vp]),
        'gs_digests_count': (i64, [vp]),
        'gs_digests_free': (None, [vp]),
        'gs_merkle_create': (i32, [vp, i32, vp, P(vp)]),
        'gs_merkle_root': (i32, [vp, vp, C.c_char_p]),
        'gs_merkle_prove_batch': (i32, [vp, vp, P(C.c_uint32), i32, vp, C.c_size_t, P(C.c_size_t)]),
        'gs_tree_free': (None, [vp]),
        'gs_stark_create': (i32, [vp, cp, C.c_size_t, i32, i32, i32, P(vp)]),
        'gs_stark_destroy': (None, [vp]),
        'gs_air_generate_trace': (i32, [cp, C.c_size_t, cp, cp, vp]),
        'gs_trace_backend': (C.c_char_p, []),
        'gs_stark_prove': (i32, [vp, cp, i32, cp, cp, cp, C.c_size_t, P(P(C.c_uint8)), P(C.c_size_t)]),
        'gs_stark_prove_ex': (i32, [vp, cp, i32, cp, cp, cp, C.c_size_t, i32, P(P(C.c_uint8)), P(C.c_size_t)]),
        'gs_stark_last_timing': (i32, [vp, P(C.c_float), P(C.c_double)]),
        'gs_ctx_profile': (i32, [vp, i32]),
        'gs_ctx_profile_report': (cp, [vp]),
        'gs_timer_begin': (i32, [vp]),
        'gs_timer_end': (i32, [vp, P(C.c_float)]),
        'gs_ntt_into': (i32, [vp, vp, vp, vp, i32]),
        'gs_stark_last_error': (cp, [vp]),
        'gs_host_stark_prove': (i32, [cp, C.c_size_t, i32, i32, i32, cp, i32, cp, cp, cp, C.c_size_t, P(P(C.c_uint8)), P(C.c_size_t), C.c_char_p, C.c_size_t]),
        'gs_host_stark_verify': (i32, [cp, C.c_size_t, i32, i32, i32, cp, i32, cp, C.c_size_t, cp, C.c_char_p, C.c_size_t]),
        'gs_vec_exp': (i32, [vp, vp, cp, P(vp)]),
        'gs_mat_mul_vector': (i32, [vp, vp, vp, P(vp)]),
        'gs_lde_cosets_into': (i32, [vp, vp, vp, vp, i32, i32]),
        'gs_mat_fill_random': (i32, [vp, vp, C.c_uint64]),
        'gs_stark_verify': (i32, [cp, C.c_size_t, i32, i32, i32, cp, i32, cp, C.c_size_t, cp, C.c_char_p, C.c_size_t]),
        'gs_stark_stage_times': (cp, [vp]),
        'gs_stark_compose_backend': (C.c_char_p, [vp]),
        'gs_stark_set_debug': (i32, [vp, i32]),
        'gs_stark_read_intermediate': (i32, [vp, i32, vp, C.c_size_t]),
        'gs_debug_modmul_probe': (i32, [vp, i32, i32, P(C.c_float)]),
        'gs_debug_butterfly_probe': (i32, [vp, i32, i32, P(C.c_float)]),
        'gs_debug_sqr_probe': (i32, [vp, i32, i32, P(C.c_float), P(C.c_uint32)]),
        'gs_debug_commit_columns': (i32, [vp, i32, P(vp), i32, P(vp)]),
        'gs_debug_tree_nodes': (i32, [vp, vp, vp, C.c_size_t]),
        'gs_field_prng': (i32, [cp, C.c_size_t, i32, C.c_char_p]),
        'gs_poly_interpolate': (i32, [cp, cp, i32, C.c_char_p]),
        'gs_poly_eval_at': (i32, [cp, i32, cp, C.c_char_p]),
        'gs_poly_mul': (i32, [cp, i32, cp, i32, C.c_char_p]),
        'gs_vec_combine': (i32, [vp, vp, vp, C.c_char_p]),
        'gs_quartic_interpolate_batch': (i32, [vp, vp, vp, P(vp)]),
        'gs_quartic_eval_batch': (i32, [vp, vp, vp, cp, P(vp)]),
        'gs_mat_stack': (i32, [vp, P(vp), i32, P(vp)]),
        'gs_mat_rows': (i32, [vp, vp, i64, i64, P(vp)]),
        'gs_mat_transpose': (i32, [vp, vp, P(vp)]),
        'gs_mat_reshape': (i32, [vp, i64, i64]),
        'gs_mat_get': (i32, [vp, vp, i64, i64, C.c_char_p]),
        'gs_hash_digest': (i32, [i32, cp, C.c_size_t, C.c_char_p]),
        'gs_merkle_verify_batch': (i32, [i32, cp, P(C.c_uint32), i32, cp, C.c_size_t]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L._gs_declared = list(sigs)
    _lib = L
    return L


def declared_symbols():
    lib()
    return list(_lib._gs_declared)


def check(ctx, rc: int):
    if rc != 0:
        msg = lib().gs_last_error(ctx)
        raise NativeError(rc, msg.decode() if msg else 'unknown error')
