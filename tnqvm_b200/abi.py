"""ctypes binding of include/mps_b200.h (libmps_b200.so).  No torch types cross this boundary."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MpsError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "lib", "libmps_b200.so")


SYMBOLS = {
    "mps_create": ([C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)], C.c_int),
    "mps_create_sharded": ([C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)], C.c_int),
    "mps_shard_partition": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "mps_shard_plan_debug": ([C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)], C.c_int),
    "mps_shard_layout": ([C.c_void_p, C.POINTER(C.c_int), C.c_void_p], C.c_int),
    "mps_destroy": ([C.c_void_p], C.c_int),
    "mps_last_error": ([C.c_void_p], C.c_char_p),
    "mps_reset": ([C.c_void_p], C.c_int),
    "mps_set_option": ([C.c_void_p, C.c_char_p, C.c_double], C.c_int),
    "mps_apply_1q": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "mps_apply_2q": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "mps_apply_layer": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "mps_apply_gates": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "mps_flush": ([C.c_void_p], C.c_int),
    "mps_sync": ([C.c_void_p], C.c_int),
    "mps_norm": ([C.c_void_p, C.c_int, C.POINTER(C.c_double)], C.c_int),
    "mps_expval_z": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)], C.c_int),
    "mps_expval_z_all": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "mps_expval_zz_pairs": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "mps_amplitude": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)], C.c_int),
    "mps_statevector": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "mps_snapshot": ([C.c_void_p], C.c_int),
    "mps_restore": ([C.c_void_p], C.c_int),
    "mps_measure": ([C.c_void_p, C.c_int], C.c_int),
    "mps_clear_measure": ([C.c_void_p], C.c_int),
    "mps_seed": ([C.c_void_p, C.c_uint64], C.c_int),
    "mps_n_measured": ([C.c_void_p, C.POINTER(C.c_int)], C.c_int),
    "mps_sample": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)], C.c_int),
    "mps_bond_dims": ([C.c_void_p, C.c_void_p], C.c_int),
    "mps_singular_values": ([C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)], C.c_int),
    "mps_discarded_weight": ([C.c_void_p, C.POINTER(C.c_double)], C.c_int),
    "mps_fidelity_estimate": ([C.c_void_p, C.POINTER(C.c_double)], C.c_int),
    "mps_get_site": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "mps_set_site": ([C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int], C.c_int),
    "mps_set_sites": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "mps_get_sites": ([C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "mps_site_device_ptr": ([C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p], C.c_int),
    "mps_resize_site": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)], C.c_int),
    "mps_stats": ([C.c_void_p, C.c_void_p, C.c_int], C.c_int),
    "mps_get_stream": ([C.c_void_p, C.POINTER(C.c_void_p)], C.c_int),
}


def load_library():
    """Load libmps_b200.so.  Raises (never falls back) when the CUDA library has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise MpsError("libmps_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
        L = C.CDLL(p)
        for name, (argt, rest) in SYMBOLS.items():
            f = getattr(L, name)
            f.argtypes = argt
            f.restype = rest
        _LIB = L
    return _LIB
