"""ctypes binding of libqsv.so — the C-ABI declared in include/qsv.h.

This is the only place Python touches native code.  There is deliberately no fallback: if the
shared library is missing or a call fails, a RuntimeError is raised (the reference raises
RuntimeError for every failure on this path too, main/circuit.py:224-288).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqsv.so")

QSV_OP_DENSE, QSV_OP_MCX, QSV_OP_PHASE, QSV_OP_DIAG, QSV_OP_SWAP = 1, 2, 3, 4, 5
QSV_OPF_REAL_MATRIX = 1
QSV_MAX_TILE_K = 5
QSV_MAX_DIAG_K = 8
QSV_MAX_TILE_T = 13
QSV_OP_HEADER_BYTES = 96


class QsvPass(C.Structure):
    _fields_ = [
        ("n_local", C.c_int32), ("tile_bits", C.c_int32), ("low_bits", C.c_int32), ("n_hi", C.c_int32),
        ("hi_pos", C.c_uint8 * 16), ("rank_bits", C.c_uint64), ("max_dense_k", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/qsv.h declares: name -> (restype, argtypes)
_P, _I, _U64, _D = C.c_void_p, C.c_int, C.c_uint64, C.c_double
_IP = C.POINTER(C.c_int)
SYMBOLS = {
    "qsv_abi_version": (C.c_int, []),
    "qsv_last_error": (C.c_char_p, []),
    "qsv_launch_count": (C.c_uint64, []),
    "qsv_reset_launch_count": (None, []),
    "qsv_init_basis": (C.c_int, [_P, _I, _U64, _U64, _P]),
    "qsv_init_basis_sparse": (C.c_int, [_P, _I, _U64, _U64, _P]),
    "qsv_init_columns": (C.c_int, [_P, _I, _U64, _P]),
    "qsv_run_pass": (C.c_int, [_P, C.POINTER(QsvPass), _P, _P, _I, _P]),
    "qsv_run_pass_sparse": (C.c_int, [_P, C.POINTER(QsvPass), _P, _P, _I, _U64, _U64, _P]),
    "qsv_run_pass_rounds": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, _P]),
    "qsv_run_pass_rounds_sparse": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, _U64, _U64, _P]),
    "qsv_run_pass_rounds_fresh": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, _U64, _U64, C.c_uint32,
                                           C.c_uint32, _P]),
    "qsv_run_pass_rounds_sums": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, _U64, _U64, C.c_uint32,
                                          C.c_uint32, _I, _IP, _P, _P]),
    "qsv_jit_source_sums": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_compile_sums": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_source": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_compile": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_wanted": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32]),
    "qsv_jit_bank_score": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "qsv_jit_source_swap": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _I, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_compile_swap": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _I, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_run_pass_rounds_swap": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_uint64), _I, _IP, _I, _I, _U64, _P]),
    "qsv_jit_source_swap_ex": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _IP, _I, _P, C.c_size_t,
                                        C.POINTER(C.c_size_t)]),
    "qsv_jit_compile_swap_ex": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32, _I, _IP, _IP, _I, _P, C.c_size_t,
                                         C.POINTER(C.c_size_t)]),
    "qsv_run_pass_rounds_swap_ex": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, C.POINTER(C.c_uint64),
                                             C.POINTER(C.c_uint64), _I, _I, _IP, _IP, _I, _I, _U64, _U64, _U64, _P]),
    "qsv_run_pass_rounds_swap_fresh": (C.c_int, [_P, C.POINTER(QsvPass), _P, C.c_uint32, _P, C.POINTER(C.c_uint64),
                                                C.POINTER(C.c_uint64), _I, _I, _IP, _IP, _I, _I, _U64, _U64, _U64,
                                                C.c_uint32, C.c_uint32, _P]),
    "qsv_jit_disk_hits": (C.c_uint64, []),
    "qsv_jit_stats": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(_D)]),
    "qsv_jit_sums_static": (C.c_int, [C.POINTER(QsvPass), _P, C.c_uint32]),
    "qsv_path_sum": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P]),
    "qsv_plan_absorbed": (C.c_int, [_P, _P, _P, _I, _U64]),
    "qsv_plan_improve_tile": (C.c_int, [_P, _P, _P, _I, _I, _U64, _IP, _I, _I, _I, _I, _I, _U64, _IP, _IP, _IP]),
    "qsv_apply_dense": (C.c_int, [_P, _I, _I, _IP, _I, _IP, C.POINTER(_D), _U64, _P]),
    "qsv_apply_diag": (C.c_int, [_P, _I, _I, _IP, C.POINTER(_D), _U64, _P]),
    "qsv_apply_phase": (C.c_int, [_P, _I, _I, _IP, _D, _D, _U64, _P]),
    "qsv_apply_cx": (C.c_int, [_P, _I, _I, _IP, _I, _U64, _P]),
    "qsv_apply_dense_wide": (C.c_int, [_P, _P, _I, _I, _IP, _P, _P]),
    "qsv_apply_perm_wide": (C.c_int, [_P, _P, _I, _I, _IP, _P, _P]),
    "qsv_apply_diag_wide": (C.c_int, [_P, _I, _I, _IP, _P, _P]),
    "qsv_grover_oracle": (C.c_int, [_P, _U64, _U64, _P]),
    "qsv_diffusion_scratch_bytes": (C.c_size_t, [_I, _I]),
    "qsv_diffusion_sums": (C.c_int, [_P, _I, _I, _P, _P, _P]),
    "qsv_diffusion_update": (C.c_int, [_P, _I, _I, _P, _I, _P]),
    "qsv_modexp_involution": (C.c_int, [_P, _I, _I, _IP, _I, _IP, _U64, _U64, _U64, _P]),
    "qsv_probs": (C.c_int, [_P, _I, _IP, _P, _P]),
    "qsv_block_sums": (C.c_int, [_P, _I, _IP, _I, _P, _P]),
    "qsv_sample_scan": (C.c_int, [_P, _I, _IP, _U64, _U64, _D, _D, _P, _P]),
    "qsv_sample_begin": (C.c_int, [_P, _P]),
    "qsv_sample_level": (C.c_int, [_P, _I, _U64, _U64, _I, _IP, _P, _P, _P]),
    "qsv_sample_leaf": (C.c_int, [_P, _I, _U64, _U64, _I, _IP, _P, _P, _P]),
    "qsv_sample_select": (C.c_int, [_P, _I, _IP, _IP, _I, _U64, _I, _P, _P]),
    "qsv_sample_sequential": (C.c_int, [_P, _I, _IP, _D, _P, _P]),
    "qsv_swap_peer": (C.c_int, [_P, C.POINTER(C.c_uint64), _I, _I, _I, _IP, _P]),
    "qsv_swap_peer_ex": (C.c_int, [_P, C.POINTER(C.c_uint64), _I, _I, _I, _I, _IP, _IP, _U64, _U64, _P]),
    "qsv_pack_slices": (C.c_int, [_P, _P, _U64, _U64, _U64, _U64, _P]),
    "qsv_unpack_slices": (C.c_int, [_P, _P, _U64, _U64, _U64, _U64, _P]),
}

_lib = None


def load():
    """Load libqsv.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libqsv.so is missing (%s): build it with `python quantum-circuits_b200/build.py`. "
            "The state-vector engine has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.qsv_abi_version() != 1:
        raise RuntimeError("libqsv.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().qsv_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed: %s" % (what, msg))


def int_array(values):
    values = list(values)
    return (C.c_int * max(1, len(values)))(*values)
