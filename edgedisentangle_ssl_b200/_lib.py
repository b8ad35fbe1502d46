"""ctypes binding of libedis.so (the C ABI in include/edis.h).

There is no CPU fallback: if the shared library is missing this module raises, and every op
raises `EdisError` with the library's message on a non-zero status.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# EDIS_LIB: alternative build of the same library (tuning variants); default is the in-tree one
LIB_PATH = os.environ.get("EDIS_LIB") or os.path.join(_PKG, "libedis.so")


class EdisError(RuntimeError):
    pass


class LayerDesc(Structure):
    """edis_layer_desc (include/edis.h)."""
    _fields_ = [("att", c_int32), ("C", c_int32), ("D", c_int32), ("Dv", c_int32),
                ("training", c_int32), ("p", c_float), ("seed", c_uint64), ("flags", c_int32),
                ("reserved", c_int32)]


ERR_STALE = -5
FLAG_PLAIN_MEAN = 1
FLAG_NO_GX = 2
FLAG_PHASE_DST, FLAG_PHASE_SRC, FLAG_PHASE_GX = 4, 8, 16


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libedis.so not found at %s: build it with `python -m edgedisentangle_ssl_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_i64p = POINTER(c_int64)
_i32p = POINTER(c_int32)
_f32p = POINTER(c_float)
_f64p = POINTER(c_double)
_descp = POINTER(LayerDesc)
_P = c_void_p  # device pointers travel as integers

# (g, d, P, ldp, Q, ldq, a, V, ldv, bias, hpre, edge_e, stats, esign, g_out, g_edge_e,
#  gP, ldgp, gQ, ldgq, ga, gV, ldgv, edge_rec, gh, workspace, workspace_bytes, stream)
_BWD_ARGS = [c_void_p, _descp, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P,
             _P, c_int64, _P, c_int64, _P, _P, c_int64, _P, _P, _P, c_int64, _P]

SIGNATURES = {
    "edis_last_error": (c_char_p, []),
    "edis_version": (c_char_p, []),
    "edis_build_adjacency_host": (c_int64, [c_int64, c_int64, _i64p, _i64p, _f64p, _i64p, _i64p, _f32p]),
    "edis_graph_create": (c_int, [c_int64, c_int64, _i64p, _i64p, c_int, c_int, POINTER(c_void_p)]),
    "edis_graph_create_rect": (c_int, [c_int64, c_int64, c_int64, _i64p, _i64p, c_int, c_int, POINTER(c_void_p)]),
    "edis_graph_destroy": (None, [c_void_p]),
    "edis_graph_info": (c_int, [c_void_p, _i64p]),
    "edis_graph_input_entries": (c_int64, [c_void_p]),
    "edis_graph_export": (c_int, [c_void_p, _i64p, _i32p, _i64p, _i64p, _i32p, _i32p]),
    "edis_graph_workspace_bytes": (c_int64, [c_void_p, c_int64]),
    "edis_edge_list_key": (c_uint64, [c_int64, c_int64, c_int64, _i64p, _i64p, c_int]),
    "edis_graph_save": (c_int, [c_void_p, c_char_p, c_uint64]),
    "edis_graph_load": (c_int, [c_char_p, c_uint64, c_int, c_int, c_int, POINTER(c_void_p)]),
    "edis_disga_fwd": (c_int, [c_void_p, _descp, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P,
                               _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "edis_disga_rec_bytes": (c_int64, [c_void_p, _descp]),
    "edis_disga_sign_bytes": (c_int64, [c_void_p, _descp]),
    "edis_disga_bwd": (c_int, _BWD_ARGS),
    "edis_disga_bwd_dst": (c_int, _BWD_ARGS),
    "edis_disga_bwd_src": (c_int, _BWD_ARGS),
    "edis_disga_sage_fwd": (c_int, [c_void_p, _descp, _P, c_int64, _P, c_int64, _P, _P, c_int64,
                                    _P, _P, _P, _P, _P, c_int64, _P]),
    "edis_disga_sage_bwd": (c_int, [c_void_p, _descp, _P, c_int64, _P, c_int64, _P, _P, c_int64,
                                    _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P,
                                    c_int64, _P]),
    "edis_disga_sage_fused_gx": (c_int, [_descp]),
    "edis_pair_score_fwd": (c_int, [_descp, c_int64, c_int64, _P, _P, c_int32, c_int32, _P, c_int64,
                                    _P, c_int64, _P, _P, _P, _P]),
    "edis_pair_score_bwd": (c_int, [_descp, c_int64, c_int64, _P, _P, c_int32, c_int32, _P, c_int64,
                                    _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "edis_pair_sign_bytes": (c_int64, [_descp, c_int64, c_int32, c_int32]),
    "edis_ssl_wmse_fwd": (c_int, [c_int64, c_int32, _P, _P, c_int64, c_int64, _P, _P, c_int64, _P]),
    "edis_ssl_wmse_bwd": (c_int, [c_int64, c_int32, _P, _P, c_int64, c_int64, _P, _P, _P]),
    "edis_nll_const_label_fwd": (c_int, [c_int64, c_int32, _P, c_int32, _P, _P, c_int64, _P]),
    "edis_nll_const_label_bwd": (c_int, [c_int64, c_int32, _P, c_int32, _P, _P, _P]),
    "edis_sp_softmax_fwd": (c_int, [c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "edis_sp_softmax_bwd": (c_int, [c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "edis_sp_matmul_fwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "edis_sp_matmul_bwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "edis_merge_pairs_host": (c_int64, [c_int64, _i64p, c_int64, _i64p, c_int64, _i64p, _i64p, _f32p]),
    "edis_rand_hits_host": (c_int64, [c_void_p, c_int64, c_int64, ctypes.c_uint32, _i64p, c_int64]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.edis_last_error().decode("utf-8", "replace")


def check(status, what):
    if status is not None and status < 0:
        raise EdisError("%s failed (%d): %s" % (what, status, last_error()))
    return status


def np_ptr(arr, ctype):
    """Host numpy array -> typed ctypes pointer (array must stay alive during the call)."""
    return arr.ctypes.data_as(POINTER(ctype))


def version():
    return lib.edis_version().decode()
