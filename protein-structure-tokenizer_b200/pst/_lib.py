"""ctypes binding of libpst_b200.so (include/pst_abi.h).  There is NO fallback: if the
shared library is missing or a symbol is absent, importing callers fail loudly."""
from __future__ import annotations

import ctypes as C
import os

LIB_NAME = "libpst_b200.so"
# PST_LIB_PATH: developer override (instrumented builds for tools/); still the same library, still no fallback
LIB_PATH = os.environ.get("PST_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

PST_ABI_VERSION = 1
PST_ERR_WORKSPACE_TOO_SMALL = -4
PST_ERR_PDB_MODEL_COUNT, PST_ERR_PDB_INSERTION_CODE, PST_ERR_PDB_MALFORMED = -8, -9, -10
PST_ERR_FILE_NOT_FOUND = -11
PST_ERR_NON_FINITE = -12  # device status: a latent reached the quantiser as Inf / NaN
PST_MAX_LEVELS = 8


class PstConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("seq_max_size", C.c_int32),
        ("max_out_len", C.c_int32),
        ("num_neighbor", C.c_int32),
        ("downsampling_ratio", C.c_int32),
        ("num_levels", C.c_int32),
        ("levels", C.c_int32 * PST_MAX_LEVELS),
        ("gnn_layers", C.c_int32),
        ("num_blocks", C.c_int32),
        ("precision", C.c_int32),
        ("max_len", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/pst_abi.h
SIGNATURES = {
    "pst_abi_version": (C.c_int, []),
    "pst_status_string": (C.c_char_p, [C.c_int]),
    "pst_weight_blob_floats": (C.c_size_t, [C.POINTER(PstConfig)]),
    "pst_model_create": (C.c_int, [C.POINTER(PstConfig), C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "pst_model_destroy": (None, [C.c_void_p]),
    "pst_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "pst_featurize_knn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pst_encode_graph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pst_quantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pst_fsq_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pst_indexes_to_codes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pst_tokenize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pst_parse_pdb": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "pst_parse_pdb_chain": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int32)]),
    "pst_parse_mmcif": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_int32)]),
    "pst_parse_pdb_batch": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_parse_pdb_files": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_read_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_last_launch_count": (C.c_int, [C.c_void_p]),
    "pst_graph_cache_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pst_graph_cache_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "pst_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pst_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python protein-structure-tokenizer_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pst_abi_version() != PST_ABI_VERSION:
        raise RuntimeError("libpst_b200.so ABI version mismatch")
    _lib = lib
    return lib


class PstError(RuntimeError):
    def __init__(self, status: int, where: str):
        msg = load().pst_status_string(status).decode()
        super().__init__(f"{where}: {msg} (status {status})")
        self.status = status


def check(status: int, where: str) -> None:
    if status != 0:
        raise PstError(status, where)
