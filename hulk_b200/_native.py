"""ctypes binding of include/hulk_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU or pure-Python fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhulk_b200.so")

# names declared in include/hulk_b200.h (tests check the library exports every one of them)
EXPORTS = [
    "hulk_b200_version", "hulk_b200_strerror", "hulk_b200_last_error", "hulk_b200_create",
    "hulk_b200_destroy", "hulk_b200_set_cws_tables", "hulk_b200_generate_cws_tables",
    "hulk_b200_generate_cws_tables_async",
    "hulk_b200_set_cws_tables_device", "hulk_b200_reset", "hulk_b200_profile_enable", "hulk_b200_profile_read", "hulk_b200_profile_timeline", "hulk_b200_set_overlap",
    "hulk_b200_new_cws", "hulk_b200_new_cws_parallel", "hulk_b200_push_reads", "hulk_b200_push_reads_fixed",
    "hulk_b200_push_reads_device", "hulk_b200_sync_inputs", "hulk_b200_flush", "hulk_b200_sync",
    "hulk_b200_finish", "hulk_b200_snapshot_async", "hulk_b200_get_stats", "hulk_b200_histogram_device_ptr", "hulk_b200_stream",
    "hulk_b200_merge_histogram", "hulk_b200_add_minimizer_count", "hulk_b200_get_histogram", "hulk_b200_get_estimates",
    "hulk_b200_get_cms", "hulk_b200_minimizers", "hulk_b200_jump_hash", "hulk_b200_jump_hash_fx", "hulk_b200_rcp_selftest", "hulk_b200_get_folded_table",
    "hulk_b200_md5_mins", "hulk_b200_sketch_json", "hulk_b200_write_json", "hulk_b200_sketch_json_minhash",
    "hulk_b200_write_json_minhash", "hulk_b200_alloc_pinned",
    "hulk_b200_free_pinned", "hulk_b200_reader_open", "hulk_b200_reader_next", "hulk_b200_reader_error",
    "hulk_b200_reader_close", "hulk_b200_sketch_reader", "hulk_b200_sketch_load", "hulk_b200_sketch_find",
    "hulk_b200_sketch_banner", "hulk_b200_sketch_free", "hulk_b200_smash",
    "hulk_b200_peer_export", "hulk_b200_peer_flags", "hulk_b200_peer_connect", "hulk_b200_peer_connect_local",
    "hulk_b200_group_create", "hulk_b200_group_destroy", "hulk_b200_group_last_error", "hulk_b200_group_size",
    "hulk_b200_group_member", "hulk_b200_group_set_cws_tables", "hulk_b200_group_generate_cws_tables",
    "hulk_b200_group_push_reads", "hulk_b200_group_push_reads_fixed", "hulk_b200_group_sync_inputs",
    "hulk_b200_group_flush", "hulk_b200_group_sync", "hulk_b200_group_finish", "hulk_b200_group_reset",
    "hulk_b200_group_get_stats", "hulk_b200_group_sketch_reader",
    "hulk_b200_generate_cws_tables_device", "hulk_b200_get_cws_tables", "hulk_b200_group_generate_cws_tables_device",
    "hulk_b200_packed_bytes", "hulk_b200_pack_bases", "hulk_b200_push_reads_packed", "hulk_b200_set_input_packing",
    "hulk_b200_minhash_enable", "hulk_b200_get_khf", "hulk_b200_get_kmv",
    "hulk_b200_group_minhash_enable", "hulk_b200_group_get_khf", "hulk_b200_group_get_kmv",
]
PEER_HANDLE_BYTES = 64

OK, EW, EK, EEMPTYSEQ, ESHORTSEQ, ESPARSE = 0, -1, -2, -3, -4, -6
EHSK, EDECAY, EBINS, ENEGBINS, ENOSKETCH = -10, -11, -12, -13, -14
EARG, ESTATE, ECUDA, ENOMEM, EIO = -20, -21, -30, -31, -40
EFASTQ, ETOOLONG, ENOSEQ = -41, -42, -43
LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)
F_ASYNC_INPUT = 1
F_INPUT_READY = 2
F_PACK_INPUT = 4


class Params(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("w", C.c_uint32), ("sketch_size", C.c_uint32), ("num_bins", C.c_int32),
        ("decay_ratio", C.c_double), ("device", C.c_int32), ("slot_begin", C.c_uint32),
        ("slot_end", C.c_uint32), ("stream", C.c_void_p), ("flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_reads", "n_bases", "n_minimizers", "n_flushes", "n_adds", "n_kernel_launches", "n_rescans",
        "h2d_bytes", "d2h_bytes", "pack_ns", "n_packed_batches")]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("launches", C.c_uint64 * 4)]


_lib = None


def load():
    """Load libhulk_b200.so (building it is __graft_entry__.build()'s / hulk_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -m hulk_b200.build). "
            "hulk_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    sig = {
        "hulk_b200_version": (C.c_char_p, []),
        "hulk_b200_strerror": (C.c_char_p, [C.c_int]),
        "hulk_b200_last_error": (C.c_char_p, [vp]),
        "hulk_b200_create": (C.c_int, [C.POINTER(Params), C.POINTER(vp)]),
        "hulk_b200_destroy": (None, [vp]),
        "hulk_b200_set_cws_tables": (C.c_int, [vp, vp, vp, vp]),
        "hulk_b200_generate_cws_tables": (C.c_int, [vp]),
        "hulk_b200_generate_cws_tables_async": (C.c_int, [vp]),
        "hulk_b200_set_cws_tables_device": (C.c_int, [vp, vp, vp, vp]),
        "hulk_b200_reset": (C.c_int, [vp]),
        "hulk_b200_profile_enable": (C.c_int, [vp, C.c_int]),
        "hulk_b200_set_overlap": (C.c_int, [vp, C.c_int]),
        "hulk_b200_profile_timeline": (C.c_int, [vp, vp, u64, C.POINTER(u64)]),
        "hulk_b200_profile_read": (C.c_int, [vp, C.POINTER(Profile)]),
        "hulk_b200_new_cws": (C.c_int, [u32, i32, u32, u32, vp, vp, vp]),
        "hulk_b200_new_cws_parallel": (C.c_int, [u32, i32, u32, u32, vp, vp, vp, u32, u64]),
        "hulk_b200_push_reads": (C.c_int, [vp, vp, vp, u64]),
        "hulk_b200_push_reads_fixed": (C.c_int, [vp, vp, u64, u32]),
        "hulk_b200_push_reads_device": (C.c_int, [vp, vp, vp, u64, u32]),
        "hulk_b200_sync_inputs": (C.c_int, [vp]),
        "hulk_b200_flush": (C.c_int, [vp]),
        "hulk_b200_sync": (C.c_int, [vp]),
        "hulk_b200_finish": (C.c_int, [vp, vp, vp]),
        "hulk_b200_snapshot_async": (C.c_int, [vp, vp, vp]),
        "hulk_b200_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "hulk_b200_histogram_device_ptr": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i32)]),
        "hulk_b200_stream": (C.c_int, [vp, C.POINTER(vp)]),
        "hulk_b200_merge_histogram": (C.c_int, [vp, vp]),
        "hulk_b200_add_minimizer_count": (C.c_int, [vp, u64]),
        "hulk_b200_get_histogram": (C.c_int, [vp, vp]),
        "hulk_b200_get_estimates": (C.c_int, [vp, vp]),
        "hulk_b200_get_cms": (C.c_int, [vp, vp]),
        "hulk_b200_minimizers": (C.c_int, [vp, vp, vp, u64, vp, u32, vp]),
        "hulk_b200_jump_hash": (C.c_int, [vp, vp, u64, i32, vp]),
        "hulk_b200_jump_hash_fx": (C.c_int, [vp, vp, u64, i32, vp, vp]),
        "hulk_b200_rcp_selftest": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
        "hulk_b200_get_folded_table": (C.c_int, [vp, vp, C.POINTER(u64)]),
        "hulk_b200_md5_mins": (None, [vp, u32, C.c_char_p]),
        "hulk_b200_sketch_json": (C.c_int64, [C.c_char_p, u64, C.c_char_p, C.c_char_p, u32, vp, vp, u32, i32, C.c_int]),
        "hulk_b200_write_json": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, u32, vp, vp, u32, i32, C.c_int]),
        "hulk_b200_sketch_json_minhash": (C.c_int64, [C.c_char_p, u64, C.c_char_p, C.c_char_p, u32, vp, vp, u32, i32,
                                                      C.c_int, vp, u32, vp, u32]),
        "hulk_b200_write_json_minhash": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, u32, vp, vp, u32, i32, C.c_int,
                                                   vp, u32, vp, u32]),
        "hulk_b200_alloc_pinned": (C.c_int, [C.POINTER(vp), u64]),
        "hulk_b200_free_pinned": (None, [vp]),
        "hulk_b200_reader_open": (C.c_int, [C.POINTER(C.c_char_p), u32, C.c_int, u64, C.POINTER(vp)]),
        "hulk_b200_reader_next": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "hulk_b200_reader_error": (C.c_char_p, [vp]),
        "hulk_b200_reader_close": (None, [vp]),
        "hulk_b200_sketch_reader": (C.c_int, [vp, vp, u64, LOG_FN, vp]),
        "hulk_b200_sketch_load": (C.c_int, [C.c_char_p, C.POINTER(vp), C.c_char_p, u64]),
        "hulk_b200_sketch_find": (C.c_int, [vp, u32, C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32), C.c_char_p, u64]),
        "hulk_b200_sketch_banner": (C.c_char_p, [vp]),
        "hulk_b200_sketch_free": (None, [vp]),
        "hulk_b200_smash": (C.c_int, [vp, vp, u32, u32, C.c_int, i32, vp]),
        "hulk_b200_peer_export": (C.c_int, [vp, vp]),
        "hulk_b200_peer_flags": (C.c_int, [vp, vp]),
        "hulk_b200_peer_connect": (C.c_int, [vp, u32, u32, vp]),
        "hulk_b200_peer_connect_local": (C.c_int, [vp, u32, u32, C.POINTER(vp)]),
        "hulk_b200_group_create": (C.c_int, [C.POINTER(Params), C.POINTER(i32), u32, C.POINTER(vp)]),
        "hulk_b200_group_destroy": (None, [vp]),
        "hulk_b200_group_last_error": (C.c_char_p, [vp]),
        "hulk_b200_group_size": (u32, [vp]),
        "hulk_b200_group_member": (vp, [vp, u32]),
        "hulk_b200_group_set_cws_tables": (C.c_int, [vp, vp, vp, vp]),
        "hulk_b200_group_generate_cws_tables": (C.c_int, [vp, C.c_int]),
        "hulk_b200_group_push_reads": (C.c_int, [vp, vp, vp, u64]),
        "hulk_b200_group_push_reads_fixed": (C.c_int, [vp, vp, u64, u32]),
        "hulk_b200_group_sync_inputs": (C.c_int, [vp]),
        "hulk_b200_group_flush": (C.c_int, [vp]),
        "hulk_b200_group_sync": (C.c_int, [vp]),
        "hulk_b200_group_finish": (C.c_int, [vp, vp, vp]),
        "hulk_b200_group_reset": (C.c_int, [vp]),
        "hulk_b200_group_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "hulk_b200_group_sketch_reader": (C.c_int, [vp, vp, u64, LOG_FN, vp]),
        "hulk_b200_generate_cws_tables_device": (C.c_int, [vp]),
        "hulk_b200_get_cws_tables": (C.c_int, [vp, vp, vp, vp]),
        "hulk_b200_group_generate_cws_tables_device": (C.c_int, [vp]),
        "hulk_b200_packed_bytes": (u64, [u64]),
        "hulk_b200_pack_bases": (C.c_int, [vp, u64, vp, vp, u64, C.POINTER(u64), i32]),
        "hulk_b200_push_reads_packed": (C.c_int, [vp, vp, vp, u64, vp, u64, u32]),
        "hulk_b200_set_input_packing": (C.c_int, [vp, i32]),
        "hulk_b200_minhash_enable": (C.c_int, [vp, C.c_int, C.c_int]),
        "hulk_b200_get_khf": (C.c_int, [vp, vp]),
        "hulk_b200_get_kmv": (C.c_int, [vp, vp, C.POINTER(u32)]),
        "hulk_b200_group_minhash_enable": (C.c_int, [vp, C.c_int, C.c_int]),
        "hulk_b200_group_get_khf": (C.c_int, [vp, vp]),
        "hulk_b200_group_get_kmv": (C.c_int, [vp, vp, C.POINTER(u32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
