"""ctypes binding of libtscode_b200.so — the C-ABI declared in include/tscode_b200.h.

The library is the product's only compute path.  If it is missing, or no CUDA device is
visible, every entry point raises: there is no CPU fallback (BASELINE.json north star).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtscode_b200.so")
_LIB = None

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/tscode_b200.h one to one
SIGNATURES = {
    "tsc_version": (C.c_int, []),
    "tsc_error_string": (C.c_char_p, [C.c_int]),
    "tsc_num_blocks_padded": (_i64, [_i64]),
    "tsc_num_slabs": (_i32, [_i32]),
    "tsc_packed_doubles": (_i64, [_i64, _i32]),
    "tsc_device_sm_count": (_i32, []),
    "tsc_pack": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "tsc_pack_blocks": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i64, _i64, _vp]),
    "tsc_rmsd_sim_tiles": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _f64, _vp, _i32, _i32, _vp]),
    "tsc_screen_operand_bytes": (_i64, [_i64, _i32]),
    "tsc_screen_ct_floats": (_i64, [_i64]),
    "tsc_screen_rows_padded": (_i64, [_i64]),
    "tsc_host_sample_pairs": (None, [_i64, _i32, _vp, _vp]),
    "tsc_host_cluster_rejects": (_i64, [_vp, _vp, _i64, _i64, _vp]),
    "tsc_host_ladder_replay": (_i64, [_i64, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tsc_host_screen_plan": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _f64, _vp, _vp, _vp]),
    "tsc_screen_max_atoms": (_i32, [_i32]),
    "tsc_pack_screen": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "tsc_rmsd_screen": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "tsc_rmsd_verify": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "tsc_rmsd_verify_incr": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp]),
    "tsc_elim_fused_ws_words": (_i64, [_i64]),
    "tsc_elim_fused_out_bytes": (_i64, [_i64]),
    "tsc_elim_fused": (C.c_int, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _vp]),
    "tsc_elim_fused_p2p": (C.c_int, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _vp]),
    "tsc_pairs_push": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tsc_rmsd_pairs": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "tsc_rmsd_pairs_idx": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _f64, _vp, _vp]),
    "tsc_group_greedy": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tsc_elim_cachebits": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "tsc_elim_round": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "tsc_elim_commit": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsc_tfd_fingerprints": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "tsc_tfd_scan": (C.c_int, [_vp, _i64, _i32, _f64, _vp, _vp, _vp]),
    "tsc_moi_moments": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "tsc_moi_scan": (C.c_int, [_vp, _i64, _f64, _vp, _vp, _vp]),
    "tsc_constraint_scores": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "tsc_embed_clash": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _f64, _f64, _i64, _vp, _vp, _vp]),
    "tsc_clash_structs": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _f64, _f64, _f64, _i64, _vp, _vp, _vp]),
    "tsc_rotcorr_pairs": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f64,
                                    _vp, _vp, _vp, _vp, _vp]),
    "tsc_host_screen_items": (_i64, [_i64, _vp, _i64, _i32, _i64, _i64, _f64, _i32, _i32, _i32, _vp, _i64]),
    "tsc_host_read_xyz": (_i64, [C.c_char_p, _i64, _vp, _vp, _i64, _vp, _vp, _i32]),
    "tsc_host_centre": (_i32, [_vp, _i64, _i32, _vp, _i32]),
    "tsc_host_write_xyz": (_i64, [_vp, _i64, _i32, C.c_char_p, C.c_char_p, _vp, _i64, _i32]),
    "tsc_host_rotcorr_chunk": (_i64, [_i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tsc_rotcorr_scan": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f64,
                                   _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsc_rotcorr_row": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32,
                                  _vp, _vp, _vp, _vp]),
    "tsc_rotcorr_commit": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "tsc_rotcorr_apply": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsc_string_embed_params": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                          _vp, _vp]),
    "tsc_cyclical_embed_params": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _vp]),
    "tsc_embed_gather": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
}


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises ExtensionMissing if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} not found: build it with `python -m tscode_b200.csrc.build` "
                "(nvcc, sm_100a).  tscode_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().tsc_error_string(rc)
        raise RuntimeError(f"tscode_b200 CUDA error {rc} in {what}: {msg.decode() if msg else '?'}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tscode_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib()
    return torch


def ptr(t):
    """Address of a torch tensor as c_void_p; None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
