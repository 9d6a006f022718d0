"""ctypes binding of liboptex_b200.so (the C-ABI declared in include/optex_b200.h).

The library is the product: if it is missing or the device is not a B200 every call raises -
there is no CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "liboptex_b200.so")

OK, EINVAL, EDEVICE, ECUDA, EWORKSPACE, ESIZE, EUNSUPPORTED = range(7)
MODES = {"chol": 0, "pca": 1, "sym": 2, "cdf": 3, "sort": 4}
GEMM_MODES = {"auto": 0, "fp32": 1, "tf32x3": 2, "tf32": 3}

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_z = C.c_size_t
_u = C.c_uint64

# name -> (restype, argtypes); must list every symbol of include/optex_b200.h (tests check it)
SIGNATURES = {
    "optex_abi_version": (_i, []),
    "optex_last_error": (C.c_char_p, []),
    "optex_device_check": (_i, []),
    "optex_launch_count": (_u, []),
    "optex_set_gemm_mode": (_i, [_i]),
    "optex_get_gemm_mode": (_i, []),
    "optex_set_pdl": (_i, [_i]),
    "optex_set_scratch_slot": (_i, [_i]),
    "optex_set_rotation_precision": (_i, [_i]),
    "optex_get_rotation_precision": (_i, []),
    "optex_debug_gemm_trace": (_i, [_p]),
    "optex_ot_workspace_bytes": (_z, [_l, _l, _i, _i]),
    "optex_ot_step": (_i, [_p, _p, _p, _p, _i, _l, _i, _l, _i, _i, _f, _p, _f, _p, _z, _p]),
    "optex_ot_step_host": (_i, [_p, _p, _p, _p, _i, _l, _i, _l, _i, _i, _f, _p, _f, _u, _u, _p]),
    "optex_ot_step_host_async": (_i, [_p, _p, _p, _p, _i, _l, _i, _l, _i, _i, _f, _p, _f, _u, _u, _i, _p]),
    "optex_ot_host_set_style": (_i, [_p, _i, _l, _i, _p]),
    "optex_ot_steps_workspace_bytes": (_z, [_l, _l, _i, _i]),
    "optex_ot_steps": (_i, [_p, _p, _i, _p, _p, _u, _u, _p, _i, _i, _i, _i, _l, _i, _l, _i, _i, _f, _p, _z, _p]),
    "optex_split_rotations": (_i, [_p, _i, _i, _p, _p]),
    "optex_fence": (_i, [_p]),
    "optex_ot_step_profile": (_i, [_p, _p, _p, _p, _i, _l, _i, _l, _i, _i, _f, _p, _z, _p, _p, _p, _p]),
    "optex_comm_unique_id_bytes": (_z, []),
    "optex_comm_unique_id": (_i, [_p, _z]),
    "optex_comm_init": (_i, [_p, _i, _i, _p]),
    "optex_comm_adopt": (_i, [_p, _i, _i, _p]),
    "optex_comm_destroy": (_i, [_p]),
    "optex_comm_rank": (_i, [_p]),
    "optex_comm_world": (_i, [_p]),
    "optex_comm_allgather_f32": (_i, [_p, _p, _p, _z, _p]),
    "optex_ot_step_sharded": (_i, [_p, _p, _p, _p, _p, _l, _l, _l, _l, _i, _i, _f, _p, _f, _p, _z, _p]),
    "optex_ot_loop": (_i, [_p, _p, _p, _i, _u, _u, _i, _l, _i, _l, _i, _i, _f, _p, _f, _p, _z, _p]),
    "optex_ot_loop_workspace_bytes": (_z, [_l, _l, _i, _i]),
    "optex_hist_match_workspace_bytes": (_z, [_l, _l, _i, _i]),
    "optex_hist_match": (_i, [_p, _p, _p, _i, _l, _i, _l, _i, _i, _f, _p, _z, _p]),
    "optex_cdf_match_workspace_bytes": (_z, [_i, _i]),
    "optex_cdf_match": (_i, [_p, _p, _p, _i, _l, _l, _i, _p, _p, _z, _p]),
    "optex_interp": (_i, [_p, _p, _p, _p, _l, _i, _p]),
    "optex_sort_match_workspace_bytes": (_z, [_i, _l, _l]),
    "optex_sort_match": (_i, [_p, _p, _p, _i, _l, _l, _p, _p, _z, _p]),
    "optex_rotation_workspace_bytes": (_z, [_i]),
    "optex_random_rotation": (_i, [_p, _i, _u, _u, _p, _p, _z, _p]),
    "optex_rotation_prepare_workspace_bytes": (_z, [_i]),
    "optex_rotation_prepare": (_i, [_p, _i, _p, _z, _p]),
    "optex_rotations_workspace_bytes": (_z, [_i, _i]),
    "optex_random_rotations": (_i, [_p, _i, _i, _u, _u, _p, _p, _z, _p]),
    "optex_rotate_forward": (_i, [_p, _p, _p, _l, _i, _p]),
    "optex_rotate_forward_block": (_i, [_p, _p, _p, _l, _i, _i, _i, _p]),
    "optex_rotate_inverse": (_i, [_p, _p, _p, _l, _i, _p, _f, _p]),
    "optex_fit_pca_workspace_bytes": (_z, [_l, _i]),
    "optex_fit_pca": (_i, [_p, _l, _i, _p, _p, _p, _p, _z, _p]),
    "optex_fit_pca_warm": (_i, [_p, _l, _i, _p, _p, _p, _p, _i, _p, _p, _z, _p]),
    "optex_debug_pca_stamps": (_i, [_p, _i]),
    "optex_debug_chain_stamps": (_i, [_p, _i]),
    "optex_pca_project": (_i, [_p, _p, _p, _l, _i, _i, _i, _p]),
    "optex_conv3x3_packed_k": (_i, [_i]),
    "optex_conv3x3_workspace_bytes": (_z, [_i, _i, _i, _i, _i, _i]),
    "optex_conv3x3": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _l, _p, _z, _p]),
    "optex_nhwc_to_nchw": (_i, [_p, _p, _i, _l, _i, _i, _p]),
    "optex_resize_workspace_bytes": (_z, [_i, _i, _i, _i, _i]),
    "optex_resize_bicubic_aa": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _z, _p]),
    "optex_rgb_to_hls": (_i, [_p, _p, _i, _l, _p]),
    "optex_hls_to_rgb": (_i, [_p, _p, _i, _l, _p]),
    "optex_lightness_transfer": (_i, [_p, _p, _p, _i, _l, _p]),
    "optex_mix_features": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p]),
    "optex_recentre_workspace_bytes": (_z, []),
    "optex_recentre": (_i, [_p, _l, _p, _l, _p, _p, _z, _p]),
}

_lib = None


class OptexError(RuntimeError):
    """A liboptex_b200 call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"liboptex_b200 status {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built - no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m optimaltextures_b200.build` "
                "(nvcc, sm_100a). optimaltextures_b200 has no CPU or PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.optex_abi_version() != 1:
            raise RuntimeError("liboptex_b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code == OK:
        return
    msg = lib().optex_last_error().decode(errors="replace")
    if code in (EINVAL, ESIZE):
        raise ValueError(f"liboptex_b200 status {code}: {msg}")
    if code == EUNSUPPORTED:
        raise NotImplementedError(f"liboptex_b200 status {code}: {msg}")
    raise OptexError(code, msg)


def mode_id(mode: str) -> int:
    """hist_mode string -> enum.  Like the reference (histmatch.py:36, `else:  # sym`) any string
    that is not chol / pca / cdf / sort selects the sym branch."""
    return MODES.get(mode, MODES["sym"])
