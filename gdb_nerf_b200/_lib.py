"""ctypes binding of the C-ABI library (include/gdb_nerf_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load the
import of any op raises.  The library is built in-tree by ``gdb_nerf_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgdbnerf_b200.so")

c_f = C.c_void_p  # device pointers travel as integers
c_i = C.c_int
c_i64 = C.c_int64
c_fl = C.c_float


class RenderTaps(C.Structure):
    _fields_ = [
        ("offsets", C.c_void_p),
        ("S_total", C.c_int64),
        ("rgbs_feat_dir", C.c_void_p),
        ("vox_feat", C.c_void_p),
        ("sigma", C.c_void_p),
        ("feat", C.c_void_p),
        ("weights", C.c_void_p),
    ]


# name -> (restype, argtypes): must list every symbol include/gdb_nerf_b200.h declares
SIGNATURES = {
    "gdb_abi_version": (c_i, []),
    "gdb_last_error_string": (C.c_char_p, []),
    "gdb_mlp_param_floats": (c_i, [c_i]),
    "gdb_planar_to_channels_last": (c_i, [c_f, c_f, c_i, c_i, c_i64, c_i, c_f]),
    "gdb_u8_to_unit_f32": (c_i, [c_f, c_f, c_i64, c_f]),
    "gdb_homography_mats": (c_i, [c_f, c_f, c_f, c_f, c_fl, c_fl, c_i, c_i, c_f, c_f]),
    "gdb_depth_values": (c_i, [c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_warp_variance_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_depth_range_fwd": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f, c_f]),
    "gdb_camera_block": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_bundle_count": (c_i, [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f]),
    "gdb_bundle_scan": (c_i, [c_f, c_f, c_i, c_f, c_f]),
    "gdb_bundle_emit": (c_i, [c_f, c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "gdb_texture_floats": (c_i64, [c_i, c_i, c_i, c_i, c_i]),
    "gdb_prepare_sources": (c_i, [c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f]),
    "gdb_render_fused_fwd": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                   c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_i, c_i, C.POINTER(RenderTaps), c_f]),
    "gdb_depth_range_from_logits_fwd": (c_i, [c_f, c_i, c_i, c_f, c_i64, c_i64, c_i64, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f,
                                              c_f, c_f]),
    "gdb_prob_head_depth_range_split_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "gdb_prob_head_depth_range_tma_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "gdb_prob_head_tma_scratch_floats": (c_i64, [c_i, c_i, c_i, c_i]),
    "gdb_prob_head_tma_counters": (c_i64, [c_i, c_i, c_i]),
    "gdb_prob_head_split_scratch_floats": (c_i64, [c_i, c_i, c_i, c_i]),
    "gdb_prob_head_split_counters": (c_i64, [c_i, c_i, c_i]),
    "gdb_prob_head_depth_range_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f, c_f, c_f]),
    "gdb_warp_variance_bwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f]),
    "gdb_depth_range_bwd": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "gdb_render_fused_bwd": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                   c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    "gdb_prepare_sources_bwd": (c_i, [c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_coarse_render_fwd": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_coarse_render_bwd": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f,
                                    c_f, c_f, c_f, c_f]),
    "gdb_bias_act_add": (c_i, [c_f, c_f, c_f, c_i64, c_i64, c_i, c_i, c_i, c_i, c_i, c_f, c_f]),
    "gdb_gate_add": (c_i, [c_f, c_f, c_f, c_f, c_i64, c_i64, c_i, c_f, c_f]),
    "gdb_se_gate_add": (c_i, [c_f, c_f, c_f, c_f, c_i, c_f, c_i64, c_i64, c_i, c_i, c_f, c_f, c_f]),
    "gdb_se_gate_add_cat": (c_i, [c_f, c_f, c_f, c_f, c_i, c_f, c_i64, c_i64, c_i, c_i, c_f, c_f, c_f, c_i, c_f]),
    "gdb_concat2_into": (c_i, [c_f, c_i, c_f, c_i, c_i64, c_f, c_i, c_i, c_f]),
    "gdb_pixel_shuffle2": (c_i, [c_f, c_f, c_i64, c_i, c_i, c_i, c_f, c_f]),
    "gdb_concat3": (c_i, [c_f, c_i, c_f, c_i, c_f, c_i, c_i64, c_f, c_f]),
    "gdb_channel_mean": (c_i, [c_f, c_i64, c_i64, c_i, c_i, c_f, c_f, c_f]),
    "gdb_adam_clip_step": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i64, c_fl, c_fl, c_fl, c_fl, c_fl, c_fl, c_fl, c_f]),
    "gdb_adam_advance": (c_i, [c_f, c_f]),
    "gdb_assemble_output": (c_i, [c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f]),
}

ABI_VERSION = 2   # include/gdb_nerf_b200.h GDB_ABI_VERSION: bumped whenever an exported signature changes

_lock = threading.Lock()
_lib = None


def _check_fresh() -> None:
    """The library must have been built from the sources next to it: the build stamp (digest of csrc/ + the header, written
    by gdb_nerf_b200.build, git-ignored like the library) is compared with the digest of the sources on disk.  A stale
    binary whose argument lists no longer match the ctypes signatures would otherwise corrupt pointers silently."""
    if os.environ.get("GDB_SKIP_DIGEST_CHECK") == "1":
        return
    from . import build as _build
    if not os.path.isdir(_build.CSRC):          # binary-only deployment: nothing to compare with
        return
    stamp = open(_build.STAMP).read().strip() if os.path.exists(_build.STAMP) else None
    if stamp != _build._digest():
        raise GdbError(f"{LIB_PATH} is stale (built from different sources than gdb_nerf_b200/csrc): "
                       "rebuild with `python -m gdb_nerf_b200.build`")


class GdbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load (once) and type the library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GdbError(
                f"{LIB_PATH} is missing: build it with `python -m gdb_nerf_b200.build` "
                "(there is no CPU or PyTorch fallback for the rendering path)")
        _check_fresh()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        got = lib.gdb_abi_version()
        if got != ABI_VERSION:
            raise GdbError(f"ABI version mismatch: library {got}, binding {ABI_VERSION}; rebuild with `python -m gdb_nerf_b200.build`")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().gdb_last_error_string().decode(errors="replace")
        raise GdbError(f"{what} failed ({code}): {msg}")
