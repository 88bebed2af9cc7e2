"""ctypes binding of libgsrast.so (include/gsrast.h).  1:1 with the `ccall`s of julia/GsrastCUDAExt.jl.

The library is built in-tree (csrc/Makefile, nvcc sm_100a).  There is no fallback: if it cannot be
loaded, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
LIB_PATH = os.path.join(CSRC, "libgsrast.so")

GSR_OK, GSR_EINVAL, GSR_ECUDA, GSR_ENOMEM, GSR_ESTATE = 0, -1, -2, -3, -4
MATH_REFERENCE, MATH_FAST, MATH_STRICT, MATH_EXPERIMENT = 0, 1, 2, 256


class GsrConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("near_plane", C.c_float),
                ("far_plane", C.c_float), ("radius_clip", C.c_int32), ("blur_eps", C.c_float),
                ("math_mode", C.c_int32)]


class GsrCamera(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("focal", C.c_float * 2), ("principal", C.c_float * 2),
                ("cam_center", C.c_float * 3), ("R_dev", C.c_void_p), ("t_dev", C.c_void_p)]


class GsrStateViews(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_rendered", C.c_int64)] + [(k, C.c_void_p) for k in (
        "radii", "grad_means2d", "means2d", "depths", "conics", "rgbs", "clamped", "tiles_touched", "points_offset",
        "normals", "keys_unsorted", "values_unsorted", "keys_sorted", "values_sorted", "ranges", "n_contrib",
        "accum_alpha")]


EXPORTS = ["gsr_version", "gsr_last_error", "gsr_create", "gsr_destroy", "gsr_release_scene_buffers",
           "gsr_memory_usage", "gsr_get_state", "gsr_forward_generation", "gsr_forward", "gsr_backward", "gsr_update_stats",
           "gsr_forward_backward_host", "gsr_identify_tile_range", "gsr_sort_pairs", "gsr_launch_count",
           "gsr_profile_enable", "gsr_profile_get", "gsr_measure_fp32_peak", "gsr_debug_exp_neg", "gsr_forward_backward_host_async",
           "gsr_host_wait", "gsr_host_timeline", "gsr_set_accumulator", "gsr_backward_render",
           "gsr_backward_gaussians_peers", "gsr_backward_gaussians_views", "gsr_export_accumulator", "gsr_ssim_forward", "gsr_ssim_backward", "gsr_photometric_loss", "gsr_forward_raw", "gsr_backward_raw", "gsr_ply_open", "gsr_ply_read",
           "gsr_ply_close", "gsr_ply_write", "gsr_ply_write_scales", "gsr_ply_last_error", "gsr_densify_masks", "gsr_prune_mask",
           "gsr_mask_offsets_scratch_words", "gsr_mask_offsets", "gsr_gather_rows", "gsr_split_children"]
STAGES = ["preprocess", "scan", "duplicate", "sort", "ranges", "render_fwd", "zero_grads", "render_bwd", "gauss_bwd",
          "presort"]


HASH_PATH = os.path.join(CSRC, "libgsrast.srchash")


def _source_hash() -> str:
    """sha256 over every file the library is built from (names + contents).  Content, not mtime: the snapshot that
    travels to a GPU box does not preserve timestamps, and a rebuild there would burn GPU minutes."""
    import hashlib
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h")) or f == "Makefile")
    files.append(os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "gsrast.h"))
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compile libgsrast.so for sm_100a with nvcc (cross-compiles without a GPU) when it is missing or was built from
    different sources (.cu / .cuh / .cpp / Makefile / include/gsrast.h)."""
    want = _source_hash()
    have = open(HASH_PATH).read().strip() if os.path.exists(HASH_PATH) else None
    if force or not os.path.exists(LIB_PATH) or have != want:
        cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libgsrast.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
        with open(HASH_PATH, "w") as fh:
            fh.write(want + "\n")
    return LIB_PATH


def load() -> C.CDLL:
    override = os.environ.get("GSRAST_LIB")  # A/B builds of the same sources with other -D flags (tools/ only)
    if not override:
        build()  # no-op unless the library is missing or stale with respect to its sources
    lib = C.CDLL(override or LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.gsr_version.restype = C.c_char_p
    lib.gsr_last_error.restype = C.c_char_p
    lib.gsr_last_error.argtypes = [vp]
    lib.gsr_create.argtypes = [C.POINTER(GsrConfig), C.POINTER(vp)]
    lib.gsr_destroy.argtypes = [vp]
    lib.gsr_release_scene_buffers.argtypes = [vp]
    lib.gsr_memory_usage.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.gsr_get_state.argtypes = [vp, C.POINTER(GsrStateViews)]
    lib.gsr_forward_generation.argtypes = [vp]
    lib.gsr_forward_generation.restype = i64
    lib.gsr_forward.argtypes = [vp, C.POINTER(GsrCamera), i64, i32, i32, vp, vp, vp, vp, vp, C.POINTER(C.c_float),
                                vp, vp, vp, C.POINTER(i64), vp]
    lib.gsr_backward.argtypes = [vp, C.POINTER(GsrCamera), i64, i32, i32, vp, vp, vp, vp, vp, C.POINTER(C.c_float),
                                 vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.gsr_update_stats.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.gsr_forward_raw.argtypes = [vp, C.POINTER(GsrCamera), i64, i32, i32, vp, vp, vp, vp, vp, i32, vp,
                                    C.POINTER(C.c_float), vp, vp, vp, C.POINTER(i64), vp]
    lib.gsr_backward_raw.argtypes = [vp, C.POINTER(GsrCamera), i64, i32, i32, vp, vp, vp, vp, vp, i32, vp,
                                     C.POINTER(C.c_float), vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.gsr_forward_backward_host.argtypes = [vp, C.POINTER(GsrCamera), i64, i32, i32, vp, vp, vp, vp, vp,
                                              C.POINTER(C.c_float), vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64), vp]
    lib.gsr_forward_backward_host_async.argtypes = lib.gsr_forward_backward_host.argtypes
    lib.gsr_host_wait.argtypes = [vp]
    lib.gsr_host_timeline.argtypes = [vp, C.POINTER(C.c_float)]
    lib.gsr_set_accumulator.argtypes = [vp, vp, i64]
    lib.gsr_backward_render.argtypes = [vp, i64, C.POINTER(C.c_float), vp, vp]
    lib.gsr_backward_gaussians_peers.argtypes = [vp, i32, i32, C.POINTER(GsrCamera), C.POINTER(vp), C.POINTER(vp), i64, i32,
                                                 i32, vp, vp, vp, vp, vp, vp]
    lib.gsr_backward_gaussians_views.argtypes = [vp, i32, C.POINTER(GsrCamera), C.POINTER(vp), i32, i32, i32, C.POINTER(vp),
                                                 i64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.gsr_export_accumulator.argtypes = [vp, i64, vp, vp]
    lib.gsr_ssim_forward.argtypes = [i32, i32, i32, i32, vp, vp, C.c_float, C.c_float, i32, vp, vp, vp, vp, vp]
    lib.gsr_ssim_backward.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.gsr_photometric_loss.argtypes = [vp, vp, vp, C.c_float, vp, vp, vp]
    lib.gsr_ply_open.argtypes = [C.c_char_p, C.POINTER(i64), C.POINTER(i32), C.POINTER(vp)]
    lib.gsr_ply_read.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.gsr_ply_close.argtypes = [vp]
    lib.gsr_ply_close.restype = None
    lib.gsr_ply_write.argtypes = [C.c_char_p, i64, i32, vp, vp, vp, vp, vp, vp]
    lib.gsr_ply_write_scales.argtypes = [C.c_char_p, i64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.gsr_ply_last_error.restype = C.c_char_p
    lib.gsr_densify_masks.argtypes = [i64, i64, vp, vp, vp, i32, C.c_float, C.c_float, vp, vp, vp]
    lib.gsr_prune_mask.argtypes = [i64, vp, vp, i32, vp, C.c_float, i32, C.c_float, vp, vp]
    lib.gsr_mask_offsets_scratch_words.argtypes = [i64]
    lib.gsr_mask_offsets_scratch_words.restype = C.c_size_t
    lib.gsr_mask_offsets.argtypes = [i64, vp, vp, vp, vp, vp]
    lib.gsr_gather_rows.argtypes = [i64, i32, vp, vp, vp, vp, i32, i64, vp]
    lib.gsr_split_children.argtypes = [i64, vp, vp, i32, vp, vp, i32, vp]
    lib.gsr_identify_tile_range.argtypes = [vp, i64, vp, vp]
    lib.gsr_sort_pairs.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    lib.gsr_launch_count.restype = i64
    lib.gsr_profile_enable.argtypes = [vp, i32]
    lib.gsr_profile_get.argtypes = [vp, C.POINTER(C.c_float)]
    lib.gsr_measure_fp32_peak.argtypes = [C.POINTER(C.c_double), vp]
    lib.gsr_debug_exp_neg.argtypes = [vp, vp, vp, vp, i64, vp]
    for name in EXPORTS:
        getattr(lib, name)  # every symbol of include/gsrast.h must resolve
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


class GsrError(RuntimeError):
    pass


def check(rc: int, handle=None):
    if rc != GSR_OK:
        msg = lib().gsr_last_error(handle)
        raise GsrError(f"libgsrast error {rc}: {msg.decode() if msg else ''}")
