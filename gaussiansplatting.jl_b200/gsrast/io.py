"""3DGS `.ply` scenes in and out (import_ply / export_ply, src/gaussians.jl:157-247) over libgsrast's native reader.

Arrays are NumPy float32 in the layout the rest of the package uses (C-contiguous == the reference's column-major with
reversed dims): points (N,3), features_dc (N,1,3), features_rest (N,R,3), opacities (N,1) pre-sigmoid, scales (N,3)
log-scales, rotations (N,4) wxyz — i.e. exactly what `GaussianRasterizer.__call__` takes once moved to the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise _lib.GsrError(f"libgsrast error {rc}: {_lib.lib().gsr_ply_last_error().decode()}")


def load_ply(path: str) -> dict:
    lib = _lib.lib()
    n, r, rd = C.c_int64(0), C.c_int32(0), C.c_void_p()
    _check(lib.gsr_ply_open(str(path).encode(), C.byref(n), C.byref(r), C.byref(rd)))
    try:
        N, R = int(n.value), int(r.value)
        out = dict(points=np.empty((N, 3), np.float32), features_dc=np.empty((N, 1, 3), np.float32),
                   features_rest=np.empty((N, R, 3), np.float32), opacities=np.empty((N, 1), np.float32),
                   scales=np.empty((N, 3), np.float32), rotations=np.empty((N, 4), np.float32))
        _check(lib.gsr_ply_read(rd, _p(out["points"]), _p(out["features_dc"]), _p(out["features_rest"]) if R else None,
                                _p(out["opacities"]), _p(out["scales"]), _p(out["rotations"])))
    finally:
        lib.gsr_ply_close(rd)
    out["max_sh_degree"] = int(round(np.sqrt(R + 1))) - 1  # gaussians.jl:236
    return out


def save_ply(path: str, points, features_dc, features_rest, opacities, scales, rotations) -> None:
    a = [np.ascontiguousarray(x, np.float32) for x in (points, features_dc, features_rest, opacities, scales, rotations)]
    N = a[0].shape[0]
    R = a[2].shape[1] if a[2].ndim == 3 else 0
    assert a[0].shape == (N, 3) and a[1].shape == (N, 1, 3) and a[3].size == N and a[4].shape == (N, 3) and a[5].shape == (N, 4)
    _check(_lib.lib().gsr_ply_write(str(path).encode(), N, R, _p(a[0]), _p(a[1]), _p(a[2]) if R and N else None, _p(a[3]), _p(a[4]),
                                    _p(a[5])))
