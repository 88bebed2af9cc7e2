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
    assert a[0].shape == (N, 3) and a[1].shape == (N, 1, 3) and a[3].size == N and a[5].shape == (N, 4)
    if a[4].ndim != 2 or a[4].shape[0] != N or a[4].shape[1] not in (1, 3):
        raise ValueError("scales must be (N,3) or, for an isotropic model, (N,1)")
    S = a[4].shape[1]  # isotropic models carry one log-scale: export_ply then writes only scale_0 (gaussians.jl:176)
    _check(_lib.lib().gsr_ply_write_scales(str(path).encode(), N, R, S, _p(a[0]), _p(a[1]), _p(a[2]) if R and N else None,
                                           _p(a[3]), _p(a[4]), _p(a[5])))


# ---- checkpoints (src/checkpoint.jl): safetensors files, structure in the dotted names, scalars in `__metadata__` ----
CHECKPOINT_FORMAT = "GaussianSplatting.jl-checkpoint-1"  # checkpoint.jl:15
_ST_DTYPES = {"F32": np.float32, "F64": np.float64, "F16": np.float16, "I64": np.int64, "I32": np.int32, "I16": np.int16,
              "I8": np.int8, "U8": np.uint8, "U16": np.uint16, "U32": np.uint32, "U64": np.uint64, "BOOL": np.bool_}


def read_safetensors(path: str):
    """The container itself: 8-byte little-endian header length, JSON header {name: {dtype, shape, data_offsets}} with an
    optional `__metadata__` string map, then the payload.  Tensors come back as (memory-mapped) C-order NumPy arrays of
    the stored shape — for a reference checkpoint that is the shape Julia sees (runtests.jl:969-973)."""
    import json
    import struct
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise ValueError(f"`{path}` is not a safetensors file (too short)")
        (hlen,) = struct.unpack("<Q", head)
        size = f.seek(0, 2)
        if hlen <= 0 or 8 + hlen > size:
            raise ValueError(f"`{path}` is not a safetensors file (header length {hlen})")
        f.seek(8)
        try:
            header = json.loads(f.read(hlen).decode("utf-8"))
        except Exception as e:
            raise ValueError(f"`{path}` is not a safetensors file (bad header: {e})") from None
    meta = header.pop("__metadata__", None) or {}
    data = np.memmap(path, np.uint8, "r", offset=8 + hlen) if size > 8 + hlen else np.zeros(0, np.uint8)
    tensors = {}
    for name, d in header.items():
        b, e = d["data_offsets"]
        dt = np.dtype(_ST_DTYPES[d["dtype"]])
        n = int(np.prod(d["shape"], dtype=np.int64)) if d["shape"] else 1
        if e - b != n * dt.itemsize or e > data.size:
            raise ValueError(f"`{path}`: tensor {name} has inconsistent offsets")
        tensors[name] = data[b:e].view(dt).reshape(d["shape"])
    return tensors, meta


def write_safetensors(path: str, tensors: dict, meta: dict | None = None) -> None:
    import json
    import struct
    inv = {np.dtype(v): k for k, v in _ST_DTYPES.items()}
    header, off, blobs = {}, 0, []
    if meta:
        header["__metadata__"] = {str(k): str(v) for k, v in meta.items()}
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name])
        header[name] = {"dtype": inv[a.dtype], "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        off += a.nbytes
        blobs.append(a)
    h = json.dumps(header, separators=(",", ":")).encode("utf-8")
    h += b" " * (-len(h) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for a in blobs:
            f.write(a.tobytes())


_GAUSSIAN_FIELDS = ("points", "features_dc", "features_rest", "scales", "rotations", "opacities")


def load_checkpoint(path: str, prefix: str = "gaussians") -> dict:
    """`load_checkpoint` + `read_state!(::GaussianModel, ckpt, prefix)` (checkpoint.jl:62-70, gaussians.jl:104-116): the
    six raw parameter arrays of the model in this package's layout, `sh_degree` / `max_sh_degree`, and the whole
    metadata map under "meta".  The stored tensors have the shapes Julia sees, in C order — (3,N) is three rows of N —
    so they are transposed here into the (N,3) / (N,K,3) arrays that are byte-identical to Julia's memory."""
    tensors, meta = read_safetensors(path)
    if meta.get("format") != CHECKPOINT_FORMAT:  # checkpoint.jl:65-68
        raise ValueError(f"`{path}` is not a GaussianSplatting.jl checkpoint (no `{CHECKPOINT_FORMAT}` in its metadata).")
    out = {}
    for f in _GAUSSIAN_FIELDS:
        key = f"{prefix}.{f}"
        if key not in tensors:
            raise KeyError(key)
        a = np.asarray(tensors[key], np.float32)
        out[f] = np.ascontiguousarray(a.transpose(*reversed(range(a.ndim))))
    out["sh_degree"] = int(meta[f"{prefix}.sh_degree"])
    out["max_sh_degree"] = int(meta[f"{prefix}.max_sh_degree"])
    out["meta"] = dict(meta)
    return out


def save_checkpoint(path: str, model: dict, prefix: str = "gaussians", meta: dict | None = None, extra: dict | None = None):
    """`write_state!(tensors, meta, prefix, ::GaussianModel)` + `save_checkpoint` (gaussians.jl:91-102, checkpoint.jl:44-54)."""
    tensors = dict(extra or {})
    for f in _GAUSSIAN_FIELDS:
        a = np.asarray(model[f], np.float32)
        tensors[f"{prefix}.{f}"] = np.ascontiguousarray(a.transpose(*reversed(range(a.ndim))))
    m = {str(k): str(v) for k, v in (meta or {}).items()}
    m[f"{prefix}.sh_degree"] = str(int(model["sh_degree"]))
    m[f"{prefix}.max_sh_degree"] = str(int(model["max_sh_degree"]))
    m["format"] = CHECKPOINT_FORMAT
    write_safetensors(path, tensors, m)
