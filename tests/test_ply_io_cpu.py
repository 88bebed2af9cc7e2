"""the native 3DGS `.ply` reader / writer (csrc/ply_io.cpp) against a NumPy restatement of
export_ply / import_ply (src/gaussians.jl:157-247).  Host-only: runs without a GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _io():
    from gsrast import io
    return io


def model(n, R, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
    return dict(points=f(n, 3), features_dc=f(n, 1, 3), features_rest=f(n, R, 3), opacities=f(n, 1), scales=f(n, 3),
                rotations=f(n, 4))


def names(R):
    return (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(3 * R)] +
            ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])


def numpy_rows(m):
    """export_ply's property matrix, one row per vertex (gaussians.jl:160-186): f_rest channel-major."""
    n, R = m["points"].shape[0], m["features_rest"].shape[1]
    rest = m["features_rest"].transpose(0, 2, 1).reshape(n, 3 * R)  # (N,R,3)[k][c] -> c*R + k
    return np.concatenate([m["points"], np.zeros((n, 3), np.float32), m["features_dc"].reshape(n, 3), rest, m["opacities"],
                           m["scales"], m["rotations"]], 1).astype(np.float32)


def numpy_write(path, m, order=None, dtype="<f4", fmt="binary_little_endian", type_name="float", extra_element=False):
    rows = numpy_rows(m)
    nm = names(m["features_rest"].shape[1])
    order = list(range(len(nm))) if order is None else order
    with open(path, "wb") as f:
        hdr = ["ply", f"format {fmt} 1.0", "comment written by the test", f"element vertex {rows.shape[0]}"]
        hdr += [f"property {type_name} {nm[j]}" for j in order]
        if extra_element:
            hdr += ["element face 0", "property list uchar int vertex_indices"]
        hdr += ["end_header"]
        f.write(("\n".join(hdr) + "\n").encode())
        if fmt == "ascii":
            for r in rows[:, order]:
                f.write((" ".join(repr(float(v)) for v in r) + "\n").encode())
        else:
            f.write(rows[:, order].astype(dtype).tobytes())


def numpy_read(path):
    """import_ply restated: parse the header, gather properties by name."""
    raw = open(path, "rb").read()
    head, data = raw.split(b"end_header\n", 1)
    props = [l.split()[2].decode() for l in head.split(b"\n") if l.startswith(b"property")]
    n = int([l for l in head.split(b"\n") if l.startswith(b"element vertex")][0].split()[2])
    a = np.frombuffer(data, "<f4", n * len(props)).reshape(n, len(props))
    col = {p: a[:, i] for i, p in enumerate(props)}
    R = sum(p.startswith("f_rest_") for p in props) // 3
    rest = np.stack([col[f"f_rest_{j}"] for j in range(3 * R)], 1).reshape(n, 3, R).transpose(0, 2, 1) if R else np.zeros((n, 0, 3), np.float32)
    return dict(points=np.stack([col[k] for k in "xyz"], 1), features_dc=np.stack([col[f"f_dc_{i}"] for i in range(3)], 1)[:, None],
                features_rest=rest, opacities=col["opacity"][:, None], scales=np.stack([col[f"scale_{i}"] for i in range(3)], 1),
                rotations=np.stack([col[f"rot_{i}"] for i in range(4)], 1)), props


def same(a, b):
    for k in ("points", "features_dc", "features_rest", "opacities", "scales", "rotations"):
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(np.asarray(a[k], np.float32), np.asarray(b[k], np.float32)), k


@pytest.mark.parametrize("n,R", [(1000, 15), (7, 0), (3, 3), (0, 8), (5000, 8)])
def test_native_write_matches_reference_layout_and_round_trips(tmp_path, n, R):
    io = _io()
    m = model(n, R, n + R)
    p = str(tmp_path / "a.ply")
    io.save_ply(p, **m)
    got, props = numpy_read(p)
    assert props == names(R)  # property order of export_ply
    same(got, m)
    back = io.load_ply(p)
    same(back, m)
    assert back["max_sh_degree"] == int(round(np.sqrt(R + 1))) - 1
    # byte-identical to the restated export_ply (minus the test's comment line)
    q = str(tmp_path / "b.ply")
    numpy_write(q, m)
    assert open(p, "rb").read() == open(q, "rb").read().replace(b"comment written by the test\n", b"")


def test_reader_keys_off_names_not_order_or_precision(tmp_path):  # gaussians.jl:199-203
    io = _io()
    m = model(257, 15, 3)
    nm = names(15)
    order = list(np.random.default_rng(1).permutation(len(nm)))
    for kw in (dict(order=order), dict(dtype="<f8", type_name="double"), dict(dtype=">f4", fmt="binary_big_endian"),
               dict(fmt="ascii"), dict(extra_element=True), dict(type_name="float32")):
        p = str(tmp_path / "c.ply")
        numpy_write(p, m, **kw)
        same(io.load_ply(p), m)


def test_reader_errors(tmp_path):
    io = _io()
    from gsrast import _lib
    m = model(10, 15, 0)
    with pytest.raises(_lib.GsrError, match="cannot open"):
        io.load_ply(str(tmp_path / "missing.ply"))
    p = str(tmp_path / "bad.ply")
    nm = names(15)
    numpy_write(p, m, order=[j for j in range(len(nm)) if nm[j] != "f_rest_44"])  # 44 f_rest: not a multiple of 3
    with pytest.raises(_lib.GsrError, match="whole number"):
        io.load_ply(p)
    numpy_write(p, m, order=[j for j in range(len(nm)) if nm[j] != "opacity"])
    with pytest.raises(_lib.GsrError, match="missing property opacity"):
        io.load_ply(p)
    numpy_write(p, m)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-100])
    with pytest.raises(_lib.GsrError, match="truncated"):
        io.load_ply(p)
    open(p, "wb").write(b"not a ply\n")
    with pytest.raises(_lib.GsrError, match="not a PLY"):
        io.load_ply(p)


def test_isotropic_model_writes_scale_0_only(tmp_path):
    """export_ply iterates axes(scales,1) (gaussians.jl:176): an isotropic (1,N) model gets `scale_0` only; the
    3-row writer must not be handed a 1-row array (it would read past the buffer)."""
    io = _io()
    m = model(300, 3, 5)
    m["scales"] = m["scales"][:, :1].copy()
    p = str(tmp_path / "iso.ply")
    io.save_ply(p, **m)
    raw = open(p, "rb").read()
    head, data = raw.split(b"end_header\n", 1)
    props = [l.split()[2].decode() for l in head.split(b"\n") if l.startswith(b"property")]
    assert [q for q in props if q.startswith("scale_")] == ["scale_0"]
    a = np.frombuffer(data, "<f4").reshape(300, len(props))
    assert np.array_equal(a[:, props.index("scale_0")], m["scales"][:, 0])
    assert np.array_equal(a[:, props.index("rot_3")], m["rotations"][:, 3])
    with pytest.raises(ValueError):
        io.save_ply(p, **{**m, "scales": np.zeros((300, 2), np.float32)})
