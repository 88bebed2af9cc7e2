"""the reference's checkpoint container (src/checkpoint.jl — safetensors with dotted names) read and
written by gsrast.io, checked against the independent `safetensors` package and the facts the reference's own test
pins (test/runtests.jl:905-980): shapes as Julia sees them, the format marker, junk files rejected."""
import numpy as np
import pytest


def _model(n, K, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
    return dict(points=f(n, 3), features_dc=f(n, 1, 3), features_rest=np.arange(n * (K - 1) * 3, dtype=np.float32).reshape(n, K - 1, 3),
                scales=f(n, 3), rotations=f(n, 4), opacities=f(n, 1), sh_degree=2, max_sh_degree=3)


def test_round_trip_and_julia_shapes(tmp_path):
    from gsrast import io
    m = _model(11, 16)
    p = str(tmp_path / "state.safetensors")
    grids = np.random.default_rng(1).random((5, 4, 8, 8, 12), dtype=np.float32)
    io.save_checkpoint(p, m, meta={"step": 30000}, extra={"bilateral.grids": grids})
    back = io.load_checkpoint(p)
    for k in ("points", "features_dc", "features_rest", "scales", "rotations", "opacities"):
        assert np.array_equal(back[k], m[k]), k
    assert back["sh_degree"] == 2 and back["max_sh_degree"] == 3 and back["meta"]["step"] == "30000"
    # the stored shapes are the ones Julia sees (runtests.jl:972-974); read with the independent implementation
    from safetensors.numpy import load_file
    st = load_file(p)
    assert st["gaussians.points"].shape == (3, 11) and st["gaussians.features_rest"].shape == (3, 15, 11)
    assert np.array_equal(st["gaussians.points"], m["points"].T)
    # Julia (3,15,n) column-major element (c,k,i)  ==  our (n,15,3)[i,k,c]
    assert st["gaussians.features_rest"][2, 7, 4] == m["features_rest"][4, 7, 2]
    assert np.array_equal(st["bilateral.grids"], grids)
    tensors, meta = io.read_safetensors(p)
    assert meta["format"] == io.CHECKPOINT_FORMAT and "sky.gaussians.points" not in tensors


def test_reads_files_written_by_the_safetensors_package(tmp_path):
    from gsrast import io
    from safetensors.numpy import save_file
    m = _model(5, 4)
    p = str(tmp_path / "ext.safetensors")
    save_file({f"gaussians.{k}": np.ascontiguousarray(m[k].transpose(*reversed(range(m[k].ndim))))
               for k in ("points", "features_dc", "features_rest", "scales", "rotations", "opacities")}, p,
              metadata={"format": io.CHECKPOINT_FORMAT, "gaussians.sh_degree": "1", "gaussians.max_sh_degree": "1"})
    back = io.load_checkpoint(p)
    assert np.array_equal(back["features_rest"], m["features_rest"]) and back["sh_degree"] == 1


def test_rejects_foreign_and_junk_files(tmp_path):  # checkpoint.jl:65-68, runtests.jl:976-978
    from gsrast import io
    junk = str(tmp_path / "junk.safetensors")
    open(junk, "wb").write(np.random.default_rng(0).integers(0, 256, 64, dtype=np.uint8).tobytes())
    with pytest.raises(ValueError):
        io.load_checkpoint(junk)
    foreign = str(tmp_path / "foreign.safetensors")
    io.write_safetensors(foreign, {"x": np.zeros((2, 2), np.float32)}, {"format": "something else"})
    with pytest.raises(ValueError, match="not a GaussianSplatting.jl checkpoint"):
        io.load_checkpoint(foreign)


def test_checkpoint_to_ply_to_rasterizer_arrays(tmp_path):
    """file -> file: a checkpoint's model exported as .ply and read back gives the same raw arrays."""
    from gsrast import io
    m = _model(9, 16)
    c, p = str(tmp_path / "a.safetensors"), str(tmp_path / "a.ply")
    io.save_checkpoint(c, m)
    g = io.load_checkpoint(c)
    io.save_ply(p, *(g[k] for k in ("points", "features_dc", "features_rest", "opacities", "scales", "rotations")))
    back = io.load_ply(p)
    for k in ("points", "features_dc", "features_rest", "scales", "rotations", "opacities"):
        assert np.array_equal(back[k], m[k]), k
