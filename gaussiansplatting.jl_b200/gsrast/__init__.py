"""gsrast — B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.

Host-side mirror of GaussianSplatting.jl's `GaussianRasterizer` / `rasterize` / rrule interface
(src/rasterization/rasterizer.jl) over the C ABI of libgsrast.so (include/gsrast.h).
`gsrast.synthetic` (NumPy only) can be imported without CUDA; everything else needs the built library.
"""
from . import synthetic  # noqa: F401

__all__ = ["Camera", "GaussianRasterizer", "rasterize", "update_stats", "synthetic"]


def __getattr__(name):  # lazy: keeps `import gsrast.synthetic` free of torch / the shared library
    if name in ("Camera", "GaussianRasterizer", "rasterize", "update_stats", "MODES"):
        from . import rasterizer
        return getattr(rasterizer, name)
    raise AttributeError(name)
