"""Host-side mirror of the reference's rasterizer interface, over the libgsrast C ABI.

Same names, argument meaning and error behaviour as src/rasterization/rasterizer.jl of GaussianSplatting.jl:

    GaussianRasterizer(width=, height=, mode=, near_plane=, far_plane=)         rasterizer.jl:60-90
    rast(means_3d, opacities, scales, rotations, sh_color, sh_remainder,
         R_w2c=None, t_w2c=None, camera=, sh_degree=, background=, ...)         rasterizer.jl:200-253
    rasterize(means_3d, shs, opacities, scales, rotations, R_w2c, t_w2c, rast=, camera=, ...)   :255-408
    its rrule / ∇rasterize                                                      :416-573  (torch.autograd.Function)
    rast.image, rast.gstate.radii, rast.gstate.grad_means2d (∇means_2d)         strategy.jl:85-86
    update_stats(max_radii, accum, denom, rast)                                 strategy.jl:107-136

PyTorch supplies device memory, streams and autograd plumbing only; every kernel on the path is in
libgsrast.so.  Tensor shapes are the reference's column-major arrays read row-major:
(3,N)->(N,3), (4,N)->(N,4), (1,N)->(N,1), (3,K,N)->(N,K,3), image (C,W,H)->(H,W,C).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import GsrCamera, GsrConfig, GsrStateViews, check

MODES = {"rgb": 3, "rgbd": 5, "rgbdn": 8}  # n_color_features, rasterizer.jl:47-51
BLOCK = 16


@dataclass
class Camera:
    """The fields of `Camera` the rasterizer reads (src/camera.jl:2-45)."""
    fx: float
    fy: float
    width: int
    height: int
    R: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=np.float32))  # w2c rotation
    t: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))      # w2c translation
    principal: tuple = (0.5, 0.5)                                               # in [0,1]

    @property
    def camera_center(self) -> np.ndarray:  # c2w[1:3,4], camera.jl:29
        R = np.asarray(self.R, np.float64)
        return (-R.T @ np.asarray(self.t, np.float64)).astype(np.float32)

    def to_c(self, R_dev=None, t_dev=None) -> GsrCamera:
        c = GsrCamera()
        R = np.asarray(self.R, np.float32)
        c.R[:] = [float(R[i, j]) for j in range(3) for i in range(3)]  # column-major
        c.t[:] = [float(v) for v in np.asarray(self.t, np.float32)]
        c.focal[:] = [float(np.float32(self.fx)), float(np.float32(self.fy))]
        c.principal[:] = [float(np.float32(self.principal[0])), float(np.float32(self.principal[1]))]
        c.cam_center[:] = [float(v) for v in self.camera_center]
        c.R_dev = R_dev
        c.t_dev = t_dev
        return c


class _DevView:
    """Zero-copy torch view of a device pointer owned by the library (via __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr, shape, typestr, device):
    if not ptr or any(s == 0 for s in shape):
        dt = {"<f4": torch.float32, "<i4": torch.int32, "|u1": torch.uint8, "<i8": torch.int64}[typestr]
        return torch.empty(tuple(shape), dtype=dt, device=device)
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name} must be a CUDA tensor: the rasterizer has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32")
    return t.contiguous()


class GeometryStateView:
    """`rast.gstate` (states.jl:2-47) as torch views into the handle's workspace."""

    def __init__(self, rast: "GaussianRasterizer"):
        self._rast = rast

    def _v(self):
        v = GsrStateViews()
        check(_lib.lib().gsr_get_state(self._rast._h, C.byref(v)), self._rast._h)
        return v

    def __getattr__(self, name):
        v, dev = self._v(), self._rast.device
        n, m = int(v.n), int(v.n_rendered)
        W, H, T = self._rast.width, self._rast.height, self._rast.n_tiles
        table = {
            "radii": (v.radii, (n,), "<i4"), "grad_means2d": (v.grad_means2d, (n, 2), "<f4"),
            "means2d": (v.means2d, (n, 2), "<f4"), "depths": (v.depths, (n,), "<f4"),
            "conics": (v.conics, (n, 3), "<f4"), "rgbs": (v.rgbs, (n, 3), "<f4"),
            "clamped": (v.clamped, (n, 3), "|u1"), "tiles_touched": (v.tiles_touched, (n,), "<i4"),
            "points_offset": (v.points_offset, (n,), "<i4"), "normals": (v.normals, (n, 3), "<f4"),
            "keys_unsorted": (v.keys_unsorted, (m,), "<i8"), "values_unsorted": (v.values_unsorted, (m,), "<i4"),
            "keys_sorted": (v.keys_sorted, (m,), "<i8"), "values_sorted": (v.values_sorted, (m,), "<i4"),
            "ranges": (v.ranges, (T, 2), "<i4"), "n_contrib": (v.n_contrib, (H, W), "<i4"),
            "accum_alpha": (v.accum_alpha, (H, W), "<f4"),
        }
        if name == "n_rendered":
            return m
        if name not in table:
            raise AttributeError(name)
        return _view(*table[name], dev)


class GaussianRasterizer:
    """`GaussianRasterizer(kab; width, height, mode, near_plane, far_plane)` — rasterizer.jl:60-90."""

    def __init__(self, *, width: int, height: int, mode: str = "rgbd", near_plane: float = 0.2,
                 far_plane: float = 1000.0, device="cuda", math_mode: str | int = "strict"):
        assert width % 16 == 0 and height % 16 == 0  # rasterizer.jl:66
        if mode not in MODES:
            raise ValueError(f"Invalid render: {mode} ∉ {tuple(MODES)}")  # rasterizer.jl:68
        if not torch.cuda.is_available():
            raise RuntimeError("GaussianRasterizer needs a CUDA device (sm_100a); there is no CPU fallback")
        self.width, self.height, self.mode = int(width), int(height), mode
        self.channels = MODES[mode]
        self.near_plane, self.far_plane = float(near_plane), float(far_plane)
        self.device = torch.device(device)
        self.grid = (width // BLOCK, height // BLOCK)
        self.n_tiles = self.grid[0] * self.grid[1]
        # "strict" (default): meets the flat 1e-5 / 1e-4 tolerances; "reference": every op in the reference's order;
        # "fast": log2-domain ex2.approx path with conditioning-dependent error; an int selects an A/B policy build
        self.math_mode = math_mode
        mm = math_mode if isinstance(math_mode, int) else {"reference": _lib.MATH_REFERENCE, "fast": _lib.MATH_FAST,
                                                            "strict": _lib.MATH_STRICT}[math_mode]
        cfg = GsrConfig(width, height, self.channels, near_plane, far_plane, 3, 0.3, mm)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_create(C.byref(cfg), C.byref(self._h)))
        self.image = torch.zeros((height, width, self.channels), dtype=torch.float32, device=self.device)
        self.gstate = GeometryStateView(self)
        self.n_rendered = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().gsr_destroy(h)
            except Exception:
                pass
            self._h = None

    def release_scene_buffers(self):  # rasterizer.jl:111-123
        check(_lib.lib().gsr_release_scene_buffers(self._h), self._h)

    def memory_usage(self) -> int:  # rasterizer.jl:127-134
        b = C.c_size_t()
        check(_lib.lib().gsr_memory_usage(self._h, C.byref(b)), self._h)
        return int(b.value) + self.image.numel() * 4

    # ---- functor: activation pre-pass + rasterize (rasterizer.jl:200-253) ---------------------------------
    def __call__(self, means_3d, opacities, scales, rotations, sh_color, sh_remainder, R_w2c=None, t_w2c=None, *,
                 camera: Camera, sh_degree: int, background=(0.0, 0.0, 0.0), covisibilities=None,
                 uncertainties=None, fused_activations: bool = True):
        """Raw parameters in (pre-sigmoid opacities, log-scales (N,3) or isotropic (N,1), features_dc | features_rest).

        fused_activations=True (default): one library call each way — sigmoid / exp / the SH concatenation and their
        pullbacks run inside the per-Gaussian kernels (gsr_forward_raw / gsr_backward_raw).  False: the reference's
        own composition — torch broadcasts, then `rasterize` on the activated arrays (autograd chains the pullbacks)."""
        if fused_activations:
            rest = None if sh_remainder is None or sh_remainder.numel() == 0 else sh_remainder
            return _RasterizeRaw.apply(means_3d, opacities, scales, rotations, sh_color, rest, R_w2c, t_w2c, self, camera,
                                       int(sh_degree), tuple(float(b) for b in background), covisibilities, uncertainties)
        shs = sh_color if sh_remainder is None or sh_remainder.numel() == 0 else torch.cat([sh_color, sh_remainder], 1)
        opacities_act = torch.sigmoid(opacities)
        if scales.shape[1] == 1:  # isotropic (rasterizer.jl:235-244)
            scales = scales.expand(-1, 3)
        scales_act = torch.exp(scales)
        return rasterize(means_3d, shs, opacities_act, scales_act, rotations, R_w2c, t_w2c, rast=self, camera=camera,
                         sh_degree=sh_degree, background=background, covisibilities=covisibilities,
                         uncertainties=uncertainties)

    def _raw_call(self, backward, means, opac, scales, rots, dc, rest, R_w2c, t_w2c, camera, sh_degree, background,
                  image=None, covis=None, uncert=None, vpixels=None, outs=None, accumulate=False):
        assert camera.width == self.width and camera.height == self.height
        n = means.shape[0]
        K = 1 + (0 if rest is None else rest.shape[1])
        iso = int(scales.shape[1] == 1)
        cam = camera.to_c(_ptr(R_w2c), _ptr(t_w2c))
        bg = (C.c_float * 3)(*[float(b) for b in background])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            if not backward:
                m = C.c_int64(0)
                check(_lib.lib().gsr_forward_raw(self._h, C.byref(cam), n, sh_degree, K, _ptr(means), _ptr(dc), _ptr(rest),
                                                 _ptr(opac), _ptr(scales), iso, _ptr(rots), bg, _ptr(image), _ptr(covis),
                                                 _ptr(uncert), C.byref(m), stream), self._h)
                self.n_rendered = int(m.value)
                return image
            dev = self.device
            if outs is None:
                outs = dict(vmeans=torch.empty((n, 3), device=dev), vfeatures_dc=torch.empty((n, 1, 3), device=dev),
                            vfeatures_rest=None if rest is None else torch.empty((n, K - 1, 3), device=dev),
                            vopacities=torch.empty((n, 1), device=dev), vscales=torch.empty_like(scales),
                            vrot=torch.empty((n, 4), device=dev))
            vR = vt = None
            if R_w2c is not None:
                vR, vt = torch.zeros((3, 3), device=dev), torch.zeros(3, device=dev)
            check(_lib.lib().gsr_backward_raw(self._h, C.byref(cam), n, sh_degree, K, _ptr(means), _ptr(dc), _ptr(rest),
                                              _ptr(opac), _ptr(scales), iso, _ptr(rots), bg, _ptr(vpixels),
                                              _ptr(outs["vmeans"]), _ptr(outs["vfeatures_dc"]), _ptr(outs["vfeatures_rest"]),
                                              _ptr(outs["vopacities"]), _ptr(outs["vscales"]), _ptr(outs["vrot"]), _ptr(vR),
                                              _ptr(vt), int(bool(accumulate)), stream), self._h)
            outs["vR"] = None if vR is None else vR.t()
            outs["vt"] = vt
            return outs

    # ---- raw stages ------------------------------------------------------------------------------------------
    def _forward(self, means, shs, opac, scales, rots, R_w2c, t_w2c, camera, sh_degree, background, covis, uncert,
                 out=None):
        assert camera.width == self.width and camera.height == self.height
        n, K = means.shape[0], shs.shape[1]
        out = self.image if out is None else out
        cam = camera.to_c(_ptr(R_w2c), _ptr(t_w2c))
        bg = (C.c_float * 3)(*[float(b) for b in background])
        m = C.c_int64(0)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_forward(self._h, C.byref(cam), n, sh_degree, K, _ptr(means), _ptr(shs), _ptr(opac),
                                         _ptr(scales), _ptr(rots), bg, _ptr(out), _ptr(covis), _ptr(uncert),
                                         C.byref(m), stream), self._h)
        self.n_rendered = int(m.value)
        return out

    def _backward(self, vpixels, means, shs, opac, scales, rots, R_w2c, t_w2c, camera, sh_degree, background,
                  outs=None, accumulate=False):
        n, K = means.shape[0], shs.shape[1]
        dev = self.device
        if outs is None:
            outs = dict(vmeans=torch.empty((n, 3), device=dev), vshs=torch.empty((n, K, 3), device=dev),
                        vopacities=torch.empty((n, 1), device=dev), vscales=torch.empty((n, 3), device=dev),
                        vrot=torch.empty((n, 4), device=dev))
        vR = vt = None
        if R_w2c is not None:  # rasterizer.jl:495-503
            vR, vt = torch.zeros((3, 3), device=dev), torch.zeros(3, device=dev)
        cam = camera.to_c(_ptr(R_w2c), _ptr(t_w2c))
        bg = (C.c_float * 3)(*[float(b) for b in background])
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            check(_lib.lib().gsr_backward(self._h, C.byref(cam), n, sh_degree, K, _ptr(means), _ptr(shs), _ptr(opac),
                                          _ptr(scales), _ptr(rots), bg, _ptr(vpixels), _ptr(outs["vmeans"]),
                                          _ptr(outs["vshs"]), _ptr(outs["vopacities"]), _ptr(outs["vscales"]),
                                          _ptr(outs["vrot"]), _ptr(vR), _ptr(vt), int(bool(accumulate)), stream),
                  self._h)
        outs["vR"] = None if vR is None else vR.t()  # library writes column-major (3,3)
        outs["vt"] = vt
        return outs


    def forward_backward_host(self, host: dict, vpixels_h, camera: Camera, sh_degree: int, background=(0.0, 0.0, 0.0),
                              out: dict | None = None, wait: bool = True):
        """gsr_forward_backward_host: inputs / outputs are HOST (ideally pinned) float32 torch tensors or NumPy
        arrays; H2D of the five parameter arrays + vpixels, forward, backward, D2H of image + gradients.
        wait=False submits asynchronously (gsr_forward_backward_host_async): consecutive submissions overlap their
        transfers; call `host_wait()` before reading `out`."""
        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        n, K = host["means"].shape[0], host["shs"].shape[1]
        out = out or {}
        cam = camera.to_c()
        bg = (C.c_float * 3)(*[float(b) for b in background])
        m = C.c_int64(0)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            fn = _lib.lib().gsr_forward_backward_host if wait else _lib.lib().gsr_forward_backward_host_async
            check(fn(
                self._h, C.byref(cam), n, sh_degree, K, hp(host["means"]), hp(host["shs"]), hp(host["opac"]),
                hp(host["scales"]), hp(host["rots"]), bg, hp(vpixels_h), hp(out.get("image")), hp(out.get("vmeans")),
                hp(out.get("vshs")), hp(out.get("vopacities")), hp(out.get("vscales")), hp(out.get("vrot")),
                C.byref(m), stream), self._h)
        self.n_rendered = int(m.value)
        return out

    # ---- multi-GPU peer-fused backward (gsrast.distributed.PeerFusedBackward drives these) ------------------
    def set_accumulator(self, gacc: torch.Tensor | None):
        """Keep the per-Gaussian accumulator in caller-provided (symmetric / peer-mapped) memory."""
        if gacc is None:
            check(_lib.lib().gsr_set_accumulator(self._h, None, 0), self._h)
        else:
            af = 16 if self.channels <= 6 else 20  # csrc/common.cuh acc_floats
            check(_lib.lib().gsr_set_accumulator(self._h, _ptr(gacc), gacc.numel() // af), self._h)
        self._ext_gacc = gacc

    def backward_render(self, vpixels, n, background=(0.0, 0.0, 0.0)):
        bg = (C.c_float * 3)(*[float(b) for b in background])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_backward_render(self._h, n, bg, _ptr(vpixels), stream), self._h)

    def backward_gaussians_peers(self, world, rank, cameras, gacc_ptrs, table_ptrs, means, shs, opac, scales, rots,
                                 sh_degree):
        n, K = means.shape[0], shs.shape[1]
        cams = (GsrCamera * world)(*[c.to_c() for c in cameras])
        ga = (C.c_void_p * world)(*[int(p) for p in gacc_ptrs])
        tb = (C.c_void_p * world)(*[int(p) for p in table_ptrs])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_backward_gaussians_peers(self._h, world, rank, cams, ga, tb, n, sh_degree, K, _ptr(means),
                                                          _ptr(shs), _ptr(opac), _ptr(scales), _ptr(rots), stream), self._h)

    def export_accumulator(self, n: int, rows: torch.Tensor):
        """gsr_export_accumulator: the accumulator of the last backward_render as exchange rows ([n][12 or 16] floats)."""
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_export_accumulator(self._h, n, _ptr(rows), stream), self._h)

    def backward_gaussians_views(self, cameras, view_gacc_ptrs, world, rank, table_ptrs, means, shs, opac, scales, rots,
                                 sh_degree, exchange_rows: bool = False):
        """gsr_backward_gaussians_views: per-Gaussian backward of a whole view batch (one accumulator per view, wherever
        it was rendered), reduced over the views and stored into every rank's table."""
        n, K = means.shape[0], shs.shape[1]
        nv = len(cameras)
        cams = (GsrCamera * nv)(*[c.to_c() for c in cameras])
        ga = (C.c_void_p * nv)(*[int(p) for p in view_gacc_ptrs])
        tb = (C.c_void_p * world)(*[(int(p) or None) for p in table_ptrs])  # 0 -> NULL: that rank receives nothing
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(_lib.lib().gsr_backward_gaussians_views(self._h, nv, cams, ga, int(bool(exchange_rows)), world, rank, tb, n,
                                                          sh_degree, K,
                                                          _ptr(means), _ptr(shs), _ptr(opac), _ptr(scales), _ptr(rots),
                                                          stream), self._h)

    def forward_generation(self) -> int:
        """Number of forwards this handle has run (gsr_forward_generation)."""
        return int(_lib.lib().gsr_forward_generation(self._h))

    def _check_generation(self, gen: int):
        """The handle keeps the per-pixel / binning state of its LAST forward only (as rast.{g,b,i}state do in the
        reference): a backward whose forward has since been overwritten would silently differentiate the wrong view."""
        now = self.forward_generation()
        if now != gen:
            raise RuntimeError(f"rasterize backward: the rasterizer ran {now - gen} more forward(s) after the one being "
                               "differentiated; its state was overwritten. Call backward before the next forward on "
                               "this GaussianRasterizer, or use one rasterizer per in-flight view.")

    def host_wait(self):
        check(_lib.lib().gsr_host_wait(self._h), self._h)

    def profile(self, enable: bool):
        check(_lib.lib().gsr_profile_enable(self._h, int(enable)), self._h)

    def stage_times_ms(self) -> dict:
        arr = (C.c_float * len(_lib.STAGES))()
        check(_lib.lib().gsr_profile_get(self._h, arr), self._h)
        return {k: float(arr[i]) for i, k in enumerate(_lib.STAGES)}


class _Rasterize(torch.autograd.Function):
    """`ChainRulesCore.rrule(::typeof(rasterize), ...)` — rasterizer.jl:552-573."""

    @staticmethod
    def forward(ctx, means, shs, opac, scales, rots, R_w2c, t_w2c, rast, camera, sh_degree, background, covis, uncert):
        means, shs, opac = _f32c(means, "means_3d"), _f32c(shs, "shs"), _f32c(opac, "opacities")
        scales, rots = _f32c(scales, "scales"), _f32c(rots, "rotations")
        # `R_w2c` is (3,3) with R[i,j] = row i, col j on the torch side; the library reads column-major
        Rc = None if R_w2c is None else _f32c(R_w2c, "R_w2c").t().contiguous()
        tc = None if t_w2c is None else _f32c(t_w2c, "t_w2c")
        # The reference returns the rasterizer-owned `rast.image`, overwritten by the next call
        # (rasterizer.jl:152-153).  Autograd must not see its output mutated, so every differentiable call
        # renders into a fresh tensor, which also becomes `rast.image`.
        out = torch.empty((rast.height, rast.width, rast.channels), dtype=torch.float32, device=rast.device)
        image = rast._forward(means, shs, opac, scales, rots, Rc, tc, camera, sh_degree, background, covis, uncert,
                              out=out)
        rast.image = image
        ctx.save_for_backward(means, shs, opac, scales, rots, Rc, tc)
        ctx.rast, ctx.camera, ctx.sh_degree, ctx.background = rast, camera, sh_degree, background
        ctx.generation = rast.forward_generation()
        return image

    @staticmethod
    def backward(ctx, vpixels):
        means, shs, opac, scales, rots, Rc, tc = ctx.saved_tensors
        ctx.rast._check_generation(ctx.generation)
        g = ctx.rast._backward(vpixels.contiguous(), means, shs, opac, scales, rots, Rc, tc, ctx.camera, ctx.sh_degree,
                               ctx.background)
        return (g["vmeans"], g["vshs"], g["vopacities"].view_as(opac), g["vscales"], g["vrot"], g["vR"], g["vt"],
                None, None, None, None, None, None)


class _RasterizeRaw(torch.autograd.Function):
    """The functor of rasterizer.jl:200-253 as ONE differentiable operator on the raw parameters."""

    @staticmethod
    def forward(ctx, means, opac, scales, rots, dc, rest, R_w2c, t_w2c, rast, camera, sh_degree, background, covis,
                uncert):
        means, opac, scales = _f32c(means, "means_3d"), _f32c(opac, "opacities"), _f32c(scales, "scales")
        rots, dc = _f32c(rots, "rotations"), _f32c(dc, "sh_color")
        rest = None if rest is None else _f32c(rest, "sh_remainder")
        Rc = None if R_w2c is None else _f32c(R_w2c, "R_w2c").t().contiguous()
        tc = None if t_w2c is None else _f32c(t_w2c, "t_w2c")
        out = torch.empty((rast.height, rast.width, rast.channels), dtype=torch.float32, device=rast.device)
        image = rast._raw_call(False, means, opac, scales, rots, dc, rest, Rc, tc, camera, sh_degree, background,
                               image=out, covis=covis, uncert=uncert)
        rast.image = image
        ctx.save_for_backward(means, opac, scales, rots, dc, rest, Rc, tc)
        ctx.rast, ctx.camera, ctx.sh_degree, ctx.background = rast, camera, sh_degree, background
        ctx.generation = rast.forward_generation()
        return image

    @staticmethod
    def backward(ctx, vpixels):
        means, opac, scales, rots, dc, rest, Rc, tc = ctx.saved_tensors
        ctx.rast._check_generation(ctx.generation)
        g = ctx.rast._raw_call(True, means, opac, scales, rots, dc, rest, Rc, tc, ctx.camera, ctx.sh_degree,
                               ctx.background, vpixels=vpixels.contiguous())
        return (g["vmeans"], g["vopacities"].view_as(opac), g["vscales"], g["vrot"], g["vfeatures_dc"],
                g["vfeatures_rest"], g["vR"], g["vt"], None, None, None, None, None, None)


def rasterize(means_3d, shs, opacities, scales, rotations, R_w2c=None, t_w2c=None, *, rast: GaussianRasterizer,
              camera: Camera, sh_degree: int, background=(0.0, 0.0, 0.0), covisibilities=None, uncertainties=None):
    """`rasterize(...)` — rasterizer.jl:255-408.  Differentiable w.r.t. the five parameter arrays (+ R_w2c, t_w2c)."""
    return _Rasterize.apply(means_3d, shs, opacities, scales, rotations, R_w2c, t_w2c, rast, camera, int(sh_degree),
                            tuple(float(b) for b in background), covisibilities, uncertainties)


def update_stats(max_radii: torch.Tensor, accum_grad_means2d: torch.Tensor, denom: torch.Tensor,
                 rast: GaussianRasterizer):
    """`update_stats!(strategy, rast.gstate.radii, rast.gstate.∇means_2d, resolution)` — strategy.jl:107-136."""
    assert max_radii.dtype == torch.int32 and accum_grad_means2d.dtype == torch.float32 and denom.dtype == torch.float32
    n = max_radii.numel()
    stream = C.c_void_p(torch.cuda.current_stream(rast.device).cuda_stream)
    check(_lib.lib().gsr_update_stats(rast._h, n, _ptr(max_radii), _ptr(accum_grad_means2d), _ptr(denom), stream),
          rast._h)
