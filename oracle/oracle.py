"""ctypes front-end of the CPU oracle (oracle/gsr_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of gsr_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package.

`Oracle.forward` / `Oracle.backward` follow the orchestration of the reference's
`rasterize` (src/rasterization/rasterizer.jl:255-408) and `∇rasterize` (:416-550)
stage by stage, calling one C function per reference kernel.

Array conventions (numpy, C-contiguous == Julia column-major with reversed dims):
  means (N,3), scales (N,3), rotations (N,4) wxyz, opacities (N,), shs (N,K,3),
  image (H,W,C), vpixels (H,W,C); ids in `values_sorted` are 1-based like the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK = 16
MODES = {"rgb": 3, "rgbd": 5, "rgbdn": 8}


def build(force: bool = False) -> None:
    """Compile liboracle_f32.so / liboracle_f64.so next to this file (gcc, seconds)."""
    targets = [os.path.join(_HERE, f) for f in ("liboracle_f32.so", "liboracle_f64.so")]
    src = os.path.join(_HERE, "gsr_oracle.c")
    if not force and all(os.path.exists(t) and os.path.getmtime(t) >= os.path.getmtime(src) for t in targets):
        return
    subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                   stdout=subprocess.DEVNULL if not force else None)


@dataclass
class OracleCamera:
    """Fields the hot path reads from `Camera` (src/camera.jl:2-45)."""
    R: np.ndarray            # (3,3) w2c rotation, R[i,j] row i col j
    t: np.ndarray            # (3,)
    focal: np.ndarray        # (2,)
    principal: np.ndarray    # (2,) in [0,1]
    cam_center: np.ndarray   # (3,)  = c2w[1:3,4]
    width: int
    height: int

    @staticmethod
    def simple(fx, fy, width, height):
        """`Camera(; fx, fy, width, height)` — src/camera.jl:37-45 (identity pose, principal 0.5)."""
        return OracleCamera(np.eye(3), np.zeros(3), np.array([fx, fy], dtype=np.float64),
                            np.array([0.5, 0.5]), np.zeros(3), int(width), int(height))


@dataclass
class OracleState:
    """GeometryState / BinningState / ImageState of the reference (src/rasterization/states.jl)."""
    n: int = 0
    n_rendered: int = 0
    depths: np.ndarray = None
    means2d: np.ndarray = None
    grad_means2d: np.ndarray = None
    rgbs: np.ndarray = None
    clamped: np.ndarray = None
    tiles_touched: np.ndarray = None
    points_offset: np.ndarray = None
    conics: np.ndarray = None
    radii: np.ndarray = None
    features: np.ndarray = None
    normals: np.ndarray = None
    keys_unsorted: np.ndarray = None
    values_unsorted: np.ndarray = None
    keys_sorted: np.ndarray = None
    values_sorted: np.ndarray = None
    ranges: np.ndarray = None
    n_contrib: np.ndarray = None
    accum_alpha: np.ndarray = None
    counts_fwd: np.ndarray = field(default_factory=lambda: np.zeros(2, np.int64))
    counts_bwd: np.ndarray = field(default_factory=lambda: np.zeros(2, np.int64))
    ambiguous: np.ndarray = None      # (H,W)  pixels with a pair near a branch threshold
    ambiguous_g: np.ndarray = None    # (N,)   Gaussians of such pairs
    cond: np.ndarray = None           # (H,W)  sum alpha/(1-alpha) over blended pairs (conditioning of T)
    image: np.ndarray = None


class Oracle:
    def __init__(self, dtype=np.float32):
        build()
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        name = "liboracle_f32.so" if self.dtype == np.float32 else "liboracle_f64.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        self.creal = C.c_float if self.dtype == np.float32 else C.c_double
        assert self.lib.orc_real_size() == self.dtype.itemsize
        r = self.creal

        class Cam(C.Structure):
            _fields_ = [("R", r * 9), ("t", r * 3), ("focal", r * 2), ("principal", r * 2),
                        ("cam_center", r * 3), ("width", C.c_int32), ("height", C.c_int32)]

        class Cfg(C.Structure):
            _fields_ = [("near_plane", r), ("far_plane", r), ("blur_eps", r), ("radius_clip", C.c_int32)]

        self.Cam, self.Cfg = Cam, Cfg
        self.lib.orc_add_blur.restype = r
        self.lib.orc_inverse.restype = r
        self.lib.orc_cumsum.restype = C.c_int64
        self.lib.orc_gaussian_normal.restype = C.c_int32

    def set_threads(self, n: int = 0) -> int:
        """OpenMP team size of the C stages (n <= 0: query).  Returns the size in effect."""
        return int(self.lib.orc_set_threads(int(n)))

    # ---------------------------------------------------------------- utils
    def arr(self, x, shape=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dtype))
        if shape is not None:
            a = a.reshape(shape)
        return a

    @staticmethod
    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def r(self, x):
        return self.creal(float(x))

    def mat_in(self, M):
        """(r,c) row/col-indexed numpy matrix -> column-major flat array."""
        return self.arr(np.asarray(M, dtype=self.dtype).T.copy().reshape(-1))

    def mat_out(self, flat, rows, cols):
        return np.asarray(flat).reshape(cols, rows).T.copy()

    def make_cam(self, cam: OracleCamera):
        c = self.Cam()
        c.R[:] = [float(v) for v in self.mat_in(cam.R)]
        c.t[:] = [float(v) for v in cam.t]
        c.focal[:] = [float(v) for v in cam.focal]
        c.principal[:] = [float(v) for v in cam.principal]
        c.cam_center[:] = [float(v) for v in cam.cam_center]
        c.width, c.height = int(cam.width), int(cam.height)
        return c

    def make_cfg(self, near=0.2, far=1000.0, blur_eps=0.3, radius_clip=3):
        g = self.Cfg()
        g.near_plane, g.far_plane, g.blur_eps, g.radius_clip = (
            float(np.float32(near)), float(np.float32(far)), float(np.float32(blur_eps)), int(radius_clip))
        return g

    # ------------------------------------------------------- scalar helpers
    def unnorm_quat2rot(self, q):
        out = np.zeros(9, self.dtype)
        self.lib.orc_unnorm_quat2rot(self.p(self.arr(q)), self.p(out))
        return self.mat_out(out, 3, 3)

    def grad_unnorm_quat2rot(self, q, vR):
        out = np.zeros(4, self.dtype)
        self.lib.orc_grad_unnorm_quat2rot(self.p(self.arr(q)), self.p(self.mat_in(vR)), self.p(out))
        return out

    def quat_scale_to_cov(self, R, s):
        out = np.zeros(9, self.dtype)
        self.lib.orc_quat_scale_to_cov(self.p(self.mat_in(R)), self.p(self.arr(s)), self.p(out))
        return self.mat_out(out, 3, 3)

    def grad_quat_scale_to_cov(self, q, s, R, vSigma, vR_extra=None):
        vq, vs = np.zeros(4, self.dtype), np.zeros(3, self.dtype)
        ex = None if vR_extra is None else self.mat_in(vR_extra)
        self.lib.orc_grad_quat_scale_to_cov(self.p(self.arr(q)), self.p(self.arr(s)), self.p(self.mat_in(R)),
                                            self.p(self.mat_in(vSigma)), self.p(ex), self.p(vq), self.p(vs))
        return vq, vs

    def pos_world_to_cam(self, R, t, p):
        out = np.zeros(3, self.dtype)
        self.lib.orc_pos_world_to_cam(self.p(self.mat_in(R)), self.p(self.arr(t)), self.p(self.arr(p)), self.p(out))
        return out

    def grad_pos_world_to_cam(self, R, t, p, v):
        vR, vt, vp = np.zeros(9, self.dtype), np.zeros(3, self.dtype), np.zeros(3, self.dtype)
        self.lib.orc_grad_pos_world_to_cam(self.p(self.mat_in(R)), self.p(self.arr(t)), self.p(self.arr(p)),
                                           self.p(self.arr(v)), self.p(vR), self.p(vt), self.p(vp))
        return self.mat_out(vR, 3, 3), vt, vp

    def covar_world_to_cam(self, R, S):
        out = np.zeros(9, self.dtype)
        self.lib.orc_covar_world_to_cam(self.p(self.mat_in(R)), self.p(self.mat_in(S)), self.p(out))
        return self.mat_out(out, 3, 3)

    def grad_covar_world_to_cam(self, R, S, vSc, vR_in):
        vR = self.mat_in(vR_in).copy()
        vS = np.zeros(9, self.dtype)
        self.lib.orc_grad_covar_world_to_cam(self.p(self.mat_in(R)), self.p(self.mat_in(S)), self.p(self.mat_in(vSc)),
                                             self.p(vR), self.p(vS))
        return self.mat_out(vR, 3, 3), self.mat_out(vS, 3, 3)

    def perspective_projection(self, mean, S, focal, res, principal):
        S2, m2 = np.zeros(4, self.dtype), np.zeros(2, self.dtype)
        res = np.ascontiguousarray(res, np.int32)
        self.lib.orc_perspective_projection(self.p(self.arr(mean)), self.p(self.mat_in(S)), self.p(self.arr(focal)),
                                            self.p(res), self.p(self.arr(principal)), self.p(S2), self.p(m2))
        return self.mat_out(S2, 2, 2), m2

    def grad_perspective_projection(self, mean, S, focal, res, principal, vS2, vm2):
        vS, vm = np.zeros(9, self.dtype), np.zeros(3, self.dtype)
        res = np.ascontiguousarray(res, np.int32)
        self.lib.orc_grad_perspective_projection(
            self.p(self.arr(mean)), self.p(self.mat_in(S)), self.p(self.arr(focal)), self.p(res),
            self.p(self.arr(principal)), self.p(self.mat_in(vS2)), self.p(self.arr(vm2)), self.p(vS), self.p(vm))
        return self.mat_out(vS, 3, 3), vm

    def add_blur(self, S2, eps):
        out, comp = np.zeros(4, self.dtype), self.creal(0)
        det = self.lib.orc_add_blur(self.p(self.mat_in(S2)), self.r(eps), self.p(out), C.byref(comp))
        return self.mat_out(out, 2, 2), det, comp.value

    def grad_add_blur(self, comp, vcomp, conic, eps):
        out = np.zeros(4, self.dtype)
        self.lib.orc_grad_add_blur(self.r(comp), self.r(vcomp), self.p(self.mat_in(conic)), self.r(eps), self.p(out))
        return self.mat_out(out, 2, 2)

    def inverse(self, X):
        out = np.zeros(4, self.dtype)
        det = self.lib.orc_inverse(self.p(self.mat_in(X)), self.p(out))
        return det, self.mat_out(out, 2, 2)

    def grad_inverse(self, Y, vY):
        out = np.zeros(4, self.dtype)
        self.lib.orc_grad_inverse(self.p(self.mat_in(Y)), self.p(self.mat_in(vY)), self.p(out))
        return self.mat_out(out, 2, 2)

    def grad_normalize(self, d, vd):
        out = np.zeros(3, self.dtype)
        self.lib.orc_grad_normalize(self.p(self.arr(d)), self.p(self.arr(vd)), self.p(out))
        return out

    def gaussian_normal(self, Rw2c, Rg, scale, mean_cam):
        n, sign = np.zeros(3, self.dtype), self.creal(0)
        k = self.lib.orc_gaussian_normal(self.p(self.mat_in(Rw2c)), self.p(self.mat_in(Rg)), self.p(self.arr(scale)),
                                         self.p(self.arr(mean_cam)), self.p(n), C.byref(sign))
        return n, int(k), sign.value

    def get_rect(self, pixel, radius, grid):
        rect = np.zeros(4, np.int32)
        self.lib.orc_get_rect(self.p(self.arr(pixel)), C.c_int32(int(radius)),
                              self.p(np.ascontiguousarray(grid, np.int32)), self.p(rect))
        return (int(rect[0]), int(rect[1])), (int(rect[2]), int(rect[3]))

    def identify_tile_range(self, keys, n_tiles):
        keys = np.ascontiguousarray(keys, np.uint64)
        ranges = np.zeros((n_tiles, 2), np.uint32)
        self.lib.orc_identify_tile_range(C.c_int64(len(keys)), self.p(keys), self.p(ranges))
        return ranges

    def colors_from_sh(self, point, cam_center, shs, degree):
        rgb, cl = np.zeros(3, self.dtype), np.zeros(3, np.uint8)
        self.lib.orc_compute_colors_from_sh(self.p(self.arr(point)), self.p(self.arr(cam_center)),
                                            self.p(self.arr(shs)), C.c_int(degree), self.p(rgb), self.p(cl))
        return rgb, cl

    def grad_color_from_sh(self, point, cam_center, shs, degree, clamped, vcolor):
        shs = self.arr(shs)
        vshs, vmean = np.zeros_like(shs), np.zeros(3, self.dtype)
        self.lib.orc_grad_color_from_sh(self.p(self.arr(point)), self.p(self.arr(cam_center)), self.p(shs),
                                        C.c_int(degree), self.p(np.ascontiguousarray(clamped, np.uint8)),
                                        self.p(self.arr(vcolor)), self.p(vshs), self.p(vmean))
        return vshs, vmean

    def sort_pairs(self, keys, values):
        keys = np.ascontiguousarray(keys, np.uint64)
        values = np.ascontiguousarray(values, np.uint32)
        ko, vo = np.zeros_like(keys), np.zeros_like(values)
        self.lib.orc_sort_pairs(C.c_int64(len(keys)), self.p(keys), self.p(values), self.p(ko), self.p(vo))
        return ko, vo

    def update_stats(self, radii, grad_means2d, width, height, max_radii, accum, denom):
        """`_update_stats!` — strategy.jl:118-136 (in place on max_radii/accum/denom)."""
        n = len(radii)
        self.lib.orc_update_stats(C.c_int64(n), self.p(np.ascontiguousarray(radii, np.int32)),
                                  self.p(self.arr(grad_means2d)), C.c_uint32(width), C.c_uint32(height),
                                  self.p(max_radii), self.p(accum), self.p(denom))

    # ------------------------------------------------- fused SSIM (§8f-2)
    def fused_ssim(self, img, ref, C1=None, C2=None, train=True):
        """`_fused_ssim` — fused_ssim.jl:354-372.  img/ref: (B,CH,H,W) C-contiguous == the reference's (W,H,CH,B).
        Returns (ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12); the last three are None when train is False."""
        img, ref = self.arr(img), self.arr(ref)
        B, CH, H, W = img.shape
        assert ref.shape == img.shape
        out = np.zeros_like(img)
        d = [np.zeros_like(img) for _ in range(3)] if train else [None, None, None]
        # Float32 keyword defaults of the reference: 0.01f0^2, 0.03f0^2 (products formed in Float32)
        c1 = float(np.float32(0.01) * np.float32(0.01)) if C1 is None else float(np.float32(C1))
        c2 = float(np.float32(0.03) * np.float32(0.03)) if C2 is None else float(np.float32(C2))
        self.lib.orc_fused_ssim(C.c_int32(W), C.c_int32(H), C.c_int32(CH), C.c_int32(B), self.p(img), self.p(ref),
                                self.r(c1), self.r(c2), C.c_int(1 if train else 0), self.p(out), self.p(d[0]),
                                self.p(d[1]), self.p(d[2]))
        return out, d[0], d[1], d[2]

    def fused_ssim_bwd(self, img, ref, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
        """`fused_ssim_bwd` — fused_ssim.jl:374-389."""
        img, ref, dL = self.arr(img), self.arr(ref), self.arr(dL_dmap)
        B, CH, H, W = img.shape
        out = np.zeros_like(img)
        self.lib.orc_fused_ssim_bwd(C.c_int32(W), C.c_int32(H), C.c_int32(CH), C.c_int32(B), self.p(img), self.p(ref),
                                    self.p(dL), self.p(self.arr(dm_dmu1)), self.p(self.arr(dm_dsigma1_sq)),
                                    self.p(self.arr(dm_dsigma12)), self.p(out))
        return out

    def photometric_loss(self, image_hwc, target, lambda_dssim=0.2):
        """The loss either side of the rasterizer (training.jl:684-694) and its pullback to the raster image:
        image_hwc (H,W,C) with rgb in channels 0..2, target (3,H,W) == the reference's (W,H,3,1).
        total = (1-l)*mean|x - t| + l*(1 - mean(fused_ssim(x; ref=t))).  Returns (total, l1, ssim_mean, vpixels (H,W,C))."""
        image_hwc = self.arr(image_hwc)
        H, W, Cc = image_hwc.shape
        x = np.ascontiguousarray(image_hwc[:, :, :3].transpose(2, 0, 1))[None]       # (1,3,H,W): permutedims + reshape
        t = self.arr(target).reshape(1, 3, H, W)
        lam = self.dtype.type(lambda_dssim)
        one = self.dtype.type(1)
        npc = x.size
        l1 = np.abs(x - t).mean(dtype=np.float64)
        m, d0, d1, d2 = self.fused_ssim(x, t, train=True)
        sm = m.mean(dtype=np.float64)
        total = float((one - lam)) * l1 + float(lam) * (1.0 - sm)
        dmap = np.full_like(x, -float(lam) / npc)
        g = self.fused_ssim_bwd(x, t, dmap, d0, d1, d2).astype(np.float64)
        g += float(one - lam) / npc * np.sign(x.astype(np.float64) - t.astype(np.float64))
        v = np.zeros((H, W, Cc), self.dtype)
        v[:, :, :3] = g[0].transpose(1, 2, 0).astype(self.dtype)
        return total, float(l1), float(sm), v

    # ---------------------------------------------------- rasterize (forward)
    def forward(self, means, shs, opacities, scales, rotations, cam: OracleCamera, *, mode="rgbd", sh_degree=0,
                background=(0.0, 0.0, 0.0), near=0.2, far=1000.0, covisibilities=None, uncertainties=None,
                state: OracleState | None = None, tile_rows=None, ambig_rel=None, ambig_cond=0.0):
        """`rasterize` — rasterizer.jl:255-408.  Returns (image (H,W,C), state).

        `state` carries stale per-Gaussian values across calls exactly like `rast.gstate`.
        `tile_rows=(y0,y1)` restricts render! to a band of tile rows (bench sampling only).
        `ambig_rel`: if set, `state.ambiguous` (H,W) flags pixels with a pair within that relative distance
        of a branch threshold (see orc_render); `ambig_cond` widens it by the conditioning of sigma."""
        lib, p = self.lib, self.p
        channels = MODES[mode]
        W, H = int(cam.width), int(cam.height)
        assert W % 16 == 0 and H % 16 == 0  # rasterizer.jl:66,281
        means = self.arr(means).reshape(-1, 3)
        n = means.shape[0]
        shs = self.arr(shs).reshape(n, -1, 3)
        K = shs.shape[1]
        opacities = self.arr(opacities).reshape(n)
        scales = self.arr(scales).reshape(n, 3)
        rotations = self.arr(rotations).reshape(n, 4)
        st = state
        if st is None or st.n < n:  # rasterizer.jl:275-278
            st = OracleState(n=n)
            z = lambda *s, dt=self.dtype: np.zeros(s, dt)
            st.depths, st.means2d, st.grad_means2d = z(n), z(n, 2), z(n, 2)
            st.rgbs, st.clamped = z(n, 3), np.zeros((n, 3), np.uint8)
            st.tiles_touched, st.points_offset = np.zeros(n, np.int32), np.zeros(n, np.int32)
            st.conics, st.radii = z(n, 3), np.zeros(n, np.int32)
            st.features = z(n, channels) if channels > 3 else None
            st.normals = z(n, 3) if channels > 5 else None
        gx, gy = -(-W // BLOCK), -(-H // BLOCK)
        grid = np.array([gx, gy], np.int32)
        if st.ranges is None:
            st.ranges = np.zeros((gx * gy, 2), np.uint32)
            st.n_contrib = np.zeros((H, W), np.uint32)
            st.accum_alpha = np.zeros((H, W), self.dtype)
        image = np.zeros((H, W, channels), self.dtype)  # fill!(rast.image, 0) rasterizer.jl:283
        ccam, ccfg = self.make_cam(cam), self.make_cfg(near, far)

        lib.orc_project(C.c_int64(n), p(means), p(scales), p(rotations), C.byref(ccam), C.byref(ccfg),
                        p(st.depths), p(st.radii), p(st.means2d), p(st.conics), p(st.normals))
        lib.orc_spherical_harmonics(C.c_int64(n), C.c_int(K), C.c_int(sh_degree), p(st.radii), p(means),
                                    p(self.arr(cam.cam_center)), p(shs), p(st.rgbs), p(st.clamped))
        lib.orc_count_tiles(C.c_int64(n), p(st.means2d), p(st.radii), p(grid), p(st.tiles_touched))
        st.n_rendered = int(lib.orc_cumsum(C.c_int64(n), p(st.tiles_touched), p(st.points_offset)))
        st.image = image
        if st.n_rendered == 0:  # rasterizer.jl:338 — zero image, NOT background
            return image, st
        m = st.n_rendered
        st.keys_unsorted, st.values_unsorted = np.zeros(m, np.uint64), np.zeros(m, np.uint32)
        st.keys_sorted, st.values_sorted = np.zeros(m, np.uint64), np.zeros(m, np.uint32)
        lib.orc_duplicate_with_keys(C.c_int64(n), p(st.means2d), p(st.depths), p(st.points_offset), p(st.radii),
                                    p(grid), p(st.keys_unsorted), p(st.values_unsorted))
        lib.orc_sort_pairs(C.c_int64(m), p(st.keys_unsorted), p(st.values_unsorted), p(st.keys_sorted),
                           p(st.values_sorted))
        st.ranges[:] = 0  # rasterizer.jl:375
        lib.orc_identify_tile_range(C.c_int64(m), p(st.keys_sorted), p(st.ranges))
        if channels > 3:
            lib.orc_pack_features(C.c_int64(n), C.c_int(channels), p(st.rgbs), p(st.depths), p(st.normals),
                                  p(st.features))
            feats = st.features
        else:
            feats = st.rgbs
        bg = np.zeros(channels, self.dtype)
        bg[:3] = np.asarray(background, self.dtype)  # feature_background rasterizer.jl:411-414
        st.counts_fwd[:] = 0
        st.ambiguous = np.zeros((H, W), np.uint8) if ambig_rel is not None else None
        st.ambiguous_g = np.zeros(n, np.uint8) if ambig_rel is not None else None
        st.cond = np.zeros((H, W), self.dtype) if ambig_rel is not None else None
        y0, y1 = tile_rows if tile_rows is not None else (0, gy)
        lib.orc_render(C.c_int(channels), C.c_int32(W), C.c_int32(H), p(st.ranges), p(st.values_sorted),
                       p(st.means2d), p(opacities), p(st.conics), p(feats), p(bg), p(image), p(st.n_contrib),
                       p(st.accum_alpha), p(covisibilities), p(uncertainties), p(st.counts_fwd),
                       C.c_int32(y0), C.c_int32(y1), p(st.ambiguous), self.r(ambig_rel or 0.0), p(st.ambiguous_g), self.r(ambig_cond), p(st.cond))
        return image, st

    # --------------------------------------------------- ∇rasterize (backward)
    def backward(self, vpixels, means, shs, opacities, scales, rotations, cam: OracleCamera, state: OracleState, *,
                 mode="rgbd", sh_degree=0, background=(0.0, 0.0, 0.0), pose_grad=False, tile_rows=None):
        """`∇rasterize` — rasterizer.jl:416-550.  Returns dict(vmeans, vshs, vopacities, vscales, vrot, vR, vt)
        and leaves `state.grad_means2d` (= rast.gstate.∇means_2d, pixel units) filled."""
        lib, p, st = self.lib, self.p, state
        channels = MODES[mode]
        W, H = int(cam.width), int(cam.height)
        means = self.arr(means).reshape(-1, 3)
        n = means.shape[0]
        shs = self.arr(shs).reshape(n, -1, 3)
        K = shs.shape[1]
        opacities = self.arr(opacities).reshape(n)
        scales = self.arr(scales).reshape(n, 3)
        rotations = self.arr(rotations).reshape(n, 4)
        vpixels = self.arr(vpixels).reshape(H, W, channels)
        f64 = lambda *s: np.zeros(s, np.float64)
        vcol, vcon, vm2, vop = f64(n, channels), f64(n, 3), f64(n, 2), f64(n)
        bg = np.zeros(channels, self.dtype)
        bg[:3] = np.asarray(background, self.dtype)
        feats = st.features if channels > 3 else st.rgbs
        gy = -(-H // BLOCK)
        y0, y1 = tile_rows if tile_rows is not None else (0, gy)
        st.counts_bwd[:] = 0
        if st.n_rendered > 0:
            lib.orc_grad_render(C.c_int(channels), C.c_int32(W), C.c_int32(H), p(st.ranges), p(st.values_sorted),
                                p(st.means2d), p(opacities), p(st.conics), p(feats), p(bg), p(vpixels),
                                p(st.n_contrib), p(st.accum_alpha), p(vcol), p(vop), p(vcon), p(vm2),
                                p(st.counts_bwd), C.c_int32(y0), C.c_int32(y1))
        st.grad_means2d[:n] = vm2.astype(self.dtype)
        vcol_r, vcon_r = vcol.astype(self.dtype), vcon.astype(self.dtype)
        vrgbs = np.ascontiguousarray(vcol_r[:, :3])
        vdepths = np.ascontiguousarray(vcol_r[:, 3]) if channels > 3 else None
        vnormals = np.ascontiguousarray(vcol_r[:, 5:8]) if channels > 5 else None
        vmeans, vscales, vrot = (np.zeros((n, 3), self.dtype), np.zeros((n, 3), self.dtype),
                                 np.zeros((n, 4), self.dtype))
        vshs = np.zeros_like(shs)
        vR = f64(9) if pose_grad else None
        vt = f64(3) if pose_grad else None
        ccam = self.make_cam(cam)
        gm2 = np.ascontiguousarray(st.grad_means2d[:n])
        lib.orc_grad_project(C.c_int64(n), p(gm2), p(vcon_r), p(vdepths), p(vnormals), p(st.conics), p(st.radii),
                             p(means), p(scales), p(rotations), C.byref(ccam), p(vmeans), p(vscales), p(vrot),
                             p(vR), p(vt))
        lib.orc_grad_spherical_harmonics(C.c_int64(n), C.c_int(K), C.c_int(sh_degree), p(means),
                                         p(self.arr(cam.cam_center)), p(shs), p(st.clamped), p(vrgbs), C.c_int(3),
                                         p(vshs), p(vmeans))
        out = dict(vmeans=vmeans, vshs=vshs, vopacities=vop.astype(self.dtype), vscales=vscales, vrot=vrot,
                   vR=None if vR is None else vR.reshape(3, 3).T.copy(), vt=vt,
                   vcolors=vcol_r, vconics=vcon_r, vmeans2d=st.grad_means2d[:n].copy())
        return out
