"""Synthetic Gaussian scenes and cameras of the shapes BASELINE.json names.

Pure NumPy (PCG64 via `default_rng(seed)`), float32, generated on the host.  Shared by the
parity tests, the CPU oracle baseline and bench.py so every arm sees the same bytes.
Array conventions: (N,3)/(N,4)/(N,K,3) C-contiguous == the reference's column-major
(3,N)/(4,N)/(3,K,N) (src/gaussians.jl:2-17).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# name -> (N, sh_degree, width, height, mode, seed, n_views)   (resolutions ×16-rounded, dataset.jl:92-94)
CONFIGS = {
    "C1": (10_000, 0, 256, 256, "rgb", 1001, 1),
    "C1d": (10_000, 0, 256, 256, "rgbd", 1001, 1),
    "C2": (1_000_000, 3, 1920, 1088, "rgbd", 1002, 1),
    "C3": (3_000_000, 3, 1920, 1088, "rgbd", 1003, 8),
    "C4": (6_000_000, 3, 3840, 2160, "rgbdn", 1004, 1),
    "C5": (500_000, 3, 1312, 848, "rgbd", 1005, 1),
}


@dataclass
class Scene:
    means: np.ndarray        # (N,3)
    scales: np.ndarray       # (N,3) activated (exp already applied)
    rotations: np.ndarray    # (N,4) wxyz, un-normalised
    opacities: np.ndarray    # (N,)  activated (sigmoid already applied)
    shs: np.ndarray          # (N,K,3)
    sh_degree: int
    width: int
    height: int
    fx: float
    fy: float

    @property
    def n(self):
        return self.means.shape[0]


def make_scene(n: int, sh_degree: int, width: int, height: int, seed: int, max_sh_degree: int | None = None) -> Scene:
    rng = np.random.default_rng(seed)
    f32 = np.float32
    fx = fy = f32(0.6 * width)
    z = rng.uniform(1.0, 30.0, n)
    x = z * rng.uniform(-1.15, 1.15, n) * (width / (2.0 * fx))
    y = z * rng.uniform(-1.15, 1.15, n) * (height / (2.0 * fy))
    means = np.stack([x, y, z], 1).astype(f32)
    sigma_px = np.exp(rng.normal(np.log(2.2), 0.6, (n, 3)))
    scales = (sigma_px * z[:, None] / fx).astype(f32)
    rotations = rng.normal(0.0, 1.0, (n, 4)).astype(f32)
    opacities = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, n)))).astype(f32)
    K = ((sh_degree if max_sh_degree is None else max_sh_degree) + 1) ** 2
    shs = np.empty((n, K, 3), f32)
    shs[:, 0, :] = rng.uniform(-1.0, 1.5, (n, 3))
    if K > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.15, (n, K - 1, 3))
    return Scene(means, scales, rotations, opacities, shs, sh_degree, width, height, float(fx), float(fy))


def make_config(name: str) -> Scene:
    n, deg, w, h, _mode, seed, _views = CONFIGS[name]
    return make_scene(n, deg, w, h, seed)


def make_vpixels(width: int, height: int, channels: int, seed: int) -> np.ndarray:
    """Cotangent of the image, (H,W,C): N(0,1)/P."""
    rng = np.random.default_rng(seed + 7919)
    return (rng.normal(0.0, 1.0, (height, width, channels)) / (width * height)).astype(np.float32)


def view_pose(view: int, n_views: int = 8, max_yaw_deg: float = 20.0, max_shift: float = 1.0):
    """w2c rotation (3,3) and translation (3,) for view `view`: yaw in ±max_yaw_deg, x-translation in ±max_shift
    (config C3 uses ±20° / ±1; bench.py's weak-scaling runs use ±2° / ±0.1 so that every rank's view carries
    the same work as the single-GPU identity view to within a few percent)."""
    if n_views <= 1:
        return np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    a = -1.0 + 2.0 * view / (n_views - 1)
    yaw = np.deg2rad(max_yaw_deg) * a
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], np.float32)
    t = np.array([a * max_shift, 0.0, 0.0], np.float32)
    return R, t
