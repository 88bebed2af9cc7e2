"""End-to-end pins of the CPU oracle: the reference's own tiny-scene rasterizer tests
(/root/reference/test/runtests.jl:697-853) re-expressed against `Oracle.forward/backward`, plus a
whole-pipeline float64 finite-difference check of `∇rasterize` (which the reference never tests,
SURVEY.md §4) and golden fixtures under tests/golden/.
"""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, OracleCamera
from gsrast.synthetic import make_scene, make_vpixels

O32 = Oracle(np.float32)
O64 = Oracle(np.float64)
SH0 = 0.28209479177387814
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rgb_2_sh(x):  # gaussians.jl:133
    return (x - 0.5) / SH0


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def grid_scene(n_side, extent, z, log_scales, raw_opacity, rng):
    xs = np.linspace(-extent, extent, n_side, dtype=np.float32)
    pts = np.array([(x, y, z) for y in xs for x in xs], np.float32)
    n = len(pts)
    colors = rng.random((n, 3)).astype(np.float32)
    shs = rgb_2_sh(colors).reshape(n, 1, 3).astype(np.float32)
    scales = np.exp(np.tile(np.asarray(log_scales, np.float32), (n, 1)))
    rots = np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))  # identity quaternion gaussians.jl:40-42
    opac = np.full(n, sigmoid(raw_opacity), np.float32)
    return pts, shs, opac, scales.astype(np.float32), rots, colors


def test_rgbdn_normal_channel():  # runtests.jl:697-742
    rng = np.random.default_rng(0)
    W, H = 64, 48
    cam = OracleCamera.simple(100.0, 100.0, W, H)
    pts, shs, opac, scales, rots, _ = grid_scene(8, 0.6, 3.0, [np.log(0.2), np.log(0.2), np.log(0.01)], 5.0, rng)
    img, st = O32.forward(pts, shs, opac, scales, rots, cam, mode="rgbdn", sh_degree=0)
    assert img.shape == (H, W, 8)
    alpha = img[:, :, 4]
    covered = alpha > 0.5
    assert covered.any()
    assert np.abs(img[:, :, 5]).max() < 1e-4 and np.abs(img[:, :, 6]).max() < 1e-4
    np.testing.assert_allclose(img[:, :, 7][covered], -alpha[covered], atol=1e-3)
    # the normal channel's cotangent must reach the rotations
    w = np.zeros((H, W, 8), np.float32)
    w[:, :, 5:8] = rng.normal(size=(H, W, 3))
    g = O32.backward(w, pts, shs, opac, scales, rots, cam, st, mode="rgbdn", sh_degree=0)
    assert g["vrot"].shape == rots.shape and np.isfinite(g["vrot"]).all() and np.abs(g["vrot"]).max() > 0


def test_sky_composite_identity():  # runtests.jl:760-797
    rng = np.random.default_rng(1)
    W, H = 64, 48
    cam = OracleCamera.simple(100.0, 100.0, W, H)
    inv_sig = np.log(0.5 / (1 - 0.5))
    pts, shs, opac, scales, rots, _ = grid_scene(6, 0.6, 3.0, [np.log(0.1)] * 3, inv_sig, rng)
    bg = np.array([0.2, 0.7, 0.4], np.float32)
    in_kernel, _ = O32.forward(pts, shs, opac, scales, rots, cam, mode="rgbd", sh_degree=0, background=bg)
    zeroed, _ = O32.forward(pts, shs, opac, scales, rots, cam, mode="rgbd", sh_degree=0)
    alpha = zeroed[:, :, 4]
    composited = zeroed[:, :, :3] + (1 - alpha)[:, :, None] * bg
    assert alpha.min() < 1e-3 and ((alpha > 0.05) & (alpha < 0.95)).any() and alpha.max() > 0.3
    assert np.abs(in_kernel[:, :, :3] - composited).max() < 1e-5
    # depth/alpha channels get a zero background (feature_background, rasterizer.jl:411-414)
    assert (in_kernel[:, :, 3:] == zeroed[:, :, 3:]).all()


def fibonacci_sphere(n):  # sky_dome.jl:52-71
    i = np.arange(1, n + 1, dtype=np.float32)
    z = np.float32(1) - np.float32(2) * (i - np.float32(0.5)) / np.float32(n)
    r = np.sqrt(np.maximum(np.float32(1) - z * z, np.float32(0)))
    golden = np.float32(np.pi * (3.0 - np.sqrt(5.0)))
    th = golden * (i - 1)
    dirs = np.stack([r * np.cos(th), r * np.sin(th), z], 1).astype(np.float32)
    return dirs, np.float32(np.sqrt(4 * np.pi / n))


def test_sky_dome_far_plane():  # runtests.jl:799-853
    W, H = 64, 48
    cam = OracleCamera.simple(100.0, 100.0, W, H)
    radius = np.float32(50.0)
    dirs, spacing = fibonacci_sphere(8192)
    pts = (dirs * radius).astype(np.float32)
    n = len(pts)
    color = np.array([0.2, 0.4, 0.9], np.float32)
    shs = np.tile(rgb_2_sh(color).astype(np.float32), (n, 1)).reshape(n, 1, 3)
    scales = np.full((n, 3), radius * spacing, np.float32)
    rots = np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))
    opac = np.full(n, 0.99, np.float32)
    # default far plane (1000) keeps the shell; far < radius culls everything (the reason far is per-rasterizer)
    img, st = O32.forward(pts, shs, opac, scales, rots, cam, mode="rgbd", sh_degree=0, far=4 * radius)
    alpha = img[:, :, 4]
    assert alpha.min() > 0.98
    opaque = alpha > 0.99
    assert opaque.any()
    for c in range(3):
        np.testing.assert_allclose(img[:, :, c][opaque], color[c], atol=1e-2)
    img2, st2 = O32.forward(pts, shs, opac, scales, rots, cam, mode="rgbd", sh_degree=0, far=40.0)
    assert st2.n_rendered == 0 and (img2 == 0).all()  # M=0 → zero image, not background (rasterizer.jl:338)
    w = np.zeros((H, W, 5), np.float32)
    w[:, :, :3] = np.random.default_rng(3).normal(size=(H, W, 3))
    g = O32.backward(w, pts, shs, opac, scales, rots, cam, st, mode="rgbd", sh_degree=0)
    assert np.isfinite(g["vshs"]).all() and np.abs(g["vshs"]).max() > 0


def _tiny_scene(deg, seed=5):
    rng = np.random.default_rng(seed)
    n, W, H = 24, 32, 32
    fx = 40.0
    z = rng.uniform(2.0, 6.0, n)
    means = np.stack([z * rng.uniform(-0.35, 0.35, n), z * rng.uniform(-0.35, 0.35, n), z], 1)
    scales = np.exp(rng.normal(np.log(3.0), 0.3, (n, 3))) * z[:, None] / fx
    rots = rng.normal(size=(n, 4))
    opac = sigmoid(rng.normal(0, 1.0, n))
    K = (deg + 1) ** 2
    shs = np.concatenate([rng.uniform(-0.5, 1.5, (n, 1, 3)), rng.normal(0, 0.2, (n, K - 1, 3))], 1)
    yaw = 0.1
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.array([0.1, -0.05, 0.2])
    cam = OracleCamera(R, t, np.array([fx, fx * 1.1]), np.array([0.48, 0.53]), -R.T @ t, W, H)
    return means, shs, opac, scales, rots, cam


@pytest.mark.parametrize("mode,deg", [("rgb", 1), ("rgbd", 3), ("rgbdn", 0)])
def test_whole_pipeline_gradient_vs_fd_float64(mode, deg):
    """`∇rasterize` (render + project + SH pullbacks) against central FDs of `rasterize` in float64."""
    means, shs, opac, scales, rots, cam = _tiny_scene(deg)
    C = {"rgb": 3, "rgbd": 5, "rgbdn": 8}[mode]
    rng = np.random.default_rng(9)
    w = rng.normal(size=(cam.height, cam.width, C))
    bg = np.array([0.3, 0.1, 0.6])

    def loss(m=means, s=shs, o=opac, sc=scales, r=rots):
        img, _ = O64.forward(m, s, o, sc, r, cam, mode=mode, sh_degree=deg, background=bg)
        return float(np.sum(img * w))

    img, st = O64.forward(means, shs, opac, scales, rots, cam, mode=mode, sh_degree=deg, background=bg)
    assert st.n_rendered > 0 and (st.radii > 0).sum() >= 12
    g = O64.backward(w, means, shs, opac, scales, rots, cam, st, mode=mode, sh_degree=deg, background=bg)

    def check(name, arr, key, kw, n_probe=10, h=1e-6):
        flat_idx = rng.choice(arr.size, size=min(n_probe, arr.size), replace=False)
        bad = 0
        for fi in flat_idx:
            idx = np.unravel_index(fi, arr.shape)
            if st.radii[idx[0]] <= 0:
                assert g[key][idx] == 0  # culled rows rely on zero-init (projection.jl:172-176)
                continue
            ap, am = arr.copy(), arr.copy()
            ap[idx] += h
            am[idx] -= h
            fd = (loss(**{kw: ap}) - loss(**{kw: am})) / (2 * h)
            an = g[key][idx]
            if not np.isclose(an, fd, rtol=2e-4, atol=1e-6 * max(1.0, np.abs(g[key]).max())):
                bad += 1  # an FD step may straddle a discontinuity (1/255 / 1e-4 / radius thresholds)
        assert bad <= 1, f"{name}: {bad} FD mismatches"

    check("means", means, "vmeans", "m")
    check("shs", shs, "vshs", "s")
    check("opacities", opac, "vopacities", "o")
    check("scales", scales, "vscales", "sc")
    check("rotations", rots, "vrot", "r")


def test_pose_gradient_vs_fd_float64():
    """vR / vt (pose optimisation path, projection.jl:243-256) against FDs; SH degree 0 so that the
    camera centre (held fixed by the reference's pullback) does not enter."""
    means, shs, opac, scales, rots, cam = _tiny_scene(0, seed=11)
    rng = np.random.default_rng(4)
    w = rng.normal(size=(cam.height, cam.width, 5))

    def loss(R=cam.R, t=cam.t):
        c = OracleCamera(R, t, cam.focal, cam.principal, cam.cam_center, cam.width, cam.height)
        img, _ = O64.forward(means, shs, opac, scales, rots, c, mode="rgbd", sh_degree=0)
        return float(np.sum(img * w))

    img, st = O64.forward(means, shs, opac, scales, rots, cam, mode="rgbd", sh_degree=0)
    g = O64.backward(w, means, shs, opac, scales, rots, cam, st, mode="rgbd", sh_degree=0, pose_grad=True)
    h = 1e-6
    for i in range(3):
        tp, tm = cam.t.copy(), cam.t.copy()
        tp[i] += h
        tm[i] -= h
        fd = (loss(t=tp) - loss(t=tm)) / (2 * h)
        assert np.isclose(g["vt"][i], fd, rtol=5e-4, atol=1e-5), (i, g["vt"][i], fd)
        for j in range(3):
            Rp, Rm = cam.R.copy(), cam.R.copy()
            Rp[i, j] += h
            Rm[i, j] -= h
            fd = (loss(R=Rp) - loss(R=Rm)) / (2 * h)
            assert np.isclose(g["vR"][i, j], fd, rtol=5e-4, atol=1e-5 * max(1, np.abs(g["vR"]).max())), (i, j)


def test_f32_matches_f64_on_c1_subsample():
    """fp32 restatement vs the same formulas in fp64 on a 2k-Gaussian 128² scene.  Pixels where a pair sits
    within 1e-4 (relative) of a branch threshold may legitimately flip; all others agree to 1e-5."""
    sc = make_scene(2000, 2, 128, 128, 77)
    cam = OracleCamera.simple(sc.fx, sc.fy, 128, 128)
    vp = make_vpixels(128, 128, 5, 77)
    args = (sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam)
    i32, s32 = O32.forward(*args, mode="rgbd", sh_degree=2)
    i64, s64 = O64.forward(*args, mode="rgbd", sh_degree=2, ambig_rel=1e-4)
    assert (s32.radii == s64.radii).mean() > 0.999 and abs(s32.n_rendered - s64.n_rendered) <= 8
    ok = s64.ambiguous == 0
    assert ok.mean() > 0.98
    d = np.abs(i32 - i64)
    assert d[:, :, [0, 1, 2, 4]][ok].max() < 1e-5
    assert (d[:, :, 3][ok] / np.maximum(1.0, i64[:, :, 3][ok])).max() < 1e-5
    assert d[:, :, [0, 1, 2, 4]].max() < 2e-2  # a flipped pair moves a pixel by at most ~alpha*T*|feature|
    g32 = O32.backward(vp, *args, s32, mode="rgbd", sh_degree=2)
    g64 = O64.backward(vp, *args, s64, mode="rgbd", sh_degree=2)
    for k in ("vmeans", "vshs", "vopacities", "vscales", "vrot"):
        ref = g64[k]
        err = np.abs(g32[k] - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err < 2e-3, (k, err)  # fp32 vs fp64 formulas; threshold flips near 1/255 dominate


def test_stale_state_semantics():
    """Culled Gaussians keep stale means_2d/depths/conics/rgbs; only radii is cleared (projection.jl:79-82)."""
    sc = make_scene(500, 0, 64, 64, 5)
    cam = OracleCamera.simple(sc.fx, sc.fy, 64, 64)
    _, st = O32.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode="rgb", sh_degree=0)
    vis = st.radii > 0
    assert vis.any()
    before = st.means2d.copy(), st.depths.copy(), st.conics.copy(), st.rgbs.copy()
    means2 = sc.means.copy()
    means2[:, 2] = -1.0  # behind the camera: everything culled
    img, st = O32.forward(means2, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode="rgb", sh_degree=0, state=st)
    assert (st.radii == 0).all() and st.n_rendered == 0 and (img == 0).all()
    for a, b in zip(before, (st.means2d, st.depths, st.conics, st.rgbs)):
        assert (a == b).all()


def test_golden_fixture_c1_small():
    """Committed fixture (tests/golden/make_golden.py): guards the oracle itself against silent drift."""
    path = os.path.join(GOLDEN, "oracle_c1_small.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated")
    z = np.load(path)
    sc = make_scene(int(z["n"]), int(z["deg"]), int(z["w"]), int(z["h"]), int(z["seed"]))
    cam = OracleCamera.simple(sc.fx, sc.fy, sc.width, sc.height)
    img, st = O32.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode="rgbd",
                          sh_degree=sc.sh_degree)
    assert (st.radii == z["radii"]).all()
    assert (st.keys_sorted == z["keys_sorted"]).all() and (st.values_sorted == z["values_sorted"]).all()
    assert (st.ranges == z["ranges"]).all() and (st.n_contrib == z["n_contrib"]).all()
    assert (st.means2d[st.radii > 0].view(np.uint32) == z["means2d_vis_bits"]).all()
    assert (st.conics[st.radii > 0].view(np.uint32) == z["conics_vis_bits"]).all()
    np.testing.assert_allclose(img, z["image"], atol=1e-6)
