"""Pins the CPU oracle against every known-answer / finite-difference test the reference's own
test-suite holds for the hot path, re-expressed from /root/reference/test/runtests.jl.

Strategy (the reference compares Float32 adjoints with Float64 central FDs at atol=1e-3, rtol=5e-3):
here the fp64 build of the oracle gives the analytic gradient in double, a 5-point central FD of the
fp64 primal checks it at 1e-6, and the fp32 build must agree with the fp64 build at fp32 accuracy.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle

O32 = Oracle(np.float32)
O64 = Oracle(np.float64)
RNG = np.random.default_rng(12345)


def fd_grad(f, x, h=1e-5):
    """5-point central difference gradient of scalar f at x (float64)."""
    x = np.array(x, np.float64)
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index

        def at(d):
            xx = x.copy()
            xx[i] += d
            return f(xx)

        g[i] = (-at(2 * h) + 8 * at(h) - 8 * at(-h) + at(-2 * h)) / (12 * h)
    return g


def close32(a32, a64, rtol=2e-4, atol=2e-5):
    scale = max(1.0, float(np.max(np.abs(a64))))
    np.testing.assert_allclose(np.asarray(a32, np.float64), a64, rtol=rtol, atol=atol * scale)


def quat_to_mat_textbook(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_quat2mat():  # runtests.jl:86-93
    from scipy.spatial.transform import Rotation
    for _ in range(20):
        rot = Rotation.from_euler("xyz", RNG.random(3))
        x, y, z, w = rot.as_quat()
        R = O32.unnorm_quat2rot([w, x, y, z])
        np.testing.assert_allclose(R, rot.as_matrix(), atol=1e-6)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)


def test_grad_unnorm_quat2rot_fd():  # runtests.jl:95-125
    for _ in range(50):
        q = RNG.normal(size=4) * (0.3 + 2 * RNG.random())
        vR = RNG.normal(size=(3, 3))
        vq = O64.grad_unnorm_quat2rot(q, vR)
        fd = fd_grad(lambda x: np.sum(vR * O64.unnorm_quat2rot(x)), q)
        np.testing.assert_allclose(vq, fd, rtol=1e-6, atol=1e-7)
        assert abs(vq @ q) / np.linalg.norm(vq) < 1e-9  # no radial component
        close32(O32.grad_unnorm_quat2rot(q, vR), vq)
        vq32 = O32.grad_unnorm_quat2rot(q, vR).astype(np.float64)
        assert abs(vq32 @ q) / np.linalg.norm(vq32) < 1e-5


def test_grad_pos_world_to_cam_fd():  # runtests.jl:127-148
    for _ in range(50):
        R, t, p, v = RNG.normal(size=(3, 3)), RNG.normal(size=3), RNG.normal(size=3), RNG.normal(size=3)
        vR, vt, vp = O64.grad_pos_world_to_cam(R, t, p, v)
        np.testing.assert_allclose(vR, fd_grad(lambda x: v @ O64.pos_world_to_cam(x, t, p), R), atol=1e-7)
        np.testing.assert_allclose(vt, fd_grad(lambda x: v @ O64.pos_world_to_cam(R, x, p), t), atol=1e-7)
        np.testing.assert_allclose(vp, fd_grad(lambda x: v @ O64.pos_world_to_cam(R, t, x), p), atol=1e-7)
        for a, b in zip(O32.grad_pos_world_to_cam(R, t, p, v), (vR, vt, vp)):
            close32(a, b)


def test_grad_covar_world_to_cam_fd():  # runtests.jl:150-173
    for _ in range(50):
        R, A = RNG.normal(size=(3, 3)), RNG.normal(size=(3, 3))
        S = A @ A.T
        vSc, vR_in = RNG.normal(size=(3, 3)), RNG.normal(size=(3, 3))
        vR, vS = O64.grad_covar_world_to_cam(R, S, vSc, vR_in)
        np.testing.assert_allclose(vR - vR_in, fd_grad(lambda x: np.sum(vSc * O64.covar_world_to_cam(x, S)), R),
                                   rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(vS, fd_grad(lambda x: np.sum(vSc * O64.covar_world_to_cam(R, x)), S),
                                   rtol=1e-6, atol=1e-6)
        a32, b32 = O32.grad_covar_world_to_cam(R, S, vSc, vR_in)
        close32(a32, vR)
        close32(b32, vS)


@pytest.mark.parametrize("inside", [True, False])
def test_grad_perspective_projection_fd(inside):  # runtests.jl:175-216
    focal = np.array([1000.0, 1000.0])
    res = np.array([1920, 1080], np.int32)
    principal = np.array([0.5, 0.5])
    tan_fov = 0.5 * res / focal
    lim = (res - principal * res) / focal + 0.3 * tan_fov
    for _ in range(30):
        if inside:
            ratio = (2 * RNG.random(2) - 1) * 0.5 * lim
        else:
            ratio = np.sign(RNG.normal(size=2)) * (1.2 + 0.5 * RNG.random(2)) * lim
        z = 2 + 4 * RNG.random()
        mean = np.array([ratio[0] * z, ratio[1] * z, z])
        A = 0.1 * RNG.normal(size=(3, 3))
        S = A @ A.T
        vS2, vm2 = RNG.normal(size=(2, 2)), RNG.normal(size=2)
        vS, vmean = O64.grad_perspective_projection(mean, S, focal, res, principal, vS2, vm2)

        def loss(m, s):
            S2, m2 = O64.perspective_projection(m, s, focal, res, principal)
            return np.sum(vS2 * S2) + vm2 @ m2

        np.testing.assert_allclose(vmean, fd_grad(lambda x: loss(x, S), mean, h=1e-4), rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(vS, fd_grad(lambda x: loss(mean, x), S, h=1e-4), rtol=1e-5, atol=1e-4)
        a32, b32 = O32.grad_perspective_projection(mean, S, focal, res, principal, vS2, vm2)
        close32(a32, vS, rtol=1e-3)
        close32(b32, vmean, rtol=1e-3)


def test_grad_quat_scale_to_cov_fd():  # runtests.jl:218-239
    for _ in range(50):
        q = RNG.normal(size=4) * (0.3 + 2 * RNG.random())
        s = np.exp(0.5 * RNG.normal(size=3))
        R = O64.unnorm_quat2rot(q)
        vS = RNG.normal(size=(3, 3))
        vq, vs = O64.grad_quat_scale_to_cov(q, s, R, vS)
        loss = lambda qq, ss: np.sum(vS * O64.quat_scale_to_cov(O64.unnorm_quat2rot(qq), ss))
        np.testing.assert_allclose(vq, fd_grad(lambda x: loss(x, s), q), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(vs, fd_grad(lambda x: loss(q, x), s), rtol=1e-6, atol=1e-6)
        a32, b32 = O32.grad_quat_scale_to_cov(q, s, O32.unnorm_quat2rot(q), vS)
        close32(a32, vq, rtol=1e-3)
        close32(b32, vs, rtol=1e-3)


def test_grad_inverse_fd():  # runtests.jl:241-266
    for _ in range(50):
        A = RNG.normal(size=(2, 2))
        X = A @ A.T + 0.5 * np.eye(2)
        b = RNG.normal(size=3)
        vY = np.array([[b[0], b[1]], [b[1], b[2]]])
        _, Y = O64.inverse(X)
        np.testing.assert_allclose(Y, np.linalg.inv(X), rtol=1e-10)
        vX = O64.grad_inverse(Y, vY)
        loss = lambda p: np.sum(vY * O64.inverse(np.array([[p[0], p[1]], [p[1], p[2]]]))[1])
        fd = fd_grad(loss, [X[0, 0], X[1, 0], X[1, 1]])
        np.testing.assert_allclose([vX[0, 0], vX[0, 1] + vX[1, 0], vX[1, 1]], fd, rtol=1e-6, atol=1e-6)
        close32(O32.grad_inverse(Y, vY), vX, rtol=1e-3)


def test_grad_add_blur_fd():  # runtests.jl:268-291
    eps = 0.3
    for _ in range(50):
        A = RNG.normal(size=(2, 2))
        S = A @ A.T + 0.5 * np.eye(2)
        vcomp = RNG.normal()
        Sb, _, comp = O64.add_blur(S, eps)
        _, conic = O64.inverse(Sb)
        vS = O64.grad_add_blur(comp, vcomp, conic, eps)
        loss = lambda p: vcomp * O64.add_blur(np.array([[p[0], p[1]], [p[1], p[2]]]), eps)[2]
        fd = fd_grad(loss, [S[0, 0], S[1, 0], S[1, 1]])
        # the reference's adjoint carries a +1e-6 regulariser in the denominator: same tolerance as runtests.jl
        np.testing.assert_allclose([vS[0, 0], vS[0, 1] + vS[1, 0], vS[1, 1]], fd, rtol=5e-3, atol=1e-4)


def test_grad_normalize_fd():  # runtests.jl:293-306
    for _ in range(50):
        d = RNG.normal(size=3) * (0.3 + 2 * RNG.random())
        vd = RNG.normal(size=3)
        out = O64.grad_normalize(d, vd)
        fd = fd_grad(lambda x: vd @ (x / np.linalg.norm(x)), d)
        np.testing.assert_allclose(out, fd, rtol=1e-6, atol=1e-7)
        close32(O32.grad_normalize(d, vd), out, rtol=1e-3)


def test_get_rect_known_answers():  # runtests.jl:308-324
    grid = [1024 // 16, 1024 // 16]
    for o in (O32, O64):
        assert o.get_rect([0, 0], 1, grid) == ((0, 0), (1, 1))
        assert o.get_rect([0, 0], 17, grid) == ((0, 0), (2, 2))
    # clamping to the grid on both sides
    assert O32.get_rect([5000.0, -5000.0], 3, grid) == ((64, 0), (64, 0))


def test_tile_ranges_known_answer():  # runtests.jl:486-494
    keys = np.array([0 << 32, 0 << 32, 1 << 32, 2 << 32, 3 << 32], np.uint64)
    ranges = O32.identify_tile_range(keys, 4)
    assert ranges.tolist() == [[0, 2], [2, 3], [3, 4], [4, 5]]
    # untouched tiles keep the pre-filled zeros (rasterizer.jl:375)
    keys = np.array([1 << 32, 1 << 32, 5 << 32], np.uint64) | np.uint64(0x3F800000)
    assert O32.identify_tile_range(keys, 7).tolist() == [[0, 0], [0, 2], [0, 0], [0, 0], [0, 0], [2, 3], [0, 0]]


def test_gaussian_normal_properties():  # runtests.jl:555-575
    from scipy.spatial.transform import Rotation
    for _ in range(50):
        q = RNG.normal(size=4) * (0.3 + 2 * RNG.random())
        scale = np.exp(0.5 * RNG.normal(size=3))
        Rw = Rotation.random(random_state=int(RNG.integers(1 << 30))).as_matrix()
        Rg = O32.unnorm_quat2rot(q)
        mc = np.array([RNG.normal(), RNG.normal(), 1 + 5 * RNG.random()])
        n, k, s = O32.gaussian_normal(Rw, Rg, scale, mc)
        assert abs(np.linalg.norm(n) - 1) < 1e-5
        assert n @ mc <= 1e-6
        assert np.float32(scale[k - 1]) == np.float32(scale).min()
        assert abs(s) == 1.0
        np.testing.assert_allclose(n, s * (Rw @ Rg[:, k - 1]), atol=1e-6)


def test_grad_gaussian_normal_fd():  # runtests.jl:577-611
    from scipy.spatial.transform import Rotation
    done = 0
    while done < 30:
        q = RNG.normal(size=4) * (0.3 + 2 * RNG.random())
        scale = np.exp(np.array([0.0, 1.0, 2.0]) + 0.1 * RNG.normal(size=3))
        Rw = Rotation.random(random_state=int(RNG.integers(1 << 30))).as_matrix()
        mc = np.array([RNG.normal(), RNG.normal(), 2 + 5 * RNG.random()])
        vn = RNG.normal(size=3)
        Rg = O64.unnorm_quat2rot(q)
        n, k, s = O64.gaussian_normal(Rw, Rg, scale, mc)
        if abs(n @ (mc / np.linalg.norm(mc))) <= 0.1:
            continue
        done += 1
        vRg = np.zeros((3, 3))
        vRg[:, k - 1] = s * (Rw.T @ vn)
        vq, vscale = O64.grad_quat_scale_to_cov(q, scale, Rg, np.zeros((3, 3)), vRg)
        assert np.all(vscale == 0)
        loss = lambda qq: vn @ O64.gaussian_normal(Rw, O64.unnorm_quat2rot(qq), scale, mc)[0]
        np.testing.assert_allclose(vq, fd_grad(loss, q), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_sh_color_and_gradient(degree):
    """Not covered by the reference's tests: SH deg 0-3 forward vs an independent real-SH
    evaluation and backward vs FD."""
    K = 16
    for _ in range(20):
        p, cam = RNG.normal(size=3) * 3, RNG.normal(size=3)
        shs = RNG.normal(size=(K, 3)) * 0.3
        rgb, cl = O64.colors_from_sh(p, cam, shs, degree)
        d = (p - cam) / np.linalg.norm(p - cam)
        x, y, z = d
        basis = [0.28209479177387814, -0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x,
                 1.0925484305920792 * x * y, -1.0925484305920792 * y * z,
                 0.31539156525252005 * (2 * z * z - x * x - y * y), -1.0925484305920792 * x * z,
                 0.5462742152960396 * (x * x - y * y), -0.5900435899266435 * y * (3 * x * x - y * y),
                 2.890611442640554 * x * y * z, -0.4570457994644658 * y * (4 * z * z - x * x - y * y),
                 0.3731763325901154 * z * (2 * z * z - 3 * x * x - 3 * y * y),
                 -0.4570457994644658 * x * (4 * z * z - x * x - y * y), 1.445305721320277 * z * (x * x - y * y),
                 -0.5900435899266435 * x * (x * x - 3 * y * y)]
        k = (degree + 1) ** 2
        ref = np.array(basis[:k]) @ shs[:k] + 0.5 + np.finfo(np.float32).eps
        # the fp64 build keeps the reference's Float32 SH constants (widened), hence 1e-7 and not 1e-12
        np.testing.assert_allclose(rgb, np.maximum(ref, 0), atol=1e-7)
        assert (cl == (ref < 0)).all()
        rgb32, cl32 = O32.colors_from_sh(p, cam, shs, degree)
        np.testing.assert_allclose(rgb32, rgb, atol=2e-6)
        vcol = RNG.normal(size=3)
        vshs, vmean = O64.grad_color_from_sh(p, cam, shs, degree, cl, vcol)
        mask = 1.0 - cl

        def loss_s(s):
            return (vcol * mask) @ (np.array(basis[:k]) @ s[:k])

        np.testing.assert_allclose(vshs, fd_grad(loss_s, shs), atol=1e-7)
        assert np.all(vshs[k:] == 0)

        def loss_p(pp):
            c, _ = O64.colors_from_sh(pp, cam, shs + 0.0, degree)
            # un-clamped colour so the FD sees the same mask as the adjoint
            dd = (pp - cam) / np.linalg.norm(pp - cam)
            return (vcol * mask) @ _sh_eval(dd, shs, degree)

        np.testing.assert_allclose(vmean, fd_grad(loss_p, p), rtol=1e-5, atol=1e-7)
        a32, b32 = O32.grad_color_from_sh(p, cam, shs, degree, cl, vcol)
        close32(a32, vshs, rtol=1e-3)
        close32(b32, vmean, rtol=2e-3, atol=2e-5)


def _sh_eval(d, shs, degree):
    x, y, z = d
    basis = [0.28209479177387814, -0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x,
             1.0925484305920792 * x * y, -1.0925484305920792 * y * z,
             0.31539156525252005 * (2 * z * z - x * x - y * y), -1.0925484305920792 * x * z,
             0.5462742152960396 * (x * x - y * y), -0.5900435899266435 * y * (3 * x * x - y * y),
             2.890611442640554 * x * y * z, -0.4570457994644658 * y * (4 * z * z - x * x - y * y),
             0.3731763325901154 * z * (2 * z * z - 3 * x * x - 3 * y * y),
             -0.4570457994644658 * x * (4 * z * z - x * x - y * y), 1.445305721320277 * z * (x * x - y * y),
             -0.5900435899266435 * x * (x * x - 3 * y * y)]
    k = (degree + 1) ** 2
    return np.array(basis[:k]) @ shs[:k]


def test_sort_is_stable_and_ascending():
    m = 20000
    keys = (RNG.integers(0, 50, m).astype(np.uint64) << np.uint64(32)) | RNG.integers(0, 40, m).astype(np.uint64)
    vals = np.arange(1, m + 1, dtype=np.uint32)
    ks, vs = O32.sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert (ks == keys[order]).all() and (vs == vals[order]).all()
    ks0, vs0 = O32.sort_pairs(keys[:0], vals[:0])
    assert len(ks0) == 0 and len(vs0) == 0


def test_update_stats():  # strategy.jl:118-136
    n = 100
    radii = RNG.integers(-1, 30, n).astype(np.int32)
    g = RNG.normal(size=(n, 2)).astype(np.float32)
    mr = RNG.integers(0, 20, n).astype(np.int32)
    acc, den = RNG.random(n).astype(np.float32), RNG.integers(0, 5, n).astype(np.float32)
    mr0, acc0, den0 = mr.copy(), acc.copy(), den.copy()
    O32.update_stats(radii, g, 1312, 848, mr, acc, den)
    vis = radii > 0
    assert (mr[~vis] == mr0[~vis]).all() and (acc[~vis] == acc0[~vis]).all() and (den[~vis] == den0[~vis]).all()
    assert (mr[vis] == np.maximum(mr0[vis], radii[vis])).all()
    np.testing.assert_allclose(acc[vis], acc0[vis] + np.hypot(g[vis, 0] * 1312 * 0.5, g[vis, 1] * 848 * 0.5), rtol=1e-6)
    assert (den[vis] == den0[vis] + 1).all()
