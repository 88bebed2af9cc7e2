"""GPU parity of the fused SSIM kernels and the photometric loss (csrc/ssim.cu) against the oracle restatement of
src/fused_ssim.jl, through the C ABI.  Tolerances: 1e-5 absolute on the SSIM map, 1e-4 relative (to the tensor's
max) on derivative maps and gradients — the floating-point bars north_star states for images / gradients."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
MAP_ATOL = 1e-5
GRAD_RTOL = 1e-4


def _mods():
    from gsrast import ssim
    from oracle.oracle import Oracle
    return ssim, Oracle


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("shape", [(2, 3, 128, 128), (1, 3, 37, 53), (1, 1, 7, 5), (1, 3, 16, 16), (3, 2, 33, 100)])
def test_ssim_forward_backward_match_oracle(shape):  # runtests.jl:512-519 shapes + ragged / tiny ones
    ssim, Oracle = _mods()
    rng = np.random.default_rng(sum(shape))
    x, ref = rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32)
    dl = rng.normal(0, 1, shape).astype(np.float32) / x.size
    o = Oracle(np.float32)
    m_o, d0_o, d1_o, d2_o = o.fused_ssim(x, ref, train=True)
    g_o = o.fused_ssim_bwd(x, ref, dl, d0_o, d1_o, d2_o)
    xd, rd, dld = (torch.from_numpy(a).cuda() for a in (x, ref, dl))
    m, d0, d1, d2 = ssim.ssim_forward(xd, rd, train=True)
    g = ssim.ssim_backward(xd, rd, dld, d0, d1, d2)
    torch.cuda.synchronize()
    assert np.abs(m.cpu().numpy() - m_o).max() <= MAP_ATOL
    for got, want in ((d0, d0_o), (d1, d1_o), (d2, d2_o), (g, g_o)):
        assert _rel(got.cpu().numpy(), want) <= GRAD_RTOL
    # inference path writes the same map and needs no derivative buffers
    m2, a, b, c = ssim.ssim_forward(xd, rd, train=False)
    assert a is None and torch.equal(m2, m)


def test_reference_known_answers_on_gpu():  # runtests.jl:499-510
    ssim, _ = _mods()
    ones = torch.ones((1, 3, 16, 16), device="cuda")
    zeros = torch.zeros_like(ones)
    assert abs(float(ssim.fused_ssim(ones, zeros).mean())) <= 1e-4
    assert float(ssim.fused_ssim(ones, ones).mean()) == pytest.approx(1.0, rel=1e-6)
    x = torch.zeros_like(ones)
    x[:, :, 0:4, 0:4] = 0.25
    x[:, :, 0:4, 4:8] = 0.5
    x[:, :, 12:16, 8:12] = 0.75
    x[:, :, 12:16, 12:16] = 1.0
    assert float(ssim.fused_ssim(x, ones).mean()) == pytest.approx(0.1035, abs=1e-3, rel=1e-3)


def test_autograd_rule_matches_oracle_pullback():  # the rrule, fused_ssim.jl:397-407
    ssim, Oracle = _mods()
    rng = np.random.default_rng(3)
    shape = (1, 3, 48, 80)
    x, ref = rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32)
    xd = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = 1.0 - ssim.fused_ssim(xd, torch.from_numpy(ref).cuda()).mean()
    loss.backward()
    o = Oracle(np.float64)
    m, d0, d1, d2 = o.fused_ssim(x, ref, train=True)
    g = o.fused_ssim_bwd(x, ref, np.full(shape, -1.0 / x.size), d0, d1, d2)
    assert float(loss) == pytest.approx(1.0 - m.mean(), rel=1e-5)
    assert _rel(xd.grad.cpu().numpy(), g) <= GRAD_RTOL


@pytest.mark.parametrize("mode,W,H", [("rgb", 64, 48), ("rgbd", 256, 256), ("rgbdn", 1920, 1088)])
def test_photometric_loss_matches_oracle(mode, W, H):  # training.jl:684-699
    ssim, Oracle = _mods()
    from gsrast import GaussianRasterizer
    rast = GaussianRasterizer(width=W, height=H, mode=mode)
    C = rast.channels
    rng = np.random.default_rng(W + H)
    img = rng.random((H, W, C), dtype=np.float32)
    tgt = rng.random((3, H, W), dtype=np.float32)
    # a few exactly equal pixels: sign(0) = 0 in the L1 pullback
    img[0, :5, :3] = tgt[:, 0, :5].T
    loss, v = ssim.photometric_loss(rast, torch.from_numpy(img).cuda(), torch.from_numpy(tgt).cuda(), 0.2)
    torch.cuda.synchronize()
    total, l1, sm, v_o = Oracle(np.float32).photometric_loss(img, tgt, 0.2)
    got = loss.cpu().numpy()
    assert got[0] == pytest.approx(total, rel=2e-5) and got[1] == pytest.approx(l1, rel=2e-5) and got[2] == pytest.approx(sm, rel=2e-5, abs=1e-6)
    v = v.cpu().numpy()
    assert (v[:, :, 3:] == 0).all()
    assert _rel(v[:, :, :3], v_o[:, :, :3]) <= GRAD_RTOL
    assert (v[0, :5, :3] == v_o[0, :5, :3]).all() or _rel(v[0, :5, :3], v_o[0, :5, :3]) <= GRAD_RTOL


def test_loss_cotangent_feeds_the_rasterizer_backward():
    """End to end on the path: raster image -> gsr_photometric_loss -> gsr_backward, against the oracle chain."""
    ssim, Oracle = _mods()
    import parity as P
    from gsrast import GaussianRasterizer
    from gsrast.synthetic import make_scene
    sc = make_scene(3000, 1, 128, 96, 77)
    cam, ocam = P.cameras(sc)
    dev = P.to_dev(sc)
    rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd")
    img = P.gpu_forward(rast, dev, cam, sc.sh_degree)
    tgt = np.random.default_rng(9).random((3, 96, 128), dtype=np.float32)
    loss, vpix = ssim.photometric_loss(rast, img, torch.from_numpy(tgt).cuda(), 0.2)
    grads = P.gpu_backward(rast, dev, cam, sc.sh_degree, vpix)
    o = P.oracle()
    ref_img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode="rgbd",
                            sh_degree=sc.sh_degree, ambig_rel=P.AMBIG_REL)
    total, l1, sm, v_o = o.photometric_loss(ref_img, tgt, 0.2)
    assert float(loss[0]) == pytest.approx(total, rel=1e-4)
    ref = o.backward(v_o, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode="rgbd",
                     sh_degree=sc.sh_degree)
    print("loss -> raster backward:", P.assert_grads_close(grads, ref, ambig_g=st.ambiguous_g))


def test_ssim_error_conventions():
    ssim, _ = _mods()
    from gsrast import _lib
    x = torch.ones((1, 3, 8, 8), device="cuda")
    with pytest.raises(ValueError):
        ssim.ssim_forward(x, torch.ones((1, 3, 8, 9), device="cuda"))
    with pytest.raises(_lib.GsrError):  # null output pointer -> GSR_EINVAL with a message
        _lib.check(_lib.lib().gsr_ssim_forward(8, 8, 3, 1, x.data_ptr(), x.data_ptr(), 1e-4, 9e-4, 0, None, None, None,
                                               None, None))
    empty = torch.ones((0, 3, 8, 8), device="cuda")
    m, *_ = ssim.ssim_forward(empty, empty, train=False)
    assert m.numel() == 0
