"""Shared helpers of the `-m gpu` parity tests: run the CUDA path through the C ABI (via the gsrast
binding), run the CPU oracle on the same seeded inputs, and compare with the tolerances north_star states.

Tolerances (BASELINE.json north_star / SURVEY.md §8d):
  * projection, sort keys, tile ranges (and every integer buffer): BIT-EXACT;
  * forward image / accum_alpha / uncertainties: <= 1e-5 ABSOLUTE on every channel (depth included) for the default
    math_mode="strict" and for math_mode="reference" — the letter of north_star, no scaling of any kind;
    the opt-in math_mode="fast" keeps 1e-5 absolute on the unit-scale channels only up to a conditioning term and
    holds the depth channel to 1e-5 * max(1, max visible depth) (see assert_image_close);
  * gradients: <= 1e-4 relative  (||Δ||∞ / ||ref||∞ per tensor) against the fp32 oracle for "strict" / "reference".
Pixels where the oracle saw a pair within `AMBIG_REL` of one of the kernel's branch thresholds
(σ<0, α<1/255, T'<1e-4) may legitimately take the other branch when exp() differs by an ulp
(libdevice expf vs glibc expf); they are excluded from the 1e-5 check, counted, and bounded.
"""
import numpy as np
import torch

from oracle.oracle import Oracle, OracleCamera
from gsrast import Camera, GaussianRasterizer

AMBIG_REL = 2e-5        # math_mode="strict" / "reference": sigma is bit-identical to the oracle, only exp() differs by ulps
AMBIG_REL_FAST = 2e-4   # math_mode="fast": contracted / prescaled sigma differs by ~1e-5 absolute near the thresholds
EPS_FAST = 2.5e-7       # 2 ulp: ex2.approx vs a correctly rounded exp, multiplied by the pixel's conditioning (see assert_image_close)
AMBIG_COND_FAST = 2.5e-6  # ... plus ~6 roundings (6e-8 each, both evaluations) of sigma's largest term (cancellation for elongated Gaussians)
IMG_ATOL = 1e-5
GRAD_RTOL = 1e-4
CH = {"rgb": 3, "rgbd": 5, "rgbdn": 8}

_ORACLES = {}


def oracle(dtype=np.float32):
    if dtype not in _ORACLES:
        _ORACLES[dtype] = Oracle(dtype)
    return _ORACLES[dtype]


def cameras(scene_or_dims, fx=None, fy=None, R=None, t=None, principal=(0.5, 0.5)):
    if hasattr(scene_or_dims, "fx"):
        W, H, fx, fy = scene_or_dims.width, scene_or_dims.height, scene_or_dims.fx, scene_or_dims.fy
    else:
        W, H = scene_or_dims
    R = np.eye(3, dtype=np.float32) if R is None else np.asarray(R, np.float32)
    t = np.zeros(3, np.float32) if t is None else np.asarray(t, np.float32)
    cam = Camera(fx=float(fx), fy=float(fy), width=W, height=H, R=R, t=t, principal=tuple(principal))
    ocam = OracleCamera(R.astype(np.float64), t.astype(np.float64), np.array([fx, fy], np.float64),
                        np.array(principal, np.float64), cam.camera_center.astype(np.float64), W, H)
    return cam, ocam


def to_dev(sc):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return dict(means=d(sc.means), shs=d(sc.shs), opac=d(sc.opacities.reshape(-1, 1)), scales=d(sc.scales),
                rots=d(sc.rotations))


def gpu_forward(rast, dev, cam, sh_degree, background=(0, 0, 0), covis=None, uncert=None, R_w2c=None, t_w2c=None):
    img = rast._forward(dev["means"], dev["shs"], dev["opac"], dev["scales"], dev["rots"], R_w2c, t_w2c, cam,
                        sh_degree, background, covis, uncert)
    torch.cuda.synchronize()
    return img


def gpu_backward(rast, dev, cam, sh_degree, vpix, background=(0, 0, 0), R_w2c=None, t_w2c=None, **kw):
    g = rast._backward(vpix, dev["means"], dev["shs"], dev["opac"], dev["scales"], dev["rots"], R_w2c, t_w2c, cam,
                       sh_degree, background, **kw)
    torch.cuda.synchronize()
    return g


def np_(t):
    return t.detach().cpu().numpy()


def assert_forward_state_bit_exact(rast, st, n):
    """Every per-Gaussian / per-instance / per-tile buffer of the forward against the oracle, bit for bit."""
    gs = rast.gstate
    radii = np_(gs.radii)
    assert (radii == st.radii[:n]).all(), f"radii differ at {np.flatnonzero(radii != st.radii[:n])[:10]}"
    vis = radii > 0
    for name, got, ref in (("means2d", gs.means2d, st.means2d), ("depths", gs.depths, st.depths),
                           ("conics", gs.conics, st.conics), ("rgbs", gs.rgbs, st.rgbs)):
        a = np_(got)[vis].view(np.uint32)
        b = np.ascontiguousarray(ref[:n][vis]).view(np.uint32)
        bad = np.flatnonzero((a != b).reshape(len(a), -1).any(1))
        assert bad.size == 0, f"{name}: {bad.size} visible rows differ bitwise, first {bad[:5]}"
    assert (np_(gs.clamped)[vis] == st.clamped[:n][vis]).all()
    assert (np_(gs.tiles_touched) == st.tiles_touched[:n]).all()
    assert (np_(gs.points_offset) == st.points_offset[:n]).all()
    if st.normals is not None:
        assert (np_(gs.normals)[vis].view(np.uint32) == st.normals[:n][vis].view(np.uint32)).all()
    assert gs.n_rendered == st.n_rendered
    if st.n_rendered:
        assert (np_(gs.keys_unsorted).view(np.uint64) == st.keys_unsorted).all()
        assert (np_(gs.values_unsorted).view(np.uint32) == st.values_unsorted).all()
        assert (np_(gs.keys_sorted).view(np.uint64) == st.keys_sorted).all()
        assert (np_(gs.values_sorted).view(np.uint32) == st.values_sorted).all()
        assert (np_(gs.ranges).view(np.uint32) == st.ranges).all()


def assert_image_close(img, st, ref_img, strict=False, max_ambig_frac=0.02):
    """Image parity on every non-ambiguous pixel; ambiguous ones are bounded by one flipped pair.

    strict=True (math_mode="strict" and "reference"): 1e-5 ABSOLUTE on every channel, depth included — the letter of
    north_star; no conditioning term, no depth rescale.
    strict=False (math_mode="fast"): 1e-5 * featmax_c + EPS_FAST * cond * featmax_c per pixel and channel, where
    featmax_c is the largest per-Gaussian feature of the channel (1 for rgb/alpha/normal, the far visible depth for
    the depth channel) and cond is the oracle's first-order error amplification of the pixel (orc_render): every
    alpha of the fast path differs from the reference-order one by up to EPS_FAST (ex2.approx, 2 ulp) plus half
    of that per unit of sigma's largest term (FMA-contracted, prescaled quadratic form); the pair's own weight
    alpha*T and the transmittance of everything behind it (conditioning alpha/(1-alpha)) inherit that error.  The
    reference on a GPU (libdevice expf <= 2 ulp, LLVM-contracted sigma) has the same sensitivity."""
    img = np_(img)
    C = img.shape[2]
    ok = st.ambiguous == 0
    assert (1.0 - ok.mean()) <= max_ambig_frac, f"ambiguous fraction {1 - ok.mean():.4f}"
    d = np.abs(img.astype(np.float64) - ref_img.astype(np.float64))
    featmax = np.ones(C)
    if C > 3 and not strict:
        vis = st.radii > 0
        featmax[3] = max(1.0, float(st.depths[: len(vis)][vis].max()) if vis.any() else 1.0)
    tol = np.broadcast_to(IMG_ATOL * featmax, d.shape).copy()
    if not strict:
        tol = tol + EPS_FAST * st.cond.astype(np.float64)[:, :, None] * featmax[None, None, :]
    bad = (d > tol) & ok[:, :, None]
    assert not bad.any(), (f"{bad.sum()} non-ambiguous values beyond tolerance (featmax {featmax.tolist()}): max err per channel "
                           f"{[float(d[:, :, c][ok].max()) for c in range(C)]} at {np.argwhere(bad)[:5].tolist()}; "
                           f"(err, tol, cond, T, n_contrib) there: "
                           f"{[(float(d[tuple(i)]), float(tol[tuple(i)]), float(st.cond[i[0], i[1]]), float(st.accum_alpha[i[0], i[1]]), int(st.n_contrib[i[0], i[1]])) for i in np.argwhere(bad)[:5]]}")
    scale = max(1.0, float(np.abs(ref_img).max()))
    assert d.max() <= 2e-2 * scale, f"ambiguous pixel error {d.max():.3e} too large for a single flipped pair"
    return dict(max_err=float(d[ok].max()) if ok.any() else 0.0,
                max_err_unit_channels=float(np.delete(d, 3, axis=2)[ok].max()) if C > 3 else float(d[ok].max()),
                ambiguous=int((~ok).sum()), max_err_ambiguous=float(d.max()))


def assert_ncontrib(rast, st, max_mismatch_frac=1e-4):
    nc = np_(rast.gstate.n_contrib).view(np.uint32)
    ok = st.ambiguous == 0
    assert (nc[ok] == st.n_contrib[ok]).all(), "n_contrib differs on non-ambiguous pixels"
    T = np_(rast.gstate.accum_alpha)
    assert np.abs(T[ok] - st.accum_alpha[ok]).max() <= IMG_ATOL
    return float((nc != st.n_contrib).mean())


def rel_err(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_grads_close(g, ref, rtol=GRAD_RTOL, keys=("vmeans", "vshs", "vopacities", "vscales", "vrot"), ambig_g=None,
                       ambig_rtol=5e-2):
    """||Δ||∞ / ||ref||∞ <= 1e-4 per tensor.  Gaussians flagged by the oracle as owning a pair within AMBIG_REL of a
    branch threshold (`ambig_g`) gain or lose that pair's whole contribution when the branch flips, so they are
    held to `ambig_rtol` instead (and counted)."""
    out, errors = {}, []
    keep = None if ambig_g is None else (np.asarray(ambig_g) == 0)
    for k in keys:
        a = np_(g[k]).reshape(ref[k].shape) if isinstance(g[k], torch.Tensor) else g[k]
        assert np.isfinite(a).all(), f"{k} has non-finite values"
        scale = max(float(np.abs(ref[k]).max()), 1e-30)
        d = np.abs(a.astype(np.float64) - ref[k].astype(np.float64)).reshape(a.shape[0], -1).max(1) / scale
        if keep is None:
            out[k] = float(d.max())
        else:
            out[k] = float(d[keep].max()) if keep.any() else 0.0
            worst_amb = float(d[~keep].max()) if (~keep).any() else 0.0
            out[k + "_ambiguous"] = worst_amb
            if worst_amb > ambig_rtol:
                errors.append(f"{k}: ambiguous-Gaussian error {worst_amb:.3e} > {ambig_rtol:g}")
        if out[k] > rtol:
            worst = int(np.argmax(np.where(keep, d, 0) if keep is not None else d))
            errors.append(f"{k}: relative error {out[k]:.3e} > {rtol:g} (row {worst})")
    if keep is not None:
        out["ambiguous_gaussians"] = int((~keep).sum())
    assert not errors, "; ".join(errors) + f" | all: {out}"
    return out


def assert_grads_as_accurate_as_reference(g, ref32, ref64, rtol=GRAD_RTOL,
                                           keys=("vmeans", "vshs", "vopacities", "vscales", "vrot"), ambig_g=None,
                                           ambig_rtol=5e-2):
    """Full-size criterion for math_mode="fast".  At 1M Gaussians the reference's own fp32 arithmetic is up to ~8e-4
    (of each tensor's max) away from an fp64 evaluation of the same formulas: (1-alpha) loses 2 digits for
    near-opaque Gaussians, T is a product of hundreds of such factors, and the T'<1e-4 termination flips on
    thousands of pixels between fp32 and fp64.  A different-but-valid fp32 evaluation order cannot match the fp32
    oracle better than that noise, so the fast path is required to be AS ACCURATE AS THE REFERENCE ARITHMETIC,
    measured against fp64 with errors relative to each tensor's max: worst row no worse than the fp32 oracle's
    worst row + rtol, rms error <= 1.05x the fp32 oracle's, and at most 1e-5 of the rows exceed the fp32 oracle's
    own error by more than rtol (none by more than 1e-3).
    Gaussians owning a pair near a branch threshold (`ambig_g`, from either oracle) are held to `ambig_rtol`."""
    out, errors = {}, []
    keep = None if ambig_g is None else (np.asarray(ambig_g) == 0)
    for k in keys:
        a = np_(g[k]).reshape(ref64[k].shape).astype(np.float64)
        assert np.isfinite(a).all(), f"{k} has non-finite values"
        scale = max(float(np.abs(ref64[k]).max()), 1e-30)
        rows = a.shape[0]
        d_gpu = np.abs(a - ref64[k]).reshape(rows, -1).max(1) / scale
        d_ref = np.abs(ref32[k].astype(np.float64) - ref64[k]).reshape(rows, -1).max(1) / scale
        if keep is not None:
            if (~keep).any() and (d_gpu - d_ref)[~keep].max() > ambig_rtol:
                errors.append(f"{k}: ambiguous row exceeds the reference's fp32 error by {(d_gpu - d_ref)[~keep].max():.3e}")
            d_gpu, d_ref = d_gpu[keep], d_ref[keep]
        excess = d_gpu - d_ref
        out[k] = dict(max_gpu=float(d_gpu.max()), max_ref=float(d_ref.max()), rms_gpu=float(np.sqrt((d_gpu ** 2).mean())),
                      rms_ref=float(np.sqrt((d_ref ** 2).mean())), worst_excess=float(excess.max()),
                      rows_over_rtol=float((excess > rtol).mean()), closer_than_reference=float((d_gpu <= d_ref).mean()))
        # the comparison is distributional: one realisation of fp32 rounding (the fp32 oracle) can be lucky on a row
        if out[k]["max_gpu"] > out[k]["max_ref"] + rtol:
            errors.append(f"{k}: worst-row error {out[k]['max_gpu']:.3e} vs reference arithmetic {out[k]['max_ref']:.3e}")
        if out[k]["rms_gpu"] > 1.05 * out[k]["rms_ref"] + 1e-8:
            errors.append(f"{k}: rms error {out[k]['rms_gpu']:.3e} vs reference arithmetic {out[k]['rms_ref']:.3e}")
        if out[k]["rows_over_rtol"] > 1e-5 or out[k]["worst_excess"] > 1e-3:
            errors.append(f"{k}: {out[k]['rows_over_rtol']:.2e} of rows exceed the reference's fp32 error by > {rtol:g} "
                          f"(worst {out[k]['worst_excess']:.3e})")
    assert not errors, "; ".join(errors) + f" | all: {out}"
    return out


def is_flat(math_mode):
    """Modes held to the flat north_star tolerances (everything but the opt-in fast mode)."""
    return math_mode != "fast"


def run_case(sc, mode, math_mode="strict", background=(0, 0, 0), R=None, t=None, principal=(0.5, 0.5),
             check_backward=True, near=0.2, far=1000.0, vpix_seed=1, grad_rtol=GRAD_RTOL):
    """Full forward(+backward) parity of one scene; returns a dict of measured errors."""
    from gsrast.synthetic import make_vpixels
    cam, ocam = cameras(sc, R=R, t=t, principal=principal)
    dev = to_dev(sc)
    rast = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, math_mode=math_mode, near_plane=near,
                              far_plane=far)
    img = gpu_forward(rast, dev, cam, sc.sh_degree, background)
    o = oracle()
    ref_img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode=mode,
                            sh_degree=sc.sh_degree, background=background, near=near, far=far,
                            ambig_rel=AMBIG_REL if is_flat(math_mode) else AMBIG_REL_FAST,
                            ambig_cond=0.0 if is_flat(math_mode) else AMBIG_COND_FAST)
    assert_forward_state_bit_exact(rast, st, sc.n)
    res = {}
    if st.n_rendered:
        res.update(assert_image_close(img, st, ref_img, strict=is_flat(math_mode)))
        res["ncontrib_mismatch"] = assert_ncontrib(rast, st)
    else:
        assert (np_(img) == 0).all()
    if check_backward:
        vp = make_vpixels(sc.width, sc.height, CH[mode], vpix_seed)
        g = gpu_backward(rast, dev, cam, sc.sh_degree, torch.from_numpy(vp).cuda(), background)
        ref = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode=mode,
                         sh_degree=sc.sh_degree, background=background)
        res.update(assert_grads_close(g, ref, rtol=grad_rtol, ambig_g=st.ambiguous_g))
        ref2 = dict(gm=ref["vmeans2d"])
        res["grad_means2d"] = assert_grads_close(dict(gm=np_(rast.gstate.grad_means2d)), ref2, rtol=grad_rtol,
                                                 keys=("gm",), ambig_g=st.ambiguous_g)["gm"]
    return res, rast, st
