"""`densify_and_prune!` through libgsrast's densification kernels against the NumPy restatement of
src/densification.jl (oracle/densify_ref.py).  Byte movement (selection, repetition, order, Adam moments, ids,
statistics) must be exact; the split children's positions / log-scales are float32 arithmetic (1e-6 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PARAMS = ("points", "features_dc", "features_rest", "scales", "rotations", "opacities")


def make(n, R, iso, seed):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
    model = dict(points=f(n, 3), features_dc=f(n, 1, 3), features_rest=f(n, R, 3),
                 scales=(rng.normal(-3.0, 1.0, (n, 1 if iso else 3))).astype(np.float32), rotations=f(n, 4),
                 opacities=rng.normal(0, 3, (n, 1)).astype(np.float32), ids=np.arange(n, dtype=np.int32))
    opt = {k: (f(*model[k].shape), np.abs(f(*model[k].shape))) for k in PARAMS}
    denom = rng.integers(0, 4, n).astype(np.float32)                    # zeros -> 0/0 = NaN -> 0
    accum = (rng.random(n).astype(np.float32) * 4e-4 * np.maximum(denom, 1)) * (denom > 0)
    stats = dict(max_radii=rng.integers(0, 40, n).astype(np.int32), accum=accum.astype(np.float32), denom=denom)
    return model, opt, stats


def keep_away_from_thresholds(model, stats, thr, gamma):
    """exp() may differ by an ulp between libdevice and NumPy: move values off the decision boundaries."""
    with np.errstate(divide="ignore", invalid="ignore"):
        g = stats["accum"] / stats["denom"]
    near = np.abs(g - thr) < 1e-3 * thr
    stats["accum"][near] *= 1.01
    s = np.exp(model["scales"]).max(1)
    near = np.abs(s - gamma) < 1e-3 * gamma
    model["scales"][near] -= 0.01
    s = np.exp(model["scales"]).max(1)
    near = np.abs(s - 0.1 * 4.0) < 1e-3
    model["scales"][near] -= 0.01
    o = 1 / (1 + np.exp(-model["opacities"]))
    near = np.abs(o - 0.005) < 1e-5
    model["opacities"][near] += 0.1


@pytest.mark.parametrize("n,R,iso,max_screen", [(20_000, 15, False, 0), (5_000, 0, True, 20), (70_000, 3, False, 20), (300, 8, False, 0)])
def test_densify_and_prune_matches_oracle(n, R, iso, max_screen):
    from gsrast import densify
    from oracle import densify_ref
    kw = dict(grad_threshold=2e-4, dense_percent=0.01, extent=4.0, pruning_extent=4.0, max_screen_size=max_screen,
              min_opacity=0.005)
    model, opt, stats = make(n, R, iso, n + R)
    keep_away_from_thresholds(model, stats, kw["grad_threshold"], np.float32(4.0) * np.float32(0.01))
    noise = np.random.default_rng(1).normal(0, 1, (2 * n, 3)).astype(np.float32)
    m_o, o_o, s_o, info_o = densify_ref.densify_and_prune(model, opt, stats, noise=noise, **kw)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m_g, o_g, s_g, info_g = densify.densify_and_prune({k: t(v) for k, v in model.items()}, {k: (t(a), t(b)) for k, (a, b) in opt.items()},
                                                      {k: t(v) for k, v in stats.items()}, noise=t(noise), **kw)
    torch.cuda.synchronize()
    assert info_g == info_o and info_o["n_clone"] > 0 and info_o["n_split"] > 0 and info_o["n_pruned"] > 0
    n_out = m_o["points"].shape[0]
    n_children = 2 * info_o["n_split"]
    assert np.array_equal(m_g["ids"].cpu().numpy(), m_o["ids"])          # selection, repetition and order
    for k in PARAMS:
        got, want = m_g[k].cpu().numpy(), m_o[k]
        assert got.shape == want.shape == (n_out,) + model[k].shape[1:], k
        if k in ("points", "scales"):  # children carry float arithmetic; everything else is moved bytes
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6, err_msg=k)
        else:
            assert np.array_equal(got, want), k
        for j in (0, 1):
            assert np.array_equal(o_g[k][j].cpu().numpy(), o_o[k][j]), (k, j)
    for k in ("max_radii", "accum", "denom"):
        assert np.array_equal(s_g[k].cpu().numpy(), s_o[k]), k
    assert n_children > 0


def test_mask_offsets_and_gather_edge_cases():
    from gsrast import densify
    for n in (0, 1, 4095, 4096, 4097, 1_200_000):
        rng = np.random.default_rng(n)
        mask = rng.random(n) < 0.3
        md = torch.from_numpy(mask).cuda()
        offs, cnt = densify.mask_offsets(md)
        assert cnt == int(mask.sum())
        if n:
            want = np.cumsum(mask) - mask
            assert np.array_equal(offs[:n].cpu().numpy(), want.astype(np.int32))
            x = torch.arange(n * 3, dtype=torch.float32, device="cuda").reshape(n, 3)
            assert torch.equal(densify.select_rows(x, md, offs, cnt), x[md])
            assert torch.equal(densify.select_rows(x, md, offs, cnt, repeat=2), torch.cat([x[md], x[md]]))
    none = torch.zeros(1000, dtype=torch.bool, device="cuda")
    offs, cnt = densify.mask_offsets(none)
    assert cnt == 0 and densify.select_rows(torch.ones((1000, 4), device="cuda"), none, offs, cnt).shape == (0, 4)
