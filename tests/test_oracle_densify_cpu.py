"""Invariants of the NumPy restatement of `densify_and_prune!` (oracle/densify_ref.py, src/densification.jl) that
follow from the reference's statements themselves: sizes, order, what is copied, what is zeroed."""
import numpy as np

from oracle import densify_ref

PARAMS = densify_ref.PARAMS


def _make(n=4000, R=3, seed=3):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
    model = dict(points=f(n, 3), features_dc=f(n, 1, 3), features_rest=f(n, R, 3), scales=rng.normal(-3, 1, (n, 3)).astype(np.float32),
                 rotations=f(n, 4), opacities=rng.normal(0, 3, (n, 1)).astype(np.float32), ids=np.arange(n, dtype=np.int32))
    opt = {k: (f(*model[k].shape), np.abs(f(*model[k].shape))) for k in PARAMS}
    denom = rng.integers(0, 4, n).astype(np.float32)
    stats = dict(max_radii=rng.integers(0, 40, n).astype(np.int32), accum=(rng.random(n) * 8e-4 * (denom > 0)).astype(np.float32), denom=denom)
    return model, opt, stats


def test_sequence_invariants():
    model, opt, stats = _make()
    n = len(model["points"])
    kw = dict(grad_threshold=2e-4, dense_percent=0.01, extent=4.0, pruning_extent=4.0, max_screen_size=0, min_opacity=0.0)
    noise = np.zeros((2 * n, 3), np.float32)  # zero noise: children sit on their parent
    m, o, s, info = densify_ref.densify_and_prune(model, opt, stats, noise=noise, **kw)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.nan_to_num(stats["accum"] / stats["denom"], nan=0.0)
    big = np.exp(model["scales"]).max(1) > np.float32(0.04)
    clone, split = (g > 2e-4) & ~big, (g >= 2e-4) & big
    assert info == dict(n_clone=int(clone.sum()), n_split=int(split.sum()), n_pruned=0)  # min_opacity 0: sigmoid > 0 always
    # order: surviving originals, clones, then the two blocks of children (densification.jl:57-60,103-117)
    want_ids = np.concatenate([np.arange(n)[~split], np.arange(n)[clone], np.arange(n)[split], np.arange(n)[split]])
    assert np.array_equal(m["ids"], want_ids)
    k0 = int((~split).sum())
    assert np.array_equal(m["points"][:k0], model["points"][~split])
    assert np.array_equal(m["points"][k0:k0 + info["n_clone"]], model["points"][clone])          # clones are exact copies
    ch = m["points"][k0 + info["n_clone"]:]
    assert np.array_equal(ch, np.concatenate([model["points"][split]] * 2))                        # zero noise
    np.testing.assert_allclose(np.exp(m["scales"][k0 + info["n_clone"]:]), np.concatenate([np.exp(model["scales"][split])] * 2) / 1.6, rtol=1e-6)
    # Adam moments: kept rows keep theirs, every appended row starts from zero (_append_optimizer!)
    for k in PARAMS:
        assert np.array_equal(o[k][0][:k0], opt[k][0][~split]) and not o[k][0][k0:].any() and not o[k][1][k0:].any()
    # statistics are re-created as zeros by densification_postfix! (:214-217)
    assert not s["max_radii"].any() and not s["accum"].any() and not s["denom"].any() and len(s["accum"]) == len(m["points"])


def test_final_prune_and_isotropic():
    model, opt, stats = _make(seed=5)
    model["scales"] = model["scales"][:, :1].copy()
    opt["scales"] = (opt["scales"][0][:, :1].copy(), opt["scales"][1][:, :1].copy())
    noise = np.random.default_rng(0).normal(0, 1, (8000, 3)).astype(np.float32)
    m, o, s, info = densify_ref.densify_and_prune(model, opt, stats, grad_threshold=2e-4, dense_percent=0.01, extent=4.0,
                                                  pruning_extent=4.0, max_screen_size=20, min_opacity=0.3, noise=noise)
    assert m["scales"].shape[1] == 1 and info["n_pruned"] > 0
    assert (1 / (1 + np.exp(-m["opacities"])) > 0.3).all() and (np.exp(m["scales"]).max(1) < 0.4).all()
    assert all(len(o[k][0]) == len(m["points"]) == len(s["denom"]) for k in PARAMS)
