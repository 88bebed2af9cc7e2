"""NumPy restatement of `densify_and_prune!` (src/densification.jl) — TEST INFRASTRUCTURE ONLY (see gsr_oracle.c).

Follows the reference statement by statement: clone (:28-60), split (:62-121) with its noise kernel (:123-136), the
final prune (:17-26), `prune_points!` (:138-196), `densification_postfix!` / `append_gaussians!` (:198-264) and the
Adam-moment bookkeeping (`_append_optimizer!` :266-281, `_prune_optimizer!` :283-292).  Arrays use this repository's
layout, byte-identical to Julia's: points (N,3), features_dc (N,1,3), features_rest (N,R,3), scales (N,3) or (N,1)
isotropic, rotations (N,4) wxyz, opacities (N,1); every optimizer is a (mu, nu) pair shaped like its parameter.
The split's N(0,1) deviates are an argument (`noise`, (n_children,3)): the reference draws them on the device."""
from __future__ import annotations

import numpy as np

PARAMS = ("points", "features_dc", "features_rest", "scales", "rotations", "opacities")
f32 = np.float32


def _sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x.astype(f32)))).astype(f32)


def _max_exp_scale(scales):
    return np.exp(scales.astype(f32)).max(axis=1)


def _quat2rot(q):
    """unnorm_quat2rot (render.jl:322-333) for (m,4) wxyz -> (m,3,3) with R[:, i, j] = row i, col j."""
    q = q.astype(f32)
    qi = f32(1) / np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    w, x, y, z = (qi * q[:, k] for k in range(4))
    R = np.empty((len(q), 3, 3), f32)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 1, 0] = 2 * (x * y + w * z); R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 0, 1] = 2 * (x * y - w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 0, 2] = 2 * (x * z + w * y); R[:, 1, 2] = 2 * (y * z - w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _append(model, opt, new):
    """append_gaussians! + _append_optimizer! + the statistics reset of densification_postfix!."""
    for k in PARAMS:
        model[k] = np.concatenate([model[k], new[k]], 0)
        mu, nu = opt[k]
        opt[k] = (np.concatenate([mu, np.zeros_like(new[k])], 0), np.concatenate([nu, np.zeros_like(new[k])], 0))
    if model.get("ids") is not None:
        model["ids"] = np.concatenate([model["ids"], new["ids"]], 0)
    n = len(model["points"])
    return dict(max_radii=np.zeros(n, np.int32), accum=np.zeros(n, f32), denom=np.zeros(n, f32))


def _prune(model, opt, stats, valid):
    """prune_points! + _prune_optimizer!."""
    for k in PARAMS:
        model[k] = model[k][valid]
        opt[k] = (opt[k][0][valid], opt[k][1][valid])
    if model.get("ids") is not None:
        model["ids"] = model["ids"][valid]
    return {k: v[valid] for k, v in stats.items()}


def densify_and_prune(model, opt, stats, *, grad_threshold, dense_percent, extent, pruning_extent, max_screen_size,
                      min_opacity, noise, n_split=2):
    model = {k: (None if v is None else np.array(v)) for k, v in model.items()}
    opt = {k: (np.array(a), np.array(b)) for k, (a, b) in opt.items()}
    stats = {k: np.array(v) for k, v in stats.items()}
    with np.errstate(divide="ignore", invalid="ignore"):
        grad = (stats["accum"].astype(f32) / stats["denom"].astype(f32)).astype(f32)  # :8
    grad[np.isnan(grad)] = 0                                                          # :9-10
    gamma = f32(extent) * f32(dense_percent)

    # ---- densify_clone! ------------------------------------------------------------------------------------
    mask = (grad > f32(grad_threshold)) & (_max_exp_scale(model["scales"]) < gamma)
    new = {k: model[k][mask] for k in PARAMS}
    new["ids"] = None if model.get("ids") is None else model["ids"][mask]
    stats = _append(model, opt, new)
    info = dict(n_clone=int(mask.sum()))

    # ---- densify_split! ------------------------------------------------------------------------------------
    n = len(model["points"])
    padded = np.zeros(n, f32)
    padded[: len(grad)] = grad                                                        # :74-75
    mask = (padded >= f32(grad_threshold)) & (_max_exp_scale(model["scales"]) > gamma)
    rep = lambda a: np.concatenate([a[mask]] * n_split, 0)                            # repeat(x[:, mask], 1, n_split)
    stds = rep(np.exp(model["scales"].astype(f32)))
    new = {k: rep(model[k]) for k in PARAMS}
    new["scales"] = np.log(stds / (f32(0.8) * f32(n_split))).astype(f32)             # :91
    m = len(new["points"])
    if m:
        xi = (stds * np.asarray(noise, f32)[:m]).astype(f32)                          # sigma .* randn (isotropic broadcasts)
        R = _quat2rot(new["rotations"])
        new["points"] = (new["points"] + ((R[:, :, 0] * xi[:, :1] + R[:, :, 1] * xi[:, 1:2]) + R[:, :, 2] * xi[:, 2:3])).astype(f32)
    new["ids"] = None if model.get("ids") is None else rep(model["ids"])
    stats = _append(model, opt, new)
    valid = np.concatenate([~mask, np.ones(m, bool)])                                 # :116
    stats = _prune(model, opt, stats, valid)
    info["n_split"] = int(mask.sum())

    # ---- final prune (:17-26) --------------------------------------------------------------------------------
    valid = _sigmoid(model["opacities"]).reshape(-1) > f32(min_opacity)
    if max_screen_size > 0:
        g2 = f32(0.1) * f32(pruning_extent)
        valid &= (stats["max_radii"] < max_screen_size) & (_max_exp_scale(model["scales"]) < g2)
    stats = _prune(model, opt, stats, valid)
    info["n_pruned"] = int((~valid).sum())
    return model, opt, stats, info
