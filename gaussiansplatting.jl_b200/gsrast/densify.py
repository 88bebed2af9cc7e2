"""Host mirror of `densify_and_prune!` (src/densification.jl) over libgsrast's densification kernels (csrc/densify.cu).

The sequence is the reference's — clone, split (+ noise), prune what was split, final prune — with Adam moments and
the densification statistics following their parameters.  Every mask, prefix sum, row gather and the split transform
is a library launch; torch only allocates and concatenates device buffers.  Tensors use the package's layout: points
(N,3), features_dc (N,1,3), features_rest (N,R,3), scales (N,3) or (N,1) isotropic, rotations (N,4), opacities (N,1)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

PARAMS = ("points", "features_dc", "features_rest", "scales", "rotations", "opacities")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def mask_offsets(mask: torch.Tensor):
    """Exclusive prefix of a bool mask and its population count (one host read)."""
    n = mask.numel()
    m8 = mask.view(torch.uint8)
    offs = torch.empty(max(n, 1), dtype=torch.int32, device=mask.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=mask.device)
    scratch = torch.empty(int(_lib.lib().gsr_mask_offsets_scratch_words(n)), dtype=torch.int32, device=mask.device)
    _lib.check(_lib.lib().gsr_mask_offsets(n, _p(m8), _p(offs), _p(cnt), _p(scratch), _stream(mask.device)))
    return offs, int(cnt.item())


def select_rows(x: torch.Tensor, mask: torch.Tensor, offs: torch.Tensor, count: int, repeat: int = 1) -> torch.Tensor:
    """`x[mask]` repeated `repeat` times block-wise (`repeat(x[:, mask], 1, r)` in the reference's layout)."""
    assert x.is_contiguous() and x.element_size() == 4
    n = x.shape[0]
    row = x[0].numel() if n else int(torch.tensor(x.shape[1:]).prod()) if x.dim() > 1 else 1
    out = torch.empty((count * repeat,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().gsr_gather_rows(n, 4 * row, _p(x), _p(mask.view(torch.uint8)), _p(offs), _p(out), repeat, count,
                                          _stream(x.device)))
    return out


def _append(model, opt, new):
    for k in PARAMS:
        model[k] = torch.cat([model[k], new[k]], 0)
        mu, nu = opt[k]
        z = torch.zeros_like(new[k])
        opt[k] = (torch.cat([mu, z], 0), torch.cat([nu, z], 0))
    if model.get("ids") is not None:
        model["ids"] = torch.cat([model["ids"], new["ids"]], 0)
    n, dev = model["points"].shape[0], model["points"].device
    return dict(max_radii=torch.zeros(n, dtype=torch.int32, device=dev), accum=torch.zeros(n, device=dev),
                denom=torch.zeros(n, device=dev))


def _prune(model, opt, stats, valid):
    offs, cnt = mask_offsets(valid)
    sel = lambda x: select_rows(x, valid, offs, cnt)
    for k in PARAMS:
        model[k] = sel(model[k])
        opt[k] = (sel(opt[k][0]), sel(opt[k][1]))
    if model.get("ids") is not None:
        model["ids"] = sel(model["ids"])
    return {k: sel(v) for k, v in stats.items()}


def densify_and_prune(model: dict, opt: dict, stats: dict, *, grad_threshold: float, dense_percent: float, extent: float,
                      pruning_extent: float, max_screen_size: int, min_opacity: float, noise: torch.Tensor, n_split: int = 2):
    """Returns (model, opt, stats, info); info = counts of cloned / split / finally pruned Gaussians."""
    lib = _lib.lib()
    model, opt = dict(model), dict(opt)
    dev = model["points"].device
    st = _stream(dev)
    iso = int(model["scales"].shape[1] == 1)
    n_grad = model["points"].shape[0]
    accum, denom = stats["accum"].contiguous(), stats["denom"].contiguous()
    gamma = float(torch.tensor(extent, dtype=torch.float32) * torch.tensor(dense_percent, dtype=torch.float32))

    # ---- clone (densification.jl:28-60) ----
    n = n_grad
    clone = torch.empty(n, dtype=torch.bool, device=dev)
    _lib.check(lib.gsr_densify_masks(n, n_grad, _p(accum), _p(denom), _p(model["scales"]), iso, grad_threshold, gamma,
                                     _p(clone.view(torch.uint8)), None, st))
    offs, cnt = mask_offsets(clone)
    new = {k: select_rows(model[k], clone, offs, cnt) for k in PARAMS}
    new["ids"] = None if model.get("ids") is None else select_rows(model["ids"], clone, offs, cnt)
    stats = _append(model, opt, new)
    info = dict(n_clone=cnt)

    # ---- split (:62-121) ----
    n = model["points"].shape[0]
    split = torch.empty(n, dtype=torch.bool, device=dev)
    _lib.check(lib.gsr_densify_masks(n, n_grad, _p(accum), _p(denom), _p(model["scales"]), iso, grad_threshold, gamma, None,
                                     _p(split.view(torch.uint8)), st))
    offs, cnt = mask_offsets(split)
    new = {k: select_rows(model[k], split, offs, cnt, repeat=n_split) for k in PARAMS}
    new["ids"] = None if model.get("ids") is None else select_rows(model["ids"], split, offs, cnt, repeat=n_split)
    m = cnt * n_split
    if m:
        nz = noise[:m].contiguous()
        _lib.check(lib.gsr_split_children(m, _p(new["points"]), _p(new["scales"]), iso, _p(new["rotations"]), _p(nz), n_split, st))
    stats = _append(model, opt, new)
    valid = torch.cat([~split, torch.ones(m, dtype=torch.bool, device=dev)])
    stats = _prune(model, opt, stats, valid)
    info["n_split"] = cnt

    # ---- final prune (:17-26) ----
    n = model["points"].shape[0]
    valid = torch.empty(n, dtype=torch.bool, device=dev)
    g2 = float(torch.tensor(0.1, dtype=torch.float32) * torch.tensor(pruning_extent, dtype=torch.float32))
    _lib.check(lib.gsr_prune_mask(n, _p(model["opacities"]), _p(model["scales"]), iso, _p(stats["max_radii"]), min_opacity,
                                  int(max_screen_size), g2, _p(valid.view(torch.uint8)), st))
    n_before = n
    stats = _prune(model, opt, stats, valid)
    info["n_pruned"] = n_before - model["points"].shape[0]
    return model, opt, stats, info
