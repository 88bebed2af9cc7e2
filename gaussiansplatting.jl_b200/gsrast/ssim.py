"""Host mirror of the reference's fused SSIM operator (src/fused_ssim.jl) and of the photometric loss that sits
between the rasterizer's image and its cotangent (src/training.jl:684-699), over libgsrast's C ABI.

Layouts: `fused_ssim` takes torch tensors of shape (B, CH, H, W), contiguous — byte-identical to the reference's
column-major (W, H, CH, B).  `photometric_loss` takes the rasterizer's own (H, W, C) image (the reference's
(C, W, H)) and a (3, H, W) target, and returns the loss terms plus the (H, W, C) cotangent `gsr_backward` consumes.
No torch kernels on these paths; without the CUDA library they raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

C1_DEFAULT = float(torch.tensor(0.01, dtype=torch.float32) * torch.tensor(0.01, dtype=torch.float32))  # 0.01f0^2
C2_DEFAULT = float(torch.tensor(0.03, dtype=torch.float32) * torch.tensor(0.03, dtype=torch.float32))  # 0.03f0^2


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check4(x, name):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous float32 CUDA tensor of shape (B, CH, H, W)")


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ssim_forward(img, ref, C1=C1_DEFAULT, C2=C2_DEFAULT, train=True):
    """`_fused_ssim(img; ref, C1, C2, train)` (fused_ssim.jl:354-372): (ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)."""
    _check4(img, "img")
    _check4(ref, "ref")
    if ref.shape != img.shape:
        raise ValueError("img and ref must have the same shape")
    B, CH, H, W = img.shape
    out = torch.empty_like(img)
    d = [torch.empty_like(img) for _ in range(3)] if train else [None, None, None]
    _lib.check(_lib.lib().gsr_ssim_forward(W, H, CH, B, _ptr(img), _ptr(ref), C1, C2, 1 if train else 0, _ptr(out),
                                           _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _stream(img.device)))
    return out, d[0], d[1], d[2]


def ssim_backward(img, ref, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """`fused_ssim_bwd` (fused_ssim.jl:374-389)."""
    for t, n in ((img, "img"), (ref, "ref"), (dL_dmap, "dL_dmap"), (dm_dmu1, "dm_dmu1"),
                 (dm_dsigma1_sq, "dm_dsigma1_sq"), (dm_dsigma12, "dm_dsigma12")):
        _check4(t, n)
    B, CH, H, W = img.shape
    out = torch.empty_like(img)
    _lib.check(_lib.lib().gsr_ssim_backward(W, H, CH, B, _ptr(img), _ptr(ref), _ptr(dL_dmap), _ptr(dm_dmu1),
                                            _ptr(dm_dsigma1_sq), _ptr(dm_dsigma12), _ptr(out), _stream(img.device)))
    return out


class _FusedSSIM(torch.autograd.Function):
    """The rrule of `_fused_ssim` (fused_ssim.jl:393-407)."""

    @staticmethod
    def forward(ctx, img, ref, C1, C2):
        train = img.requires_grad
        m, d0, d1, d2 = ssim_forward(img.detach(), ref, C1, C2, train=train)
        if train:
            ctx.save_for_backward(img.detach(), ref, d0, d1, d2)
        return m

    @staticmethod
    def backward(ctx, delta):
        img, ref, d0, d1, d2 = ctx.saved_tensors
        return ssim_backward(img, ref, delta.contiguous(), d0, d1, d2), None, None, None


def fused_ssim(img, ref, C1=C1_DEFAULT, C2=C2_DEFAULT):
    """`fused_ssim(img; ref, C1, C2)` (fused_ssim.jl:391-395): the SSIM map, differentiable w.r.t. img."""
    return _FusedSSIM.apply(img, ref, C1, C2)


def photometric_loss(rast, image, target, lambda_dssim=0.2):
    """total = (1-l)*mean|x - t| + l*(1 - mean ssim) on the rgb channels of the rasterizer's image, and its cotangent.

    rast: the GaussianRasterizer that produced `image` ((H, W, C)); target: (3, H, W).
    Returns (loss, vpixels): loss = float32 CUDA tensor {total, l1, mean ssim} (no sync), vpixels (H, W, C)."""
    H, W, Cc = rast.height, rast.width, rast.channels
    if tuple(image.shape) != (H, W, Cc) or not image.is_contiguous() or image.dtype != torch.float32:
        raise ValueError(f"image must be a contiguous float32 ({H}, {W}, {Cc}) tensor")
    if tuple(target.shape) != (3, H, W) or not target.is_contiguous() or target.dtype != torch.float32:
        raise ValueError(f"target must be a contiguous float32 (3, {H}, {W}) tensor")
    vpix = torch.empty_like(image)
    loss = torch.empty(3, dtype=torch.float32, device=image.device)
    _lib.check(_lib.lib().gsr_photometric_loss(rast._h, _ptr(image), _ptr(target), float(lambda_dssim), _ptr(vpix),
                                               _ptr(loss), _stream(image.device)), rast._h)
    return loss, vpix
