"""A training-shaped step on the C2 scene from RAW parameters: functor forward -> photometric loss -> backward to
raw-parameter gradients.  Compares (a) the reference's composition (torch sigmoid / exp / cat broadcasts + `rasterize`
+ torch L1/permute + fused_ssim, pullbacks chained by autograd) with (b) the fused path (gsr_forward_raw,
gsr_photometric_loss, gsr_backward_raw).  CUDA events around K steps; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussiansplatting.jl_b200"))
from gsrast import Camera, GaussianRasterizer, ssim  # noqa: E402
from gsrast.synthetic import make_config  # noqa: E402


def main():
    sc = make_config("C2")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    raw = dict(means=t(sc.means), opac=t(np.log(sc.opacities / (1 - sc.opacities)).reshape(-1, 1).clip(-12, 12).astype(np.float32)),
               scales=t(np.log(sc.scales).astype(np.float32)), rots=t(sc.rotations), dc=t(sc.shs[:, :1]), rest=t(sc.shs[:, 1:]))
    cam = Camera(fx=sc.fx, fy=sc.fy, width=sc.width, height=sc.height)
    rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd")
    target = torch.rand((3, sc.height, sc.width), device="cuda")
    lam = 0.2

    def composed():
        leaves = {k: v.detach().requires_grad_(True) for k, v in raw.items()}
        img = rast(leaves["means"], leaves["opac"], leaves["scales"], leaves["rots"], leaves["dc"], leaves["rest"], camera=cam,
                   sh_degree=3, fused_activations=False)
        x = img[:, :, :3].permute(2, 0, 1).unsqueeze(0).contiguous()           # training.jl:684-686
        tt = target.unsqueeze(0)
        loss = (1 - lam) * (x - tt).abs().mean() + lam * (1 - ssim.fused_ssim(x, tt).mean())
        loss.backward()
        return loss

    def fused():
        img = rast._raw_call(False, raw["means"], raw["opac"], raw["scales"], raw["rots"], raw["dc"], raw["rest"], None, None,
                             cam, 3, (0, 0, 0), image=rast.image)
        loss, vpix = ssim.photometric_loss(rast, img, target, lam)
        rast._raw_call(True, raw["means"], raw["opac"], raw["scales"], raw["rots"], raw["dc"], raw["rest"], None, None, cam, 3,
                       (0, 0, 0), vpixels=vpix, outs=outs)
        return loss

    n, K = sc.n, 16
    outs = dict(vmeans=torch.empty((n, 3), device="cuda"), vfeatures_dc=torch.empty((n, 1, 3), device="cuda"),
                vfeatures_rest=torch.empty((n, K - 1, 3), device="cuda"), vopacities=torch.empty((n, 1), device="cuda"),
                vscales=torch.empty((n, 3), device="cuda"), vrot=torch.empty((n, 4), device="cuda"))
    res = {}
    for name, fn in (("composed", composed), ("fused", fused)):
        for _ in range(5):
            l = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 30
        e0.record()
        for _ in range(steps):
            l = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms_per_step": round(e0.elapsed_time(e1) / steps, 4), "loss": float(l if l.dim() == 0 else l[0])}
    print(json.dumps({"workload": "C2 from raw parameters: functor fwd + L1/D-SSIM loss + bwd to raw-parameter gradients", **res}))


if __name__ == "__main__":
    main()
