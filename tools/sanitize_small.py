"""Small forward+backward of every mode / math mode for compute-sanitizer (memcheck, racecheck, initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
import numpy as np, torch
from gsrast import Camera, GaussianRasterizer, update_stats
from gsrast.synthetic import make_scene, make_vpixels

for mode, C, deg in (("rgb", 3, 0), ("rgbd", 5, 3), ("rgbdn", 8, 2)):
    for mm in ("reference", "fast"):
        sc = make_scene(3000, deg, 160, 128, 7)
        cam = Camera(fx=sc.fx, fy=sc.fy, width=160, height=128)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        means, shs, opac, scales, rots = d(sc.means), d(sc.shs), d(sc.opacities.reshape(-1, 1)), d(sc.scales), d(sc.rotations)
        rast = GaussianRasterizer(width=160, height=128, mode=mode, math_mode=mm)
        covis = torch.zeros(sc.n, dtype=torch.uint8, device="cuda")
        unc = torch.zeros((128, 160), device="cuda")
        rast._forward(means, shs, opac, scales, rots, None, None, cam, deg, (0.1, 0.2, 0.3), covis, unc)
        g = rast._backward(d(make_vpixels(160, 128, C, 3)), means, shs, opac, scales, rots, None, None, cam, deg, (0.1, 0.2, 0.3))
        mr = torch.zeros(sc.n, dtype=torch.int32, device="cuda"); acc = torch.zeros(sc.n, device="cuda"); den = torch.zeros(sc.n, device="cuda")
        update_stats(mr, acc, den, rast)
        torch.cuda.synchronize()
        print(mode, mm, "M", rast.n_rendered, float(g["vmeans"].abs().sum()))
print("done")
