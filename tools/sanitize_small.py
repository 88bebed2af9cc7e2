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

# raw-parameter path, SSIM operator and photometric loss
from gsrast import ssim
sc = make_scene(2000, 3, 160, 128, 9)
cam = Camera(fx=sc.fx, fy=sc.fy, width=160, height=128)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rast = GaussianRasterizer(width=160, height=128, mode="rgbd")
raw = dict(means=d(sc.means), opac=d(np.log(sc.opacities / (1 - sc.opacities)).reshape(-1, 1).clip(-12, 12).astype(np.float32)),
           scales=d(np.log(sc.scales).astype(np.float32)), rots=d(sc.rotations), dc=d(sc.shs[:, :1]), rest=d(sc.shs[:, 1:]))
for iso in (False, True):
    scl = raw["scales"][:, :1].contiguous() if iso else raw["scales"]
    img = rast._raw_call(False, raw["means"], raw["opac"], scl, raw["rots"], raw["dc"], raw["rest"], None, None, cam, 3, (0, 0, 0),
                         image=rast.image)
    loss, vpix = ssim.photometric_loss(rast, img, torch.rand((3, 128, 160), device="cuda"), 0.2)
    g = rast._raw_call(True, raw["means"], raw["opac"], scl, raw["rots"], raw["dc"], raw["rest"], None, None, cam, 3, (0, 0, 0),
                       vpixels=vpix)
    torch.cuda.synchronize()
    print("raw iso", iso, float(loss[0]), float(g["vscales"].abs().sum()))
x, t = torch.rand((2, 3, 37, 53), device="cuda"), torch.rand((2, 3, 37, 53), device="cuda")
m, d0, d1, d2 = ssim.ssim_forward(x, t, train=True)
gg = ssim.ssim_backward(x, t, torch.rand_like(x), d0, d1, d2)
torch.cuda.synchronize()
print("ssim", float(m.mean()), float(gg.abs().sum()))
print("done 2")

# densification kernels
from gsrast import densify
n = 3000
rng = np.random.default_rng(2)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
model = dict(points=t(f(n, 3)), features_dc=t(f(n, 1, 3)), features_rest=t(f(n, 15, 3)), scales=t(rng.normal(-3, 1, (n, 3)).astype(np.float32)),
             rotations=t(f(n, 4)), opacities=t(rng.normal(0, 3, (n, 1)).astype(np.float32)), ids=t(np.arange(n, dtype=np.int32)))
opt = {k: (torch.randn_like(model[k]), torch.rand_like(model[k])) for k in densify.PARAMS}
den = t(rng.integers(0, 4, n).astype(np.float32))
stats = dict(max_radii=t(rng.integers(0, 40, n).astype(np.int32)), accum=t((rng.random(n) * 8e-4).astype(np.float32)) * (den > 0), denom=den)
m, o, s, info = densify.densify_and_prune(model, opt, stats, grad_threshold=2e-4, dense_percent=0.01, extent=4.0, pruning_extent=4.0,
                                          max_screen_size=20, min_opacity=0.005, noise=torch.randn((2 * n, 3), device="cuda"))
torch.cuda.synchronize()
print("densify", info, tuple(m["points"].shape))
print("done 3")
