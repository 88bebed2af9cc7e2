"""Debug helper (GPU): C2 gradients in both math modes vs the oracle; prints the worst rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import parity as P
from gsrast import GaussianRasterizer
from gsrast.synthetic import make_config, make_vpixels

sc = make_config("C2")
cam, ocam = P.cameras(sc)
dev = P.to_dev(sc)
vp = make_vpixels(sc.width, sc.height, 5, 1002)
o = P.oracle()
ref_img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode="rgbd", sh_degree=3, ambig_rel=2e-4)
ref = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode="rgbd", sh_degree=3)
keep = st.ambiguous_g == 0
for mm in ("reference", "fast"):
    rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd", math_mode=mm)
    P.gpu_forward(rast, dev, cam, 3)
    nc = P.np_(rast.gstate.n_contrib).view(np.uint32)
    print(mm, "n_contrib mismatches:", int((nc != st.n_contrib).sum()), "on non-ambiguous:", int(((nc != st.n_contrib) & (st.ambiguous == 0)).sum()))
    g = P.gpu_backward(rast, dev, cam, 3, torch.from_numpy(vp).cuda())
    for k in ("vopacities", "vmeans", "vscales", "vrot", "vshs"):
        a = P.np_(g[k]).reshape(ref[k].shape).astype(np.float64)
        scale = np.abs(ref[k]).max()
        d = np.abs(a - ref[k]).reshape(a.shape[0], -1).max(1) / scale
        d[~keep] = 0
        w = np.argsort(d)[-3:][::-1]
        print(mm, k, "worst rows", [(int(i), float(d[i])) for i in w])
    a = P.np_(g["vopacities"]).reshape(-1)
    print(mm, "row 81569: gpu", a[81569], "oracle", ref["vopacities"][81569], "S_e/op check; opacity", sc.opacities[81569])
    acc = P.np_(rast.gstate.grad_means2d)[81569]
    print(mm, "grad_means2d gpu", acc, "oracle", ref["vmeans2d"][81569])
