"""Debug helper (GPU): C2 gradients of the default (strict) mode vs the fp32 and fp64 oracles over several backward runs —
how much of the error is the GPU's atomic-order noise (run-to-run spread), how much is common to all runs, which rows
carry it and how the fp32 ORACLE itself sits against the fp64 one on those rows."""
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
ref_img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode="rgbd", sh_degree=3, ambig_rel=P.AMBIG_REL)
ref = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode="rgbd", sh_degree=3)
o64 = P.oracle(np.float64)
_, st64 = o64.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode="rgbd", sh_degree=3, ambig_rel=P.AMBIG_REL)
ref64 = o64.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st64, mode="rgbd", sh_degree=3)
keep = (st.ambiguous_g == 0) & (st64.ambiguous_g == 0)
MM = os.environ.get("MATH_MODE", "strict")
MM = int(MM) if MM.isdigit() else MM
print("math_mode", MM)
rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd", math_mode=MM)
P.gpu_forward(rast, dev, cam, 3)
runs = []
NRUNS = int(os.environ.get('NRUNS', '6'))
for r in range(NRUNS):
    g = P.gpu_backward(rast, dev, cam, 3, torch.from_numpy(vp).cuda())
    runs.append({k: P.np_(g[k]).reshape(ref[k].shape).astype(np.float64).copy() for k in ("vopacities", "vmeans", "vscales", "vrot", "vshs")})
for k in ("vopacities", "vmeans", "vscales", "vrot", "vshs"):
    scale = np.abs(ref[k]).max()
    rowerr = lambda a, b: np.where(keep, np.abs(a - b).reshape(a.shape[0], -1).max(1) / scale, 0.0)
    e32 = [rowerr(r[k], ref[k]) for r in runs]
    e64 = [rowerr(r[k], ref64[k]) for r in runs]
    eo = rowerr(ref[k].astype(np.float64), ref64[k])
    spread = rowerr(runs[0][k], runs[1][k])
    w = int(np.argmax(e32[0]))
    print(k, "max err vs fp32 oracle per run", [f"{e.max():.2e}" for e in e32], "| vs fp64 oracle", [f"{e.max():.2e}" for e in e64],
          "| fp32 oracle vs fp64 oracle", f"{eo.max():.2e}", "| run0 vs run1", f"{spread.max():.2e}",
          "| rows over 5e-5 (run 0)", int((e32[0] > 5e-5).sum()), "| worst row", w, "its err vs fp64", f"{e64[0][w]:.2e}", "oracle32 vs 64 there", f"{eo[w]:.2e}",
          "scales", sc.scales[w], "radius", int(st.radii[w]))

# tail of the per-run maximum (atomic-order noise): many runs, vrot and vscales only
if NRUNS > 6:
    for k in ("vrot", "vscales", "vopacities"):
        scale = np.abs(ref[k]).max()
        mx = sorted(float(np.where(keep, np.abs(r[k] - ref[k]).reshape(r[k].shape[0], -1).max(1) / scale, 0.0).max()) for r in runs)
        print(k, "per-run max over", NRUNS, "runs: min %.2e median %.2e p90 %.2e max %.2e" % (mx[0], mx[len(mx) // 2], mx[int(0.9 * len(mx))], mx[-1]))
