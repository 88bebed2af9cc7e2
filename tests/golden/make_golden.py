"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is Julia-only and cannot run here, so these are NOT reference outputs: they freeze the
oracle's own results on small seeded scenes so that (a) the oracle cannot drift silently and (b) the
`-m gpu` parity tests have a committed target that does not need the oracle at all.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
from oracle.oracle import Oracle, OracleCamera  # noqa: E402
from gsrast.synthetic import make_scene, make_vpixels  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    o = Oracle(np.float32)
    n, deg, w, h, seed = 3000, 2, 128, 96, 4242
    sc = make_scene(n, deg, w, h, seed)
    cam = OracleCamera.simple(sc.fx, sc.fy, w, h)
    img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode="rgbd", sh_degree=deg,
                        ambig_rel=2e-5)
    vp = make_vpixels(w, h, 5, seed)
    g = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, st, mode="rgbd", sh_degree=deg)
    vis = st.radii > 0
    np.savez_compressed(
        os.path.join(HERE, "oracle_c1_small.npz"),
        n=n, deg=deg, w=w, h=h, seed=seed,
        radii=st.radii, keys_sorted=st.keys_sorted, values_sorted=st.values_sorted, ranges=st.ranges,
        n_contrib=st.n_contrib, means2d_vis_bits=st.means2d[vis].view(np.uint32),
        conics_vis_bits=st.conics[vis].view(np.uint32), rgbs_vis_bits=st.rgbs[vis].view(np.uint32),
        depths_vis_bits=st.depths[vis].view(np.uint32), image=img, accum_alpha=st.accum_alpha,
        ambiguous=st.ambiguous, vmeans=g["vmeans"], vshs=g["vshs"], vopacities=g["vopacities"],
        vscales=g["vscales"], vrot=g["vrot"], vmeans2d=g["vmeans2d"])
    print("wrote oracle_c1_small.npz: V=%d M=%d" % (vis.sum(), st.n_rendered))


if __name__ == "__main__":
    main()
