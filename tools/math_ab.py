"""A/B of the compositing arithmetic policies (csrc/render.cu, built with EXTRA_NVFLAGS=-DGSR_POLICY_AB) against the
fp32 CPU oracle with the FLAT north_star tolerances: which operations have to stay in the reference's order for the
image to stay within 1e-5 absolute and the gradients within 1e-4 relative, and what each relaxation buys in time.

    python tools/math_ab.py [C1d C5 C2 ...]      -> one JSON line per (config, policy) in gpurun_out/math_ab.jsonl

Test / measurement infrastructure: uses oracle/.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200"), os.path.join(ROOT, "tests")]
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity as P  # noqa: E402
from gsrast import GaussianRasterizer, _lib  # noqa: E402
from gsrast.synthetic import CONFIGS, make_config, make_vpixels  # noqa: E402


def policy(sig, expk, col, div, acc):
    return sig | (expk << 1) | (col << 3) | (div << 5) | (acc << 6)


MODES = [("reference", "reference"), ("strict", "strict"), ("fast", "fast"),
         ("strict,splitexp", 256 + policy(1, 2, 0, 0, 0)), ("strict,refdepth", 256 + policy(1, 1, 2, 0, 0)),
         ("strict,div", 256 + policy(1, 1, 0, 1, 0))]


def exp_probe(out):
    rng = np.random.default_rng(0)
    sig = np.concatenate([rng.uniform(0, 6, 1 << 24), rng.uniform(0, 90, 1 << 20), [0.0, 5.5412635, 1e-30, 87.0]]).astype(np.float32)
    s = torch.from_numpy(sig).cuda()
    a, b = torch.empty_like(s), torch.empty_like(s)
    _lib.check(_lib.lib().gsr_debug_exp_neg(C.c_void_p(s.data_ptr()), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), None, s.numel(), None))
    torch.cuda.synchronize()
    exact = np.exp(-sig.astype(np.float64))
    ulp = np.spacing(exact.astype(np.float32)).astype(np.float64)
    lo = sig <= 6
    res = {}
    for name, t in (("split", a), ("libdevice", b)):
        err = np.abs(t.cpu().numpy().astype(np.float64) - exact) / ulp
        res[name] = {"max_ulp_sigma<=6": float(err[lo].max()), "mean_ulp_sigma<=6": float(err[lo].mean()),
                     "max_ulp_sigma<=90": float(err[exact > 1e-37].max())}
    import math
    glibc = np.array([np.float32(math.exp(-float(x))) for x in sig[:200000]])  # correctly rounded stand-in for glibc expf
    res["split_equals_rounded_exp_frac"] = float((a.cpu().numpy()[:200000] == glibc).mean())
    res["libdevice_equals_rounded_exp_frac"] = float((b.cpu().numpy()[:200000] == glibc).mean())
    out.write(json.dumps({"exp_probe": res}) + "\n")
    out.flush()
    print("exp probe:", res)


def run_config(name, out):
    n, deg, W, H, mode, seed, _ = CONFIGS[name]
    sc = make_config(name)
    Cn = P.CH[mode]
    cam, ocam = P.cameras(sc)
    dev = P.to_dev(sc)
    o = P.oracle()
    ref_img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode=mode, sh_degree=deg,
                            ambig_rel=P.AMBIG_REL)
    vp = make_vpixels(W, H, Cn, seed)
    ref = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode=mode, sh_degree=deg)
    vpd = torch.from_numpy(vp).cuda()
    ok = st.ambiguous == 0
    keepg = st.ambiguous_g == 0
    for label, mm in MODES:
        try:
            rast = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=mm)
        except Exception as e:
            print(label, "unavailable:", e)
            continue
        img = P.np_(P.gpu_forward(rast, dev, cam, deg))
        g = P.gpu_backward(rast, dev, cam, deg, vpd)
        d = np.abs(img.astype(np.float64) - ref_img.astype(np.float64))
        rec = {"config": name, "policy": label, "ambiguous_px": int((~ok).sum()), "ambiguous_frac": float((~ok).mean()),
               "img_max_err_per_channel": [float(d[:, :, c][ok].max()) for c in range(Cn)],
               "img_max_err": float(d[ok].max()), "img_px_over_1e-5": int((d.max(2) > 1e-5)[ok].sum()),
               "img_max_err_ambiguous": float(d[~ok].max()) if (~ok).any() else 0.0,
               "ncontrib_mismatch_nonambig": int((P.np_(rast.gstate.n_contrib).view(np.uint32) != st.n_contrib)[ok].sum())}
        gr = {}
        for k in ("vmeans", "vshs", "vopacities", "vscales", "vrot"):
            a = P.np_(g[k]).reshape(ref[k].shape).astype(np.float64)
            scale = max(float(np.abs(ref[k]).max()), 1e-30)
            dd = np.abs(a - ref[k].astype(np.float64)).reshape(a.shape[0], -1).max(1) / scale
            gr[k] = {"max_rel": float(dd[keepg].max()), "max_rel_ambiguous": float(dd[~keepg].max()) if (~keepg).any() else 0.0,
                     "rows_over_1e-4": int((dd[keepg] > 1e-4).sum())}
        gm = np.abs(P.np_(rast.gstate.grad_means2d).astype(np.float64) - ref["vmeans2d"]).max(1) / max(np.abs(ref["vmeans2d"]).max(), 1e-30)
        gr["grad_means2d"] = {"max_rel": float(gm[keepg].max())}
        rec["grads"] = gr
        rec["ambiguous_gaussians"] = int((~keepg).sum())
        # timing: per-stage events, 10 reps
        rast.profile(True)
        acc = {}
        for _ in range(10):
            P.gpu_forward(rast, dev, cam, deg)
            P.gpu_backward(rast, dev, cam, deg, vpd)
            for k, v in rast.stage_times_ms().items():
                acc[k] = acc.get(k, 0.0) + v / 10
        rast.profile(False)
        rec["ms"] = {k: round(v, 4) for k, v in acc.items()}
        rec["ms_step"] = round(sum(acc.values()), 4)
        rec["pass_flat"] = bool(rec["img_max_err"] <= 1e-5 and all(v["max_rel"] <= 1e-4 for v in gr.values()))
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(name, label, "img", f"{rec['img_max_err']:.2e}", [f"{x:.1e}" for x in rec["img_max_err_per_channel"]], "amb",
              rec["ambiguous_px"], f"{rec['img_max_err_ambiguous']:.1e}", "grads", {k: f"{v['max_rel']:.1e}" for k, v in gr.items()},
              "fwd/bwd ms", rec["ms"]["render_fwd"], rec["ms"]["render_bwd"], "step", rec["ms_step"], "PASS" if rec["pass_flat"] else "FAIL")
        del rast


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "math_ab.jsonl"), "w") as out:
        exp_probe(out)
        for name in (sys.argv[1:] or ["C1d", "C5", "C2"]):
            run_config(name, out)
