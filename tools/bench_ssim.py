"""Times the fused SSIM / photometric-loss kernels at the C2 image size (1920x1088, rgb of an :rgbd image) and
reports them against the HBM and FP32 rooflines.  Prints one JSON line.

Method: CUDA events around 8*R back-to-back launches through the C ABI with preallocated buffers (no host work in
the timed region); the launches rotate over 8 independent image sets (>= 400 MB in total, 126 MB L2), so every
launch finds its inputs in HBM, not in L2."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussiansplatting.jl_b200"))
from gsrast import GaussianRasterizer, _lib  # noqa: E402

NSETS, REPS = 8, 6


def p(t):
    return C.c_void_p(t.data_ptr())


def timed(launch):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(NSETS):
        launch(i, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        for i in range(NSETS):
            launch(i, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (REPS * NSETS)


def main():
    W, H, Cc = 1920, 1088, 5
    peak, fp32 = 6545.3, 72.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    lib = _lib.lib()
    rast = GaussianRasterizer(width=W, height=H, mode="rgbd")
    pk = C.c_double(0.0)
    if lib.gsr_measure_fp32_peak(C.byref(pk), None) == 0 and pk.value > 0:
        fp32 = pk.value
    xs = [torch.rand((1, 3, H, W), device="cuda") for _ in range(NSETS)]
    ts = [torch.rand((1, 3, H, W), device="cuda") for _ in range(NSETS)]
    ms_ = [torch.empty_like(xs[0]) for _ in range(NSETS)]
    ds = [[torch.empty_like(xs[0]) for _ in range(3)] for _ in range(NSETS)]
    dl = [torch.full_like(xs[0], -0.2 / xs[0].numel()) for _ in range(NSETS)]
    gs = [torch.empty_like(xs[0]) for _ in range(NSETS)]
    imgs = [torch.rand((H, W, Cc), device="cuda") for _ in range(NSETS)]
    tg3 = [t[0].contiguous() for t in ts]
    vps = [torch.empty_like(imgs[0]) for _ in range(NSETS)]
    loss = torch.empty(3, device="cuda")
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    n = xs[0].numel()
    # algorithmic flops per pixel-channel: 2 passes x 11 taps x 2 flops per convolved quantity, + products + map math
    f_fwd, f_bwd = 5 * 44 + 3 + 45, 3 * 44 + 3 + 6
    res = {}

    def add(name, ms, nbytes, flops):
        res[name] = {"ms": round(ms, 5), "alg_bytes": nbytes, "gbs": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / peak, 4),
                     "alg_flops": flops, "tflops": round(flops / ms / 1e9, 2), "fp32_frac": round(flops / ms / 1e9 / fp32, 4)}

    ms = timed(lambda i, st: lib.gsr_ssim_forward(W, H, 3, 1, p(xs[i]), p(ts[i]), c1, c2, 1, p(ms_[i]), p(ds[i][0]), p(ds[i][1]), p(ds[i][2]), st))
    add("ssim_fwd_train", ms, 24 * n, f_fwd * n)          # 2 reads + 4 writes
    ms = timed(lambda i, st: lib.gsr_ssim_forward(W, H, 3, 1, p(xs[i]), p(ts[i]), c1, c2, 0, p(ms_[i]), None, None, None, st))
    add("ssim_fwd_infer", ms, 12 * n, (f_fwd - 30) * n)
    ms = timed(lambda i, st: lib.gsr_ssim_backward(W, H, 3, 1, p(xs[i]), p(ts[i]), p(dl[i]), p(ds[i][0]), p(ds[i][1]), p(ds[i][2]), p(gs[i]), st))
    add("ssim_bwd", ms, 28 * n, f_bwd * n)                # 6 reads + 1 write
    ms = timed(lambda i, st: lib.gsr_photometric_loss(rast._h, p(imgs[i]), p(tg3[i]), 0.2, p(vps[i]), p(loss), st))
    # image rgb + target read twice, 3 maps written + read, cotangent (all C channels) written
    add("photometric_loss", ms, (2 * (4 + 4) + 2 * 12) * n + 4 * W * H * Cc, (f_fwd + f_bwd) * n)
    print(json.dumps({"workload": f"{W}x{H} rgb of an :rgbd raster image, {NSETS} rotating image sets (inputs never L2-resident)",
                      "hbm_peak_gbs": peak, "fp32_peak_tflops": round(fp32, 1), "kernels": res}))


if __name__ == "__main__":
    main()
