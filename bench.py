#!/usr/bin/env python3
"""bench.py — forward+backward raster steps/s on BASELINE.json's headline configuration.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement of the reference kernels

Workload (config C2): synthetic 1M Gaussians, SH degree 3, one 1920x1088 (= 1080p rounded up to x16, as the
reference's dataset loader does) view per step, mode :rgbd, rasterize forward + backward.
N > 1: weak scaling — parameters replicated, every rank renders its own view of the same scene per step,
then the per-Gaussian gradients (59 floats / Gaussian) are summed with one NCCL all-reduce; value = views/s
over all ranks (a "step" stays one view's forward+backward).

One JSON line on stdout (rank 0).  See DESIGN.md §"Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "fwd+bwd raster steps/sec @1M Gaussians 1080p"
UNIT = "steps/s"
WORKLOAD = "C2: synthetic 1M Gaussians, SH degree 3, one 1920x1088 view, :rgbd, rasterize forward+backward"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for _, ln in self.lines]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def stage_bytes(N, V, M, T, P, C, K, k):
    """ALGORITHMIC bytes per stage and step — SURVEY.md §8(d)."""
    Cp = C if C > 3 else 3
    tile_bits, depth_bits = int(np.ceil(np.log2(max(T, 2)))), 27  # significant key bits (Appendix A.4)
    presort = os.environ.get("GSR_PRESORT", "1") != "0"
    # depth pre-sort (default): 4 radix passes over the N Gaussians' (depth, index) pairs, then only the tile digits
    # over the M instances; without it every significant bit is sorted over M
    passes = int(np.ceil(tile_bits / 8)) if presort else int(np.ceil((tile_bits + depth_bits) / 8))
    ppasses = int(np.ceil((1 + depth_bits) / 8))
    return {
        "presort": (8 * N + 12 * N + 8 * N + 24 * N * ppasses) if presort else 0,
        "preprocess": 40 * N + 4 * N + V * (12 * k + 28 + 4 * Cp + 3 + 4),
        "scan": 8 * N,
        "duplicate": 20 * V + 12 * M,
        "sort": 8 * M + 24 * M * passes,
        "ranges": 8 * M + 8 * T,
        "render_fwd": M * (28 + 4 * C) + P * (4 * C + 8),
        "zero_grads": 4 * N * (C + 6),
        "render_bwd": M * (28 + 4 * C) + P * (4 * C + 8) + 2 * 4 * (C + 6) * M,
        "gauss_bwd": N * (44 + 3 + 12 * k) + V * (12 + 8 + 12 + 4 * Cp) + N * (40 + 12 * K) + 12 * N,
    }


def cpu_oracle_step(sc, mode, band=None):
    """One forward+backward of the CPU restatement (oracle/, OpenMP over all host cores).  Returns seconds,
    counts.  `band=(y0,y1)` restricts the two compositing stages to a band of tile rows (bounded sample)."""
    from oracle.oracle import Oracle, OracleCamera
    from gsrast.synthetic import make_vpixels
    o = Oracle(np.float32)
    cam = OracleCamera.simple(sc.fx, sc.fy, sc.width, sc.height)
    C = {"rgb": 3, "rgbd": 5, "rgbdn": 8}[mode]
    vp = make_vpixels(sc.width, sc.height, C, 1002)
    t0 = time.perf_counter()
    img, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode=mode,
                        sh_degree=sc.sh_degree, tile_rows=band)
    t1 = time.perf_counter()
    o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, st, mode=mode,
               sh_degree=sc.sh_degree, tile_rows=band)
    t2 = time.perf_counter()
    return t2 - t0, (t1 - t0, t2 - t1), st


def run_reference(args, real_stdout):
    """--impl reference: the reference's kernels cannot run on a CPU (`@kernel cpu=false`) and there is no Julia
    toolchain, so this arm times the CPU restatement (oracle/gsr_oracle.c, kind "port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from gsrast.synthetic import make_config
    sc = make_config("C2")
    cores = os.cpu_count() or 1
    gy = sc.height // 16
    y0, y1 = gy * 7 // 16, gy * 7 // 16 + max(1, gy // 8)  # a band of tile rows in the middle of the image
    # full per-Gaussian stages + sort every step; compositing fwd/bwd on the band, scaled by instance share
    times = []
    share = None
    for i in range(args.warmup + args.steps):
        t, _, st = cpu_oracle_step(sc, "rgbd", band=(y0, y1))
        if share is None:
            r = st.ranges.astype(np.int64)
            gx = sc.width // 16
            inst_band = int((r[y0 * gx:y1 * gx, 1] - r[y0 * gx:y1 * gx, 0]).sum())
            share = inst_band / max(1, st.n_rendered)
        if i >= args.warmup:
            times.append(t)
    # estimate the non-compositing part once (band of zero rows)
    t_rest, _, _ = cpu_oracle_step(sc, "rgbd", band=(gy, gy))  # empty band: compositing skipped
    t_med = float(np.median(times))
    t_comp_band = max(t_med - t_rest, 1e-6)
    t_full = t_rest + t_comp_band / share
    value = 1.0 / t_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_full, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "views_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": (f"per step: all per-Gaussian stages + sort at full size, compositing fwd+bwd on tile "
                                    f"rows [{y0},{y1}) of {gy} ({100 * share:.1f}% of the tile instances), scaled by that "
                                    f"share; measured {1e3 * t_med:.0f} ms/sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference kernels (the Julia reference cannot execute on CPU); not the reference binary",
    }
    emit(line, real_stdout)
    return 0


def emit(line: dict, real_stdout_fd: int):
    """The ONE JSON line goes to the real stdout; everything else any library prints during the run (NCCL's version
    banner, warnings) was diverted to stderr by `quiet_stdout`."""
    os.write(real_stdout_fd, (json.dumps(line) + "\n").encode())


def quiet_stdout() -> int:
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the process (C libraries included)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gsrast", choices=["gsrast", "reference"])
    ap.add_argument("--math", default="strict", choices=["strict", "reference", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--config", default="C2")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    real_stdout = quiet_stdout()
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import torch.distributed as dist
    from gsrast import Camera, GaussianRasterizer, _lib
    from gsrast.synthetic import CONFIGS, make_config, make_vpixels, view_pose

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgsrast has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, deg, W, H, mode, seed, _ = CONFIGS[args.config]
    sc = make_config(args.config)
    C = {"rgb": 3, "rgbd": 5, "rgbdn": 8}[mode]
    K = sc.shs.shape[1]
    R, t = view_pose(rank, world, max_yaw_deg=2.0, max_shift=0.1)  # near-identical work per rank (weak scaling)
    cam = Camera(fx=sc.fx, fy=sc.fy, width=W, height=H, R=R, t=t)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = dict(means=pin(sc.means), shs=pin(sc.shs), opac=pin(sc.opacities.reshape(-1, 1)), scales=pin(sc.scales),
                rots=pin(sc.rotations))
    vpix_h = pin(make_vpixels(W, H, C, seed))
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    vpix = vpix_h.to(dev, non_blocking=True)
    rast = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=args.math, device=dev)

    # one flat gradient table (59 floats / Gaussian at K=16) so that the cross-GPU reduction is one collective
    from gsrast.distributed import GradientTable, allreduce_gradients_
    table = GradientTable(n, K, dev)
    flat, outs = table.flat, table.outs()

    def step_nccl():  # baseline multi-GPU path: per-rank backward over all Gaussians, then one NCCL all-reduce
        rast._forward(d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cam, deg, (0, 0, 0), None, None)
        rast._backward(vpix, d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cam, deg, (0, 0, 0),
                       outs=dict(outs))
        allreduce_gradients_(table)

    # N > 1: per-Gaussian backward fused with the gradient reduction over NVLink peer memory (backward_peers.cu)
    fused, reduction = None, "none (single GPU)"
    if world > 1:
        reduction = "NCCL all-reduce of the 59-float/Gaussian table"
        try:
            from gsrast.distributed import PeerFusedBackward
            cams_all = [Camera(fx=sc.fx, fy=sc.fy, width=W, height=H, R=Rv, t=tv)
                        for Rv, tv in (view_pose(r, world, max_yaw_deg=2.0, max_shift=0.1) for r in range(world))]
            rast_f = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=args.math, device=dev)
            fused = PeerFusedBackward(rast_f, n, K, cams_all)
            reduction = "peer-fused: P2P loads of the moment accumulators + P2P stores of the reduced rows (NVLink)"
        except Exception as e:  # symmetric memory unavailable: keep the NCCL path
            sys.stderr.write(f"[bench] peer-fused path unavailable ({e}); using NCCL all-reduce\n")
            fused = None

    def step():
        if fused is not None:
            fused.step(d, vpix, deg)
        else:
            step_nccl()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    barrier()
    launches0 = _lib.lib().gsr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = time.time()
    launches = int(_lib.lib().gsr_launch_count() - launches0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop(w0, w1) if sampler else None
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)
    nccl_value = None
    if world > 1 and fused is not None:  # the same step with backward_gaussians + NCCL all-reduce, for comparison
        for _ in range(3):
            step_nccl()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_nccl()
        e1.record()
        barrier()
        msn = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(msn, op=dist.ReduceOp.MAX)
        nccl_value = world * args.steps / (float(msn.item()) * 1e-3)

    # ---- end to end: host buffers in, host buffers out ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        out_h = dict(image=torch.empty((H, W, C)).pin_memory(), vmeans=torch.empty((n, 3)).pin_memory(),
                     vshs=torch.empty((n, K, 3)).pin_memory(), vopacities=torch.empty((n, 1)).pin_memory(),
                     vscales=torch.empty((n, 3)).pin_memory(), vrot=torch.empty((n, 4)).pin_memory())
        h2d = sum(v.numel() * 4 for v in host.values()) + vpix_h.numel() * 4
        d2h = sum(v.numel() * 4 for v in out_h.values())
        flat_h = torch.empty_like(flat, device="cpu").pin_memory() if world > 1 else None

        def step_e2e():
            if world == 1:  # the C-ABI host entry point, pipelined: step k+1's H2D overlaps step k's compute / D2H
                rast.forward_backward_host(host, vpix_h, cam, deg, out=out_h, wait=False)
            else:
                dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                vp = vpix_h.to(dev, non_blocking=True)
                if fused is not None:
                    img, _ = fused.step(dd, vp, deg)
                    src_flat = fused.table_flat
                else:
                    img = rast._forward(dd["means"], dd["shs"], dd["opac"], dd["scales"], dd["rots"], None, None, cam,
                                        deg, (0, 0, 0), None, None)
                    rast._backward(vp, dd["means"], dd["shs"], dd["opac"], dd["scales"], dd["rots"], None, None, cam, deg,
                                   (0, 0, 0), outs=dict(outs))
                    allreduce_gradients_(table)
                    src_flat = flat
                out_h["image"].copy_(img, non_blocking=True)
                flat_h.copy_(src_flat, non_blocking=True)
                torch.cuda.synchronize()

        for _ in range(3):
            step_e2e()
        rast.host_wait()
        barrier()
        ke = max(5, min(args.steps, 20))
        e0.record()
        for _ in range(ke):
            step_e2e()
        rast.host_wait()  # every step's D2H has landed in the host buffers
        e1.record()
        barrier()
        mse = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(mse, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ke / (float(mse.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(mse.item()) / ke, "steps": ke,
               "api": "gsr_forward_backward_host_async + gsr_host_wait (C ABI, pinned host buffers, double-buffered "
                      "staging: H2D / compute / D2H of consecutive steps overlap)" if world == 1 else
                      "pinned torch copies + gsr_forward / backward + gradient reduction + D2H"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-stage device times (separate pass, CUDA events inside the library on the launching stream) -----
    rast.profile(True)
    acc = {}
    reps = 10
    for _ in range(reps):
        rast._forward(d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cam, deg, (0, 0, 0), None, None)
        rast._backward(vpix, d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cam, deg, (0, 0, 0),
                       outs=dict(outs))
        torch.cuda.synchronize()
        for k, v in rast.stage_times_ms().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    rast.profile(False)
    fp32 = _lib.C.c_double(0)
    _lib.check(_lib.lib().gsr_measure_fp32_peak(_lib.C.byref(fp32), None))
    fp32_peak = float(fp32.value)

    radii = rast.gstate.radii
    V = int((radii > 0).sum())
    M = int(rast.n_rendered)
    T, P = rast.n_tiles, W * H
    k_used = (deg + 1) ** 2
    bts = stage_bytes(n, V, M, T, P, C, K, k_used)
    hbm_peak, peak_src = measured_peaks()
    stages = {}
    for name, msv in acc.items():
        gbs = bts[name] / (msv * 1e-3) / 1e9 if msv > 0 else None
        stages[name] = {"ms": round(msv, 4), "alg_bytes": int(bts[name]), "gbs": None if gbs is None else round(gbs, 1),
                        "hbm_frac": None if gbs is None else round(gbs / hbm_peak, 4)}
    traffic = {}
    try:  # DRAM bytes per launch from the committed ncu capture (profiles/), same workload
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["per_launch_bytes"]
    except Exception:
        pass
    dominant = max(acc, key=lambda k2: acc[k2])
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": stages[dominant]["gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": stages[dominant]["hbm_frac"], "traffic": traffic.get(dominant), "peak_source": peak_src,
                "traffic_source": "profiles/ncu_traffic.json (ncu dram bytes per launch)" if dominant in traffic else None,
                "ms": stages[dominant]["ms"], "share_of_step": round(acc[dominant] / sum(acc.values()), 3),
                "note": "the compositing kernels are FP32/SFU-bound, not HBM-bound (SURVEY.md §8d): see roofline_fp32"}
    step_bytes = sum(bts.values())

    cpu_baseline, roofline_fp32 = None, None
    if not args.no_cpu_baseline and world == 1:
        t_cpu, (tf, tb), st = cpu_oracle_step(sc, mode)
        cpu_baseline = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": f"1 full forward+backward step of the same workload ({tf:.2f}s fwd + {tb:.2f}s bwd)",
                        "note": "CPU restatement of the reference kernels (oracle/gsr_oracle.c, OpenMP); the Julia "
                                "reference has no CPU backend for this path"}
        Ef, Bf = (int(x) for x in st.counts_fwd)
        Eb, Bb = (int(x) for x in st.counts_bwd)
        fl = {"render_fwd": 14 * Ef + (2 + 3 * C) * Bf, "render_bwd": 14 * Eb + (30 + 9 * C) * Bb}
        roofline_fp32 = {"peak": round(fp32_peak, 2), "unit": "TFLOP/s", "peak_source": "FFMA micro-benchmark, this run",
                         "pairs": {"evaluated_fwd": Ef, "blended_fwd": Bf, "evaluated_bwd": Eb, "blended_bwd": Bb}}
        t_lower = 0.0
        for name in acc:
            tl = bts[name] / (hbm_peak * 1e9)
            if name in fl:
                tf32 = fl[name] / (fp32_peak * 1e12)
                ach = fl[name] / (acc[name] * 1e-3) / 1e12
                roofline_fp32[name] = {"alg_flops": fl[name], "achieved": round(ach, 3), "frac": round(ach / fp32_peak, 4)}
                tl = max(tl, tf32)
            t_lower += tl
        roofline_fp32["t_lower_ms"] = round(1e3 * t_lower, 4)
        roofline_fp32["step_frac_of_binding_roofline"] = round(1e3 * t_lower / sum(acc.values()), 4)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "N": n, "V": V, "M": M, "tiles": T, "sh_degree": deg, "mode": mode,
                   "math_mode": args.math, "views_per_step_per_gpu": 1,
                   "parallelism": f"view-sharded x{world}", "gradient_reduction": reduction,
                   "value_with_nccl_allreduce": nccl_value,
                   "l2": "inputs larger than L2: 236 MB parameters + 236 MB gradients + ~0.5 GB state per step vs 126 MB L2"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "roofline_fp32": roofline_fp32, "stages": stages,
        "step_hbm": {"alg_bytes": int(step_bytes), "gbs": round(step_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                     "frac": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak, 4)},
        "cpu_baseline": cpu_baseline,
    }
    emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
