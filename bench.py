#!/usr/bin/env python3
"""bench.py — the rasterizer hot path on BASELINE.json's configurations.

    python bench.py --gpus N --steps K --warmup W [--config C2|C3|C4|C5]      # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W [--config ...]   # CPU restatement of the reference

Default (what the driver runs) is config C2, BASELINE.json's headline: synthetic 1M Gaussians, SH degree 3, one
1920x1088 (= 1080p rounded up to x16, as the reference's dataset loader does) view per step, mode :rgbd, rasterize
forward + backward.  The other configurations of BASELINE.json:
  C3  3M Gaussians, a batch of 8 posed 1080p views per step, view-sharded over the ranks (8/N views per rank), ONE
      gradient exchange per batch (fused into the per-Gaussian backward, over NVLink when N > 1)   — strong scaling
  C4  6M Gaussians, forward-only 3840x2160 :rgbdn render (scripts/render-views.jl), one view per rank per step
  C5  500k Gaussians, 1312x848 (1297x840 rounded up): forward + backward + update_stats! (strategy.jl:118-136)
N > 1 (C2/C5): weak scaling — parameters replicated, every rank renders its own view per step, the per-Gaussian backward
reduces the gradients across the ranks inside the kernel (peer memory); value = views/s over all ranks.

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

# kind: train = one view fwd+bwd per rank and step; batch = V views per step over all ranks; render = forward only
META = {
    "C2": dict(kind="train", stats=False, scaling="weak", unit="steps/s",
               metric="fwd+bwd raster steps/sec @1M Gaussians 1080p",
               workload="C2: synthetic 1M Gaussians, SH degree 3, one 1920x1088 view, :rgbd, rasterize forward+backward"),
    "C3": dict(kind="batch", stats=False, scaling="strong", unit="steps/s",
               metric="fwd+bwd raster view-batch steps/sec @3M Gaussians, 8 views 1080p per step",
               workload="C3: synthetic 3M Gaussians, SH degree 3, batch of 8 posed 1920x1088 views per step, :rgbd, "
                        "view-sharded over the ranks, one gradient exchange per batch"),
    "C4": dict(kind="render", stats=False, scaling="weak", unit="views/s",
               metric="forward render views/sec @6M Gaussians 3840x2160",
               workload="C4: synthetic 6M Gaussians, SH degree 3, forward-only 3840x2160 :rgbdn render (render-views path)"),
    "C5": dict(kind="train", stats=True, scaling="weak", unit="steps/s",
               metric="training-shaped raster steps/sec @500k Gaussians 1312x848 (fwd+bwd+update_stats)",
               workload="C5: synthetic 500k Gaussians, SH degree 3, one 1312x848 view, :rgbd, rasterize forward+backward + "
                        "per-Gaussian 2D-gradient accumulation (update_stats!)"),
}
CH = {"rgb": 3, "rgbd": 5, "rgbdn": 8}


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


def stage_bytes(N, V, M, T, P, C, K, k, backward=True):
    """ALGORITHMIC bytes per stage and view — SURVEY.md §8(d)."""
    Cp = C if C > 3 else 3
    tile_bits, depth_bits = int(np.ceil(np.log2(max(T, 2)))), 27  # significant key bits (Appendix A.4)
    presort = os.environ.get("GSR_PRESORT", "1") != "0"
    # depth pre-sort (default): 4 radix passes over the N Gaussians' (depth, index) pairs, then only the tile digits
    # over the M instances; without it every significant bit is sorted over M
    passes = int(np.ceil(tile_bits / 8)) if presort else int(np.ceil((tile_bits + depth_bits) / 8))
    ppasses = int(np.ceil((1 + depth_bits) / 8))
    b = {
        "presort": (8 * N + 12 * N + 24 * N * ppasses) if presort else 0,
        "preprocess": 40 * N + 4 * N + V * (12 * k + 28 + 4 * Cp + 3 + 4),
        "scan": 8 * N + (4 * N if presort else 0),
        "duplicate": 20 * V + 12 * M + (4 * N if presort else 0),
        "sort": 8 * M + 24 * M * passes,
        "ranges": 8 * M + 8 * T,
        "render_fwd": M * (28 + 4 * C) + P * (4 * C + 8),
    }
    if backward:
        b.update({
            "zero_grads": 4 * N * (C + 6),
            "render_bwd": M * (28 + 4 * C) + P * (4 * C + 8) + 2 * 4 * (C + 6) * M,
            "gauss_bwd": N * (44 + 3 + 12 * k) + V * (12 + 8 + 12 + 4 * Cp) + N * (40 + 12 * K) + 12 * N,
        })
    return b


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle (CPU restatement of the reference kernels, OpenMP).  Test / measurement infrastructure only.
# ---------------------------------------------------------------------------------------------------------------
def cpu_threads():
    """All host cores, regardless of a launcher's OMP_NUM_THREADS=1 (torchrun exports that to its workers)."""
    from oracle.oracle import Oracle
    o = Oracle(np.float32)
    return o, o.set_threads(os.cpu_count() or 1)


def cpu_oracle_view(o, sc, mode, backward=True):
    """One forward (+backward) of the identity view on the CPU.  Returns seconds (fwd, bwd) and the oracle state."""
    from oracle.oracle import OracleCamera
    from gsrast.synthetic import make_vpixels
    cam = OracleCamera.simple(sc.fx, sc.fy, sc.width, sc.height)
    t0 = time.perf_counter()
    _, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, mode=mode, sh_degree=sc.sh_degree)
    t1 = time.perf_counter()
    if backward:
        vp = make_vpixels(sc.width, sc.height, CH[mode], 1002)
        o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, cam, st, mode=mode, sh_degree=sc.sh_degree)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, st


def cpu_step_seconds(o, sc, cfg, mode, n_views):
    """Seconds of one `step` of configuration cfg on the CPU + a description of what was timed.  A batch step (C3) is
    timed on ONE of its views and scaled by the view count (bounded sample)."""
    meta = META[cfg]
    tf, tb, st = cpu_oracle_view(o, sc, mode, backward=meta["kind"] != "render")
    if meta["kind"] == "batch":
        return (n_views * (tf + tb),
                f"1 of the {n_views} views of a batch timed in full ({tf:.2f}s fwd + {tb:.2f}s bwd), x{n_views}", st)
    if meta["kind"] == "render":
        return tf, f"1 full forward of the same workload ({tf:.2f}s)", st
    return tf + tb, f"1 full forward+backward step of the same workload ({tf:.2f}s fwd + {tb:.2f}s bwd)", st


def run_reference(args, real_stdout):
    """--impl reference: the reference's kernels cannot run on a CPU (`@kernel cpu=false`) and there is no Julia
    toolchain, so this arm times the CPU restatement (oracle/gsr_oracle.c, kind "port") on all host cores: FULL steps of
    the same workload, as many of the requested K as fit in ~150 s (`steps` reports the count actually timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from gsrast.synthetic import CONFIGS, make_config
    cfg = args.config
    n, deg, W, H, mode, seed, n_views = CONFIGS[cfg]
    sc = make_config(cfg)
    o, cores = cpu_threads()
    meta = META[cfg]
    t_first, sample, _ = cpu_step_seconds(o, sc, cfg, mode, n_views)  # warm-up (page-in, first touch)
    per_timed = t_first / n_views if meta["kind"] == "batch" else t_first  # wall seconds one timed sample takes
    k = max(1, min(args.steps, int(150.0 / max(per_timed, 1e-3)) - 1))
    times = []
    for _ in range(k):
        t, sample, _ = cpu_step_seconds(o, sc, cfg, mode, n_views)
        times.append(t)
    t_med = float(np.median(times))
    value = 1.0 / t_med  # one unit per step: a view (train / render) or a batch
    line = {
        "impl": "reference", "metric": meta["metric"], "value": value, "unit": meta["unit"], "n_gpus": args.gpus,
        "steps": k, "steps_requested": args.steps, "warmup": 1, "ms_per_step": 1e3 * t_med, "higher_is_better": True,
        "scaling": meta["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": meta["workload"], "views_per_step": n_views if meta["kind"] == "batch" else 1},
        "cpu_baseline": {"value": value, "unit": meta["unit"], "cores": cores, "kind": "port",
                         "sample": f"each timed step: {sample}; {k} steps timed, median"},
        "e2e": {"value": value, "unit": meta["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference kernels (the Julia reference cannot execute on a CPU: `@kernel cpu=false`, "
                "no Julia toolchain); OpenMP team = all host cores regardless of OMP_NUM_THREADS"
                + ("; ms_per_step is the batch estimate, the wall time of a timed sample is 1/8 of it"
                   if meta["kind"] == "batch" else ""),
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
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="how a step with several views in total combines them: one fused per-Gaussian backward over all "
                         "views (default), or a per-view backward accumulating into the table + NCCL all-reduce")
    ap.add_argument("--config", default="C2", choices=sorted(META))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    real_stdout = quiet_stdout()
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import torch.distributed as dist
    from gsrast import Camera, GaussianRasterizer, _lib, update_stats
    from gsrast.distributed import GradientTable, ViewBatchBackward, allreduce_gradients_, views_for_rank
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

    cfg, meta = args.config, META[args.config]
    kind = meta["kind"]
    n, deg, W, H, mode, seed, cfg_views = CONFIGS[cfg]
    sc = make_config(cfg)
    C = CH[mode]
    K = sc.shs.shape[1]
    # the views of one step over all ranks: `world` near-identical views (weak scaling: the same work per rank to within
    # a few percent) for train / render; the configuration's posed batch (±20 deg yaw, ±1 shift) for a batch step
    V = cfg_views if kind == "batch" else world
    poses = [view_pose(v, V, max_yaw_deg=20.0, max_shift=1.0) if kind == "batch"
             else view_pose(v, V, max_yaw_deg=2.0, max_shift=0.1) for v in range(V)]
    cams = [Camera(fx=sc.fx, fy=sc.fy, width=W, height=H, R=Rv, t=tv) for Rv, tv in poses]
    mine = views_for_rank(V, rank, world)
    if not mine:
        raise SystemExit(f"{cfg}: {V} views cannot occupy {world} ranks")
    cam = cams[mine[0]]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = dict(means=pin(sc.means), shs=pin(sc.shs), opac=pin(sc.opacities.reshape(-1, 1)), scales=pin(sc.scales),
                rots=pin(sc.rotations))
    vpix_h = pin(make_vpixels(W, H, C, seed))  # the same cotangent for every view
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    vpix = vpix_h.to(dev, non_blocking=True)
    rast = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=args.math, device=dev)
    stats = None
    if meta["stats"]:
        stats = (torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev))

    # one flat gradient table (59 floats / Gaussian at K=16) so that a cross-GPU reduction is one collective
    table = GradientTable(n, K, dev) if kind != "render" else None
    outs = table.outs() if table is not None else None

    def fwd(c, out=None):
        return rast._forward(d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, c, deg, (0, 0, 0), None,
                             None, out=out)

    def step_plain():  # per-view backward into the table (accumulating over this rank's views), NCCL all-reduce if N > 1
        for j, v in enumerate(mine):
            fwd(cams[v])
            rast._backward(vpix, d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cams[v], deg, (0, 0, 0),
                           outs=dict(outs), accumulate=(j > 0))
            if stats is not None:
                update_stats(*stats, rast)
        if world > 1:
            allreduce_gradients_(table)

    # fused: one accumulator per view, ONE per-Gaussian backward (+ the gradient exchange over NVLink peer memory when
    # N > 1) per step — csrc/backward_peers.cu.  Used whenever a step has more than one view in total.
    fused, reduction = None, "none (one view on one GPU)"
    if kind != "render" and V > 1:
        reduction = "accumulating per-view backward" if world == 1 else "NCCL all-reduce of the 59-float/Gaussian table"
        if args.exchange == "fused":
            try:
                rast_f = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=args.math, device=dev)
                fused = ViewBatchBackward(rast_f, n, K, cams)
                reduction = ("fused view-batch backward: one per-Gaussian pass over all views' moment accumulators"
                             if world == 1 else
                             "peer-fused: P2P loads of the views' moment accumulators + P2P stores of the reduced rows (NVLink)")
            except Exception as e:  # symmetric memory unavailable: keep the plain path
                sys.stderr.write(f"[bench] fused path unavailable ({e}); using {reduction}\n")
                fused = None
    vp_mine = {v: vpix for v in mine}

    def step_fused():
        fused.step(d, vp_mine, deg)
        if stats is not None:
            update_stats(*stats, fused.rast)

    def step_render():
        fwd(cam)

    step = step_render if kind == "render" else (step_fused if fused is not None else step_plain)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    barrier()
    launches0 = _lib.lib().gsr_launch_count()
    w0 = time.time()
    total_ms = timed(step, args.steps, 0)
    w1 = time.time()
    launches = int(_lib.lib().gsr_launch_count() - launches0)
    clocks = sampler.stop(w0, w1) if sampler else None
    ms_per_step = total_ms / args.steps
    units_per_step = 1 if kind == "batch" else world  # batches, or views over all ranks
    value = units_per_step * args.steps / (total_ms * 1e-3)
    alt_value, scatter_value, scatter_err = None, None, None
    if fused is not None:  # the same step through the other exchange path, for comparison
        alt_value = units_per_step * args.steps / (timed(step_plain, args.steps, 3) * 1e-3)
    if fused is not None and world > 1 and not args.no_parity_check:
        # ... and with reduce-scatter semantics (every rank keeps the reduced rows of ITS Gaussian slice only, what a
        # Gaussian-sharded optimizer consumes): the same kernel without the all-gather half of the exchange
        rast_s = GaussianRasterizer(width=W, height=H, mode=mode, math_mode=args.math, device=dev)
        fs = ViewBatchBackward(rast_s, n, K, cams, scatter_only=True)
        scatter_value = units_per_step * args.steps / (timed(lambda: fs.step(d, vp_mine, deg), args.steps, 3) * 1e-3)
        fused.step(d, vp_mine, deg)
        torch.cuda.synchronize()
        lo_g, hi_g = fs.slice_rows()
        es = torch.tensor([max(float((fs.views[k][lo_g:hi_g].double() - fused.views[k][lo_g:hi_g].double()).abs().max()
                                     / fused.views[k].double().abs().max().clamp_min(1e-30)) for k in fused.views)], device=dev)
        dist.all_reduce(es, op=dist.ReduceOp.MAX)
        scatter_err = float(es.item())
        del fs, rast_s

    # ---- parity of the multi-view / multi-GPU value path, outside the timed region ------------------------------
    parity_check = None
    if fused is not None and not args.no_parity_check:
        fused.step(d, vp_mine, deg)
        torch.cuda.synchronize()
        got = {k: v.clone() for k, v in fused.views.items()}
        step_plain()
        torch.cuda.synchronize()
        keys = ("vrot", "vmeans", "vscales", "vopacities", "vshs")
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        e_other = max(rel(got[k], outs[k]) for k in keys)
        # rank 0 alone renders ALL views of the step on its own GPU and accumulates (the single-GPU path the -m gpu tests
        # pin against the oracle); every rank's fused table must equal it
        e_single, identical = None, None
        if world > 1:
            if rank == 0:
                for v in range(V):
                    fwd(cams[v])
                    rast._backward(vpix, d["means"], d["shs"], d["opac"], d["scales"], d["rots"], None, None, cams[v], deg,
                                   (0, 0, 0), outs=dict(outs), accumulate=(v > 0))
            dist.broadcast(table.flat, src=0)
            es = torch.tensor([max(rel(got[k], outs[k]) for k in keys), e_other], device=dev)
            dist.all_reduce(es, op=dist.ReduceOp.MAX)
            e_single, e_other = float(es[0].item()), float(es[1].item())
            ref0 = fused.table_flat.clone()
            dist.broadcast(ref0, src=0)
            same = torch.tensor([1.0 if torch.equal(ref0, fused.table_flat) else 0.0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            identical = bool(same.item())
            del ref0
        # Both sides of these comparisons carry fp32 atomic-order noise (the oracle-pinned tolerance of either is 1e-4
        # of the tensor's max, vrot being the sensitive one), so two of them may differ by up to 2e-4.
        worst = max(e_other, e_single or 0.0)
        parity_check = {"max_rel": worst, "ok": bool(worst <= 2e-4 and identical is not False),
                        "fused_vs_plain_path": e_other, "fused_vs_all_views_on_rank0": e_single,
                        "tables_bit_identical_across_ranks": identical, "tolerance": 2e-4,
                        "note": "max over ranks and tensors of ||a-b||inf / ||b||inf between two GPU results, each held to "
                                "1e-4 against the oracle elsewhere (hence 2e-4 here): tools/peers_check.py compares the "
                                "same kernels with the sum of per-view ORACLE gradients at 1e-4 (profiles/), "
                                "tests/test_gpu_parity.py the single-GPU forms"}

    # ---- end to end: host buffers in, host buffers out --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, dev, rank, world, kind, meta, rast, fused, host, vpix_h, d, cams, mine, deg, n, K, C,
                      W, H, table, stats, update_stats, allreduce_gradients_, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-stage device times (separate pass, CUDA events inside the library on the launching stream) -----
    rast.profile(True)
    acc = {}
    reps = 10
    for _ in range(reps):
        fwd(cam)
        if kind != "render":
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
    Vis = int((radii > 0).sum())
    M = int(rast.n_rendered)
    T, P = rast.n_tiles, W * H
    k_used = (deg + 1) ** 2
    bts = stage_bytes(n, Vis, M, T, P, C, K, k_used, backward=kind != "render")
    hbm_peak, peak_src = measured_peaks()
    stages = {}
    for name, msv in acc.items():
        if name not in bts or msv <= 0:
            continue
        gbs = bts[name] / (msv * 1e-3) / 1e9
        stages[name] = {"ms": round(msv, 4), "alg_bytes": int(bts[name]), "gbs": round(gbs, 1),
                        "hbm_frac": round(gbs / hbm_peak, 4)}
    traffic = {}
    if cfg == "C2":
        try:  # DRAM bytes per launch from the committed ncu capture (profiles/), same workload
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["per_launch_bytes"]
        except Exception:
            pass
    t_view = sum(s["ms"] for s in stages.values())
    dominant = max(stages, key=lambda k2: stages[k2]["ms"])
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": stages[dominant]["gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": stages[dominant]["hbm_frac"], "traffic": traffic.get(dominant),
                "peak_source": peak_src,
                "traffic_source": "profiles/ncu_traffic.json (ncu dram bytes per launch)" if dominant in traffic else None,
                "ms": stages[dominant]["ms"], "share_of_step": round(stages[dominant]["ms"] / t_view, 3),
                "note": "the compositing kernels are FP32-issue-bound, not HBM-bound (SURVEY.md §8d): see roofline_fp32"}
    view_bytes = sum(s["alg_bytes"] for s in stages.values())

    cpu_baseline, roofline_fp32 = None, None
    if not args.no_cpu_baseline and world == 1:
        o, cores = cpu_threads()
        t_cpu, sample, st = cpu_step_seconds(o, sc, cfg, mode, cfg_views)
        cpu_baseline = {"value": 1.0 / t_cpu, "unit": meta["unit"], "cores": cores, "kind": "port", "sample": sample,
                        "note": "CPU restatement of the reference kernels (oracle/gsr_oracle.c, OpenMP); the Julia "
                                "reference has no CPU backend for this path"}
        Ef, Bf = (int(x) for x in st.counts_fwd)
        Eb, Bb = (int(x) for x in st.counts_bwd)
        fl = {"render_fwd": 14 * Ef + (2 + 3 * C) * Bf}
        if kind != "render":
            fl["render_bwd"] = 14 * Eb + (30 + 9 * C) * Bb
        roofline_fp32 = {"peak": round(fp32_peak, 2), "unit": "TFLOP/s", "peak_source": "FFMA micro-benchmark, this run",
                         "pairs": {"evaluated_fwd": Ef, "blended_fwd": Bf, "evaluated_bwd": Eb, "blended_bwd": Bb},
                         "note": "flops of the reference's pair counts, identity view (the view `stages` times)"}
        t_lower = 0.0
        for name, sg in stages.items():
            tl = sg["alg_bytes"] / (hbm_peak * 1e9)
            if name in fl:
                tf32 = fl[name] / (fp32_peak * 1e12)
                ach = fl[name] / (sg["ms"] * 1e-3) / 1e12
                roofline_fp32[name] = {"alg_flops": fl[name], "achieved": round(ach, 3), "frac": round(ach / fp32_peak, 4)}
                tl = max(tl, tf32)
            t_lower += tl
        roofline_fp32["t_lower_ms"] = round(1e3 * t_lower, 4)
        roofline_fp32["step_frac_of_binding_roofline"] = round(1e3 * t_lower / t_view, 4)

    line = {
        "metric": meta["metric"], "value": value, "unit": meta["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": meta["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": meta["workload"], "N": n, "V": Vis, "M": M, "tiles": T, "sh_degree": deg, "mode": mode,
                   "math_mode": args.math, "views_per_step": V, "views_per_step_per_gpu": len(mine),
                   "parallelism": f"view-sharded x{world}", "gradient_reduction": reduction,
                   "value_through_the_plain_path": alt_value,
                   "value_with_reduce_scatter_only": scatter_value, "reduce_scatter_slice_vs_allreduce_table": scatter_err,
                   "l2": f"inputs larger than L2: {4 * n * (11 + 3 * K) / 1e6:.0f} MB parameters"
                         + ("" if kind == "render" else " + as many of gradients") + " + per-view state vs 126 MB L2"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "parity_check": parity_check,
        "roofline": roofline, "roofline_fp32": roofline_fp32, "stages": stages,
        "view_hbm": {"alg_bytes": int(view_bytes), "stage_ms": round(t_view, 4),
                     "gbs": round(view_bytes / (t_view * 1e-3) / 1e9, 1),
                     "frac": round(view_bytes / (t_view * 1e-3) / 1e9 / hbm_peak, 4),
                     "note": "stages, their bytes and fractions are per view (identity pose), measured on rank 0"},
        "cpu_baseline": cpu_baseline,
    }
    emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, torch, dist, dev, rank, world, kind, meta, rast, fused, host, vpix_h, d_res, cams, mine, deg, n, K, C, W, H,
            table, stats, update_stats, allreduce_gradients_, barrier):
    """The same metric with HOST buffers on both sides of every step: pinned inputs -> device, the step, results -> pinned
    host, all inside the timed region.  What one step of one rank uploads / downloads:
      train, N = 1   gsr_forward_backward_host_async (C ABI): 5 parameter arrays + cotangent up; image + 5 gradient arrays down
      train, N > 1   the same uploads per rank; down: this rank's image + ITS SLICE of the reduced gradient table (the ranks'
                     slices tile the table: the job downloads every gradient exactly once)
      batch          parameters once + one cotangent per local view up; this rank's table slice + its views' images down
      render         the camera (164 B, passed by value) up — the scene is the resident model, as in render-views.jl; image down
    Uploads, compute and downloads of consecutive steps overlap (two device slots, three streams)."""
    unit = meta["unit"]
    par_bytes = sum(v.numel() for v in host.values()) * 4
    img_bytes = H * W * C * 4
    units_per_step = 1 if kind == "batch" else world
    cam = cams[mine[0]]

    if kind == "train" and world == 1:
        out_h = dict(image=torch.empty((H, W, C)).pin_memory(), vmeans=torch.empty((n, 3)).pin_memory(),
                     vshs=torch.empty((n, K, 3)).pin_memory(), vopacities=torch.empty((n, 1)).pin_memory(),
                     vscales=torch.empty((n, 3)).pin_memory(), vrot=torch.empty((n, 4)).pin_memory())
        h2d = par_bytes + vpix_h.numel() * 4
        d2h = sum(v.numel() * 4 for v in out_h.values())

        def step_e2e():  # the C-ABI host entry point, pipelined: step k+1's H2D overlaps step k's compute / D2H
            rast.forward_backward_host(host, vpix_h, cam, deg, out=out_h, wait=False)
            if stats is not None:
                update_stats(*stats, rast)

        def drain():
            rast.host_wait()
        api = ("gsr_forward_backward_host_async + gsr_host_wait (C ABI, pinned host buffers, double-buffered staging: "
               "H2D / compute / D2H of consecutive steps overlap)")
    else:
        cur = torch.cuda.current_stream(dev)
        s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        n_local = len(mine)
        per = 4 + 3 + 3 + 1 + 3 * K
        lo = (n * per // world) * rank
        hi = n * per if rank == world - 1 else (n * per // world) * (rank + 1)
        slots = []
        for _ in range(2):
            sl = dict(ev_up=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_down=torch.cuda.Event(),
                      img=[torch.empty((H, W, C), device=dev) for _ in range(n_local)],
                      img_h=[torch.empty((H, W, C)).pin_memory() for _ in range(n_local)])
            if kind != "render":
                sl["d"] = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
                sl["vp"] = [torch.empty_like(vpix_h, device=dev) for _ in range(n_local)]
                sl["grad_d"] = torch.empty(hi - lo, device=dev)
                sl["grad_h"] = torch.empty(hi - lo).pin_memory()
            slots.append(sl)
        h2d = (par_bytes + n_local * vpix_h.numel() * 4) if kind != "render" else 164
        d2h = n_local * img_bytes + ((hi - lo) * 4 if kind != "render" else 0)
        r_used = fused.rast if fused is not None else rast
        state = {"i": 0}

        def step_e2e():
            sl = slots[state["i"] & 1]
            state["i"] += 1
            if kind != "render":
                with torch.cuda.stream(s_up):
                    s_up.wait_event(sl["ev_done"])  # the step that last used this slot has consumed its inputs
                    for k2, v in host.items():
                        sl["d"][k2].copy_(v, non_blocking=True)
                    for vp in sl["vp"]:
                        vp.copy_(vpix_h, non_blocking=True)
                    sl["ev_up"].record(s_up)
                cur.wait_event(sl["ev_up"])
            cur.wait_event(sl["ev_down"])  # this slot's result buffers were downloaded
            dd = sl["d"] if kind != "render" else d_res
            if kind == "render":
                rast._forward(dd["means"], dd["shs"], dd["opac"], dd["scales"], dd["rots"], None, None, cam, deg, (0, 0, 0),
                              None, None, out=sl["img"][0])
            elif fused is not None:
                imgs = {}
                fused.step(dd, {v: sl["vp"][j] for j, v in enumerate(mine)}, deg, images=imgs if n_local > 1 else None)
                if stats is not None:
                    update_stats(*stats, r_used)
                for j, v in enumerate(mine):
                    sl["img"][j].copy_(imgs[v] if n_local > 1 else r_used.image, non_blocking=True)
                sl["grad_d"].copy_(fused.table_flat[lo:hi], non_blocking=True)  # the next exchange overwrites the table
            else:
                for j, v in enumerate(mine):
                    rast._forward(dd["means"], dd["shs"], dd["opac"], dd["scales"], dd["rots"], None, None, cams[v], deg,
                                  (0, 0, 0), None, None, out=sl["img"][j])
                    rast._backward(sl["vp"][j], dd["means"], dd["shs"], dd["opac"], dd["scales"], dd["rots"], None, None,
                                   cams[v], deg, (0, 0, 0), outs=dict(table.outs()), accumulate=(j > 0))
                    if stats is not None:
                        update_stats(*stats, rast)
                if world > 1:
                    allreduce_gradients_(table)
                sl["grad_d"].copy_(table.flat[lo:hi], non_blocking=True)
            sl["ev_done"].record(cur)
            with torch.cuda.stream(s_down):
                s_down.wait_event(sl["ev_done"])
                for j in range(n_local):
                    sl["img_h"][j].copy_(sl["img"][j], non_blocking=True)
                if kind != "render":
                    sl["grad_h"].copy_(sl["grad_d"], non_blocking=True)
                sl["ev_down"].record(s_down)

        def drain():
            torch.cuda.synchronize()
        api = ("pinned host buffers; uploads / step / downloads of consecutive steps on three streams (two device slots); "
               + ("gsr_forward" if kind == "render" else
                  "gsr_forward + gsr_backward_render per view, gsr_backward_gaussians_views per step" if fused is not None
                  else "gsr_forward + gsr_backward per view (+ NCCL all-reduce)"))

    for _ in range(3):
        step_e2e()
    drain()
    barrier()
    ke = max(5, min(args.steps, 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ke):
        step_e2e()
    drain()  # every step's D2H has landed in the host buffers
    e1.record()
    barrier()
    mse = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(mse, op=dist.ReduceOp.MAX)
    return {"value": units_per_step * ke / (float(mse.item()) * 1e-3), "unit": unit, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "bytes_are": "per rank", "ms_per_step": float(mse.item()) / ke, "steps": ke,
            "api": api}


if __name__ == "__main__":
    sys.exit(main())
