"""torchrun check (>= 2 GPUs) of the fused per-Gaussian backward + gradient exchange (csrc/backward_peers.cu):

  * against the ORACLE: the table every rank ends up with equals the sum over the batch's views of the CPU restatement's
    gradients (small scene; 1e-4 relative, the flat tolerance of the single-GPU tests);
  * against this repo's other path: backward_gaussians + NCCL all-reduce of the 59-float table;
  * one view per rank and several views per rank (V = 2 x world, capped at 16), odd Gaussian counts;
  * timing of both exchange paths.

    torchrun --nproc-per-node N tools/peers_check.py [c2] [odd]     -> gpurun_out/peers_check_w<N>[_c2].json

Test / measurement infrastructure: uses oracle/ (small scenes only).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gsrast import Camera, GaussianRasterizer  # noqa: E402
from gsrast.distributed import GradientTable, PeerFusedBackward, ViewBatchBackward, allreduce_gradients_, views_for_rank  # noqa: E402
from gsrast.synthetic import make_config, make_scene, make_vpixels, view_pose  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
big = "c2" in sys.argv[1:]
odd = "odd" in sys.argv[1:]
sc = make_config("C2") if big else make_scene(50_001 if odd else 50_000, 3, 640, 368, 7)
mode, C, deg, K = "rgbd", 5, 3, 16
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = dict(means=d(sc.means), shs=d(sc.shs), opac=d(sc.opacities.reshape(-1, 1)), scales=d(sc.scales), rots=d(sc.rotations))
yaw, shift = (2.0, 0.1) if big else (20.0, 1.0)
n = sc.n
KEYS = ("vrot", "vmeans", "vscales", "vopacities", "vshs")
report = {"world": world, "N": n, "scene": "C2" if big else f"{n} Gaussians, SH3, 640x368, :rgbd", "cases": []}


def make_views(V):
    cams, poses = [], []
    for v in range(V):
        R, t = view_pose(v, V, max_yaw_deg=yaw, max_shift=shift)
        cams.append(Camera(fx=sc.fx, fy=sc.fy, width=sc.width, height=sc.height, R=R, t=t))
        poses.append((R, t))
    vps = [make_vpixels(sc.width, sc.height, C, 100 + v) for v in range(V)]
    return cams, poses, vps


def oracle_sum(poses, vps):
    """Sum over the views of the CPU restatement's gradients (rank 0 computes, everybody receives)."""
    out = None
    if rank == 0:
        import parity as P
        o = P.oracle()
        total, amb = None, np.zeros(n, bool)
        for (R, t), vp in zip(poses, vps):
            _, ocam = P.cameras(sc, R=R, t=t)
            _, st = o.forward(sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, mode=mode, sh_degree=deg,
                              ambig_rel=P.AMBIG_REL)
            g = o.backward(vp, sc.means, sc.shs, sc.opacities, sc.scales, sc.rotations, ocam, st, mode=mode, sh_degree=deg)
            amb |= st.ambiguous_g != 0
            total = {k: g[k].astype(np.float64) for k in KEYS} if total is None else {k: total[k] + g[k] for k in KEYS}
        out = ({k: torch.from_numpy(np.ascontiguousarray(v.reshape(n, -1))) for k, v in total.items()}, torch.from_numpy(amb))
    box = [out]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def rel_errs(views, ref, keep=None):
    errs = {}
    for k in KEYS:
        a, b = views[k].double().reshape(n, -1).cpu(), ref[k].double().reshape(n, -1).cpu()
        dd = (a - b).abs().amax(1) / b.abs().max().clamp_min(1e-30)
        errs[k] = float(dd[keep].max() if keep is not None else dd.max())
    return errs


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


ok = True
for V in ([world] if big else [world, min(16, 2 * world)]):
    cams, poses, vps = make_views(V)
    mine = views_for_rank(V, rank, world)
    vp_dev = {v: d(vps[v]) for v in mine}
    # path A: accumulate this rank's views into one table, then ONE NCCL all-reduce
    rast_a = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, device=dev)
    table = GradientTable(n, K, dev)

    def step_nccl():
        for j, v in enumerate(mine):
            rast_a._forward(params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None, None,
                            cams[v], deg, (0, 0, 0), None, None)
            rast_a._backward(vp_dev[v], params["means"], params["shs"], params["opac"], params["scales"], params["rots"],
                             None, None, cams[v], deg, (0, 0, 0), outs=table.outs(), accumulate=(j > 0))
        allreduce_gradients_(table)

    step_nccl(); torch.cuda.synchronize()
    ref_nccl = {k: v.clone() for k, v in table.outs().items()}
    # path B: one accumulator per view, one fused backward + exchange per batch
    rast_b = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, device=dev)
    fused = ViewBatchBackward(rast_b, n, K, cams)
    views = fused.step(params, vp_dev, deg)
    torch.cuda.synchronize()
    case = {"views": V, "views_per_rank": len(mine)}
    case["fused_vs_nccl_path"] = rel_errs(views, ref_nccl)
    ok = ok and max(case["fused_vs_nccl_path"].values()) < 1e-4
    if not big:
        ref, amb = oracle_sum(poses, vps)
        keep = ~amb
        case["fused_vs_oracle_sum"] = rel_errs(views, ref, keep)
        case["nccl_path_vs_oracle_sum"] = rel_errs(ref_nccl, ref, keep)
        case["ambiguous_gaussians_excluded"] = int(amb.sum())
        ok = ok and max(case["fused_vs_oracle_sum"].values()) <= 1e-4 and max(case["nccl_path_vs_oracle_sum"].values()) <= 1e-4
    if world > 1:  # reduce-scatter form: every rank keeps the reduced rows of its own slice only
        rast_c = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, device=dev)
        fs = ViewBatchBackward(rast_c, n, K, cams, scatter_only=True)
        fs.table_flat.fill_(float("nan"))  # rows outside the slice must stay untouched
        sv = fs.step(params, vp_dev, deg)
        torch.cuda.synchronize()
        lo, hi = fs.slice_rows()
        e = max(float((sv[k][lo:hi].double() - views[k][lo:hi].double()).abs().max() / views[k].double().abs().max().clamp_min(1e-30))
                for k in KEYS)
        untouched = all(bool(torch.isnan(sv[k][:lo]).all()) and bool(torch.isnan(sv[k][hi:]).all()) for k in KEYS)
        case["scatter_only_slice_vs_fused"] = e
        case["scatter_only_rows_outside_slice_untouched"] = untouched
        ok = ok and e <= 2e-5 and untouched
        del fs, rast_c
    case["ms_nccl_path"] = timeit(step_nccl)
    case["ms_fused"] = timeit(lambda: fused.step(params, vp_dev, deg))
    report["cases"].append(case)
    if rank == 0:
        print(json.dumps(case))
    del fused, rast_a, rast_b

flag = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
report["all_ranks_ok"] = bool(flag.item())
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = f"peers_check_w{world}" + ("_c2" if big else "") + ("_odd" if odd else "") + ".json"
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
    print(f"N={n} world={world}: all ranks ok = {bool(flag.item())}")
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
