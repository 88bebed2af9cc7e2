"""torchrun helper: per-stage times of one fused C2 step (library event timers) with the view accumulators in peer-mapped
memory vs staged through ordinary device memory.  Measurement infrastructure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
import numpy as np, torch, torch.distributed as dist
from gsrast import Camera, GaussianRasterizer
from gsrast.distributed import ViewBatchBackward
from gsrast.synthetic import make_config, make_vpixels, view_pose
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sc = make_config("C2")
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(means=sc.means, shs=sc.shs, opac=sc.opacities.reshape(-1, 1), scales=sc.scales, rots=sc.rotations).items()}
cams = [Camera(fx=sc.fx, fy=sc.fy, width=sc.width, height=sc.height, R=R, t=t) for R, t in (view_pose(v, world, max_yaw_deg=2.0, max_shift=0.1) for v in range(world))]
vp = torch.from_numpy(make_vpixels(sc.width, sc.height, 5, 1002)).to(dev)
for stage in (True,):
    rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd", device=dev)
    f = ViewBatchBackward(rast, sc.n, 16, cams)
    for _ in range(5): f.step(d, {rank: vp}, 3)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): f.step(d, {rank: vp}, 3)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 30], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rast.profile(True)
    acc = {}
    for _ in range(10):
        f.step(d, {rank: vp}, 3); torch.cuda.synchronize()
        for k, v in rast.stage_times_ms().items(): acc[k] = acc.get(k, 0) + v / 10
    rast.profile(False)
    if rank == 0:
        print("exchange-row accumulators: ms/step", round(float(t), 4), {k: round(v, 4) for k, v in acc.items() if k in ("render_bwd", "gauss_bwd", "zero_grads", "render_fwd")})
    del f, rast
dist.barrier(); dist.destroy_process_group()
