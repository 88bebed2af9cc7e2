"""torchrun test (>= 2 GPUs): the peer-fused backward equals backward_gaussians + NCCL all-reduce; timing of both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
import numpy as np, torch, torch.distributed as dist
from gsrast import Camera, GaussianRasterizer
from gsrast.distributed import GradientTable, PeerFusedBackward, allreduce_gradients_
from gsrast.synthetic import make_scene, make_config, make_vpixels, view_pose

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
big = len(sys.argv) > 1 and sys.argv[1] == "c2"
sc = make_config("C2") if big else make_scene(50_001 if len(sys.argv) > 2 else 50_000, 3, 640, 368, 7)
mode, C, deg, K = "rgbd", 5, 3, 16
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = dict(means=d(sc.means), shs=d(sc.shs), opac=d(sc.opacities.reshape(-1, 1)), scales=d(sc.scales), rots=d(sc.rotations))
yaw, shift = (2.0, 0.1) if big else (20.0, 1.0)
cams = []
for r in range(world):
    R, t = view_pose(r, world, max_yaw_deg=yaw, max_shift=shift)
    cams.append(Camera(fx=sc.fx, fy=sc.fy, width=sc.width, height=sc.height, R=R, t=t))
vpix = d(make_vpixels(sc.width, sc.height, C, 100 + rank))
n = sc.n

# baseline: per-rank backward + NCCL all-reduce
rast_a = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, device=dev)
table = GradientTable(n, K, dev)
def step_nccl():
    rast_a._forward(params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None, None, cams[rank], deg, (0, 0, 0), None, None)
    rast_a._backward(vpix, params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None, None, cams[rank], deg, (0, 0, 0), outs=table.outs())
    allreduce_gradients_(table)
step_nccl(); torch.cuda.synchronize()
ref = {k: v.clone() for k, v in table.outs().items()}
gm_ref = rast_a.gstate.grad_means2d.clone()

# peer-fused
rast_b = GaussianRasterizer(width=sc.width, height=sc.height, mode=mode, device=dev)
fused = PeerFusedBackward(rast_b, n, K, cams)
img, views = fused.step(params, vpix, deg)
torch.cuda.synchronize()
ok = True
for k in ("vrot", "vmeans", "vscales", "vopacities", "vshs"):
    a, b = views[k].double(), ref[k].double()
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    if rank == 0: print(f"{k}: rel err vs NCCL path {err:.3e}")
    ok = ok and err < 2e-4
gerr = float((rast_b.gstate.grad_means2d - gm_ref).abs().max() / gm_ref.abs().max())
ok = ok and gerr < 1e-4
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_nccl = timeit(step_nccl)
t_fused = timeit(lambda: fused.step(params, vpix, deg))
flag = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"grad_means2d rel err {gerr:.3e}")
    print(f"N={n} world={world}: NCCL path {t_nccl:.3f} ms/step, peer-fused {t_fused:.3f} ms/step; all ranks ok = {bool(flag.item())}")
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
