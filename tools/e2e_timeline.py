import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
import numpy as np, torch, time
from gsrast import Camera, GaussianRasterizer, _lib
from gsrast.synthetic import make_config, make_vpixels
sc = make_config("C2"); n, K = sc.n, 16
cam = Camera(fx=sc.fx, fy=sc.fy, width=sc.width, height=sc.height)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
host = dict(means=pin(sc.means), shs=pin(sc.shs), opac=pin(sc.opacities.reshape(-1, 1)), scales=pin(sc.scales), rots=pin(sc.rotations))
vp = pin(make_vpixels(sc.width, sc.height, 5, 1002))
out = dict(image=torch.empty((sc.height, sc.width, 5)).pin_memory(), vmeans=torch.empty((n, 3)).pin_memory(), vshs=torch.empty((n, K, 3)).pin_memory(),
           vopacities=torch.empty((n, 1)).pin_memory(), vscales=torch.empty((n, 3)).pin_memory(), vrot=torch.empty((n, 4)).pin_memory())
rast = GaussianRasterizer(width=sc.width, height=sc.height, mode="rgbd")
if os.environ.get("E2E_SIDE_STREAM"):
    torch.cuda.set_stream(torch.cuda.Stream())
for _ in range(3): rast.forward_backward_host(host, vp, cam, 3, out=out, wait=False)
rast.host_wait(); rast.profile(True)
t0 = time.perf_counter(); cpu = []
for i in range(8):
    a = time.perf_counter(); rast.forward_backward_host(host, vp, cam, 3, out=out, wait=False); cpu.append((a - t0, time.perf_counter() - t0))
rast.host_wait(); t1 = time.perf_counter()
arr = (C.c_float * 12)(); _lib.lib().gsr_host_timeline(rast._h, arr)
print("8 steps wall", (t1 - t0) * 1e3 / 8, "ms/step")
print("cpu submit [start,end] ms:", [(round(a * 1e3, 2), round(b * 1e3, 2)) for a, b in cpu])
for k in range(2):
    v = [round(arr[6 * k + j], 2) for j in range(6)]
    print(f"slot{k}: h2d {v[0]}->{v[1]} ({v[1]-v[0]:.2f}) compute {v[2]}->{v[3]} ({v[3]-v[2]:.2f}) d2h {v[4]}->{v[5]} ({v[5]-v[4]:.2f})")
