"""PCIe floor of the end-to-end number: 278 MB (one C2 step's parameters + cotangent) host->device, device->host and
both at once, from pinned memory.  Under torchrun every rank runs the same copies CONCURRENTLY on its own GPU (barrier
in front), which is what `e2e` at N GPUs competes with: the ranks share the host's memory system and PCIe root ports.

    python tools/pcie_probe.py                                   # one GPU
    torchrun --nproc-per-node N tools/pcie_probe.py              # N GPUs at once -> gpurun_out/pcie_probe_n<N>.json
"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 278_000_000 // 4
h_in = torch.empty(n).pin_memory(); h_out = torch.empty(n).pin_memory()
d_in = torch.empty(n, device="cuda"); d_out = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


res = {"n_gpus": world, "bytes": 4 * n}
for name, u, d in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
    run(u, d, 2)
    dt = run(u, d)
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    res[name] = {"ms_per_278MB_slowest_rank": round(dt * 1e3, 3), "GBps_per_gpu_per_direction": round(0.278 / dt, 1),
                 "GBps_aggregate_per_direction": round(world * 0.278 / dt, 1)}
    if rank == 0:
        print(f"{name}: {dt*1e3:.2f} ms per 278 MB on the slowest of {world} GPU(s) -> {0.278/dt:.1f} GB/s per GPU and direction")
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/pcie_probe_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
