"""Feasibility probe (2+ GPUs, torchrun): symmetric memory rendezvous, peer pointers, barrier, peer read bandwidth."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 12_000_000
buf = symm_mem.empty(n, dtype=torch.float32, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
buf.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (n,), torch.float32)
torch.cuda.synchronize()
x = pt.clone(); torch.cuda.synchronize()
ok = bool((x == float(peer + 1)).all())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.copy_(pt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
if rank == 0:
    print("buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None)
    print("peer read ok", ok, f"{n*4/ms/1e6:.1f} GB/s peer copy")
    e0.record()
    for _ in range(20): hdl.barrier(channel=0)
    e1.record(); torch.cuda.synchronize(); print("barrier us", e0.elapsed_time(e1) / 20 * 1e3)
else:
    for _ in range(20): hdl.barrier(channel=0)
dist.barrier(); dist.destroy_process_group()
