"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: view sharding covers every view once, the flat
gradient table keeps `vrot` 16-byte aligned and round-trips its views, and the all-reduce of tables / stats
gives sum / max semantics.  The kernels themselves need a GPU (tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
    from gsrast.distributed import GradientTable, allreduce_gradients_, allreduce_stats_, views_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, K, n_views = 1001, 16, 8
    table = GradientTable(n, K, "cpu")
    mine = views_for_rank(n_views, rank, world)
    # "render": view v contributes (v+1) to every gradient entry of segment s scaled by its index
    for j, v in enumerate(mine):
        for si, (name, t) in enumerate(table.outs().items()):
            contrib = torch.full_like(t, float((v + 1) * (si + 1)))
            if j == 0:
                t.copy_(contrib)      # accumulate=0 overwrites
            else:
                t.add_(contrib)       # accumulate=1 adds
    allreduce_gradients_(table)
    expect = sum(v + 1 for v in range(n_views))
    ok = all(bool((t == expect * (si + 1)).all()) for si, (name, t) in enumerate(table.outs().items()))
    mr = torch.tensor([rank + 3, 7 - rank], dtype=torch.int32)
    acc, den = torch.tensor([1.0 + rank, 2.0]), torch.tensor([1.0, float(rank)])
    allreduce_stats_(mr, acc, den)
    ok = ok and mr.tolist() == [world + 2, 7] and acc.tolist() == [sum(1.0 + r for r in range(world)), 2.0 * world]
    ok = ok and den.tolist() == [float(world), float(sum(range(world)))]
    q.put((rank, mine, ok, table.nbytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharding_and_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    views = sorted(v for _, mine, _, _ in res for v in mine)
    assert views == list(range(8))                        # every view rendered exactly once
    assert all(ok for _, _, ok, _ in res)
    assert res[0][3] == 1001 * 59 * 4                     # 59 floats per Gaussian at K=16


def test_gradient_table_layout():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
    from gsrast.distributed import GradientTable, views_for_rank
    for n in (1, 3, 1000, 1001):
        for K in (1, 4, 9, 16):
            t = GradientTable(n, K, "cpu")
            o = t.outs()
            assert o["vrot"].shape == (n, 4) and o["vshs"].shape == (n, K, 3) and o["vopacities"].shape == (n, 1)
            assert o["vrot"].data_ptr() % 16 == 0
            assert t.flat.numel() == n * (3 + 3 * K + 1 + 3 + 4)
            o["vshs"].fill_(2.0)
            assert float(t.flat.sum()) == 2.0 * n * 3 * K  # views alias the flat buffer without overlap
    assert views_for_rank(8, 3, 4) == [3, 7] and views_for_rank(3, 3, 4) == [] and views_for_rank(8, 0, 1) == list(range(8))


def test_view_owner_matches_round_robin_sharding():
    """ViewBatchBackward's accumulator addressing: view v lives on rank v % world at slot v // world — exactly the
    j-th entry of views_for_rank(V, rank, world) — for every batch size up to the kernel's 16-view limit."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussiansplatting.jl_b200")]
    from gsrast.distributed import view_owner, views_for_rank
    for world in (1, 2, 3, 4, 8):
        for V in range(1, 17):
            seen = set()
            for r in range(world):
                for j, v in enumerate(views_for_rank(V, r, world)):
                    assert view_owner(v, world) == (r, j)
                    seen.add(v)
            assert seen == set(range(V))
            slots = (V + world - 1) // world
            assert all(view_owner(v, world)[1] < slots for v in range(V))
