"""View-sharded data parallelism for the rasterizer hot path.

The reference has no batching and no multi-GPU code (one camera per `step!`, src/training.jl:587-591); the
semantics here are defined as: batch gradient = sum over views of the per-view `∇rasterize` gradients, i.e.
what B sequential reference pullbacks accumulate to.  One process per GPU (torchrun); Gaussian parameters are
replicated; rank r renders views r, r+G, ... into ONE flat gradient table (the kernels accumulate), and a single
NCCL all-reduce over NVLink/NVSwitch sums the tables.  Densification statistics reduce with sum / max.

Host-side logic only — the kernels live in libgsrast.so.  Covered on CPU by tests/test_distributed_cpu.py (gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# segment order keeps `vrot` (float4, 128-bit stores) 16-byte aligned at offset 0 for any N
SEGMENTS = (("vrot", 4), ("vmeans", 3), ("vscales", 3), ("vopacities", 1), ("vshs", None))


def views_for_rank(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin view sharding: every view is rendered by exactly one rank."""
    return list(range(rank, n_views, world))


class GradientTable:
    """One contiguous fp32 buffer holding the 3+3K+1+3+4 floats per Gaussian that `∇rasterize` returns
    (rasterizer.jl:549), so that the cross-GPU reduction is a single collective."""

    def __init__(self, n: int, K: int, device):
        self.n, self.K = n, K
        per = sum(s if s is not None else 3 * K for _, s in SEGMENTS)
        self.flat = torch.zeros(n * per, dtype=torch.float32, device=device)
        self.views, off = {}, 0
        for name, s in SEGMENTS:
            width = 3 * K if s is None else s
            v = self.flat[off:off + n * width]
            self.views[name] = v.view(n, K, 3) if s is None else v.view(n, width)
            off += n * width
        assert self.views["vrot"].data_ptr() % 16 == 0

    def outs(self) -> dict:
        return dict(self.views)

    def zero_(self):
        self.flat.zero_()
        return self

    def nbytes(self) -> int:
        return self.flat.numel() * 4


def allreduce_gradients_(table: GradientTable, group=None, async_op: bool = False):
    """Sum the per-rank gradient tables in place (NCCL on GPU, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(table.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return None


def allreduce_stats_(max_radii: torch.Tensor, accum: torch.Tensor, denom: torch.Tensor, group=None):
    """Densification statistics of `_update_stats!` (strategy.jl:118-136) across ranks: sums and a max."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(denom, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group)


def render_view_batch(rast, params: dict, cameras: list, vpixels: list, table: GradientTable, sh_degree: int,
                      background=(0.0, 0.0, 0.0), rank: int = 0, world: int = 1, group=None, images: list | None = None):
    """Forward + backward of this rank's share of a view batch, gradients accumulated into `table`, then the
    all-reduce.  `params` holds activated device tensors: means, shs, opac, scales, rots."""
    mine = views_for_rank(len(cameras), rank, world)
    if not mine:
        table.zero_()
    for j, v in enumerate(mine):
        img = rast._forward(params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None,
                            None, cameras[v], sh_degree, background, None, None)
        if images is not None:
            images.append((v, img.clone()))
        rast._backward(vpixels[v], params["means"], params["shs"], params["opac"], params["scales"], params["rots"],
                       None, None, cameras[v], sh_degree, background, outs=table.outs(), accumulate=(j > 0))
    allreduce_gradients_(table, group=group)
    return mine


class PeerFusedBackward:
    """Per-Gaussian backward fused with the cross-GPU gradient reduction over NVLink peer memory
    (csrc/backward_peers.cu): replaces `backward_gaussians + all_reduce(table)`.

    Every rank keeps its moment accumulator (48/64 B per Gaussian) and its gradient table in symmetric memory
    (torch.distributed._symmetric_memory).  After the local compositing backward, rank r loads all ranks' accumulator
    rows for ITS slice of Gaussians over NVLink, applies each view's ∇project / ∇SH chain, sums, and stores the
    reduced rows into every rank's table.  Result == the all-reduced table (same layout as `GradientTable`)."""

    def __init__(self, rast, n: int, K: int, cameras: list, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        assert len(cameras) == self.world, "one camera (view) per rank"
        self.rast, self.n, self.K, self.cameras = rast, n, K, cameras
        af = 12 if rast.channels <= 6 else 16
        dev = rast.device
        self.gacc = symm_mem.empty(n * af, dtype=torch.float32, device=dev)
        self.h_gacc = symm_mem.rendezvous(self.gacc, self.group)
        per = 4 + 3 + 3 + 1 + 3 * K
        self.table_flat = symm_mem.empty(n * per, dtype=torch.float32, device=dev)
        self.h_table = symm_mem.rendezvous(self.table_flat, self.group)
        self.gacc_ptrs = list(self.h_gacc.buffer_ptrs)
        self.table_ptrs = list(self.h_table.buffer_ptrs)
        rast.set_accumulator(self.gacc)
        self.views, off = {}, 0
        for name, s in SEGMENTS:
            width = 3 * K if s is None else s
            v = self.table_flat[off:off + n * width]
            self.views[name] = v.view(n, K, 3) if s is None else v.view(n, width)
            off += n * width

    def step(self, params: dict, vpixels, sh_degree: int, background=(0.0, 0.0, 0.0)):
        """forward + compositing backward of this rank's view, then the fused reduce; returns the table views."""
        r, cam = self.rast, self.cameras[self.rank]
        img = r._forward(params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None, None, cam,
                         sh_degree, background, None, None)
        r.backward_render(vpixels, self.n, background)
        self.h_gacc.barrier(channel=0)    # every rank's accumulator is complete
        r.backward_gaussians_peers(self.world, self.rank, self.cameras, self.gacc_ptrs, self.table_ptrs, params["means"],
                                   params["shs"], params["opac"], params["scales"], params["rots"], sh_degree)
        self.h_table.barrier(channel=0)   # every rank's table is complete (and nobody still reads my accumulator)
        return img, self.views
