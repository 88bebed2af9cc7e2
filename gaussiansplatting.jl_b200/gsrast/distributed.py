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


def view_owner(view: int, world: int) -> tuple[int, int]:
    """(rank that renders `view`, its index among that rank's views) under round-robin sharding."""
    return view % world, view // world


class ViewBatchBackward:
    """A batch of V views (V <= 16) sharded round-robin over the ranks — several views per rank, or all of them on one
    GPU — with ONE fused per-Gaussian backward + exchange per batch (csrc/backward_peers.cu, gsr_backward_gaussians_views).

    Every view gets its own moment accumulator (64 / 80 B per Gaussian, second moments in fp64).  After this rank's forwards
    and compositing backwards, rank r reduces ITS slice of the Gaussians over all V views and stores the finished rows into
    every rank's table: compute + reduce-scatter + all-gather in one kernel, once per batch.  Across GPUs the accumulators
    are published as fp32 EXCHANGE rows (48 / 64 B, gsr_export_accumulator) in symmetric memory, which the owner of a
    slice loads over NVLink.  On a single GPU (world == 1, no process group needed) the same kernel reads the accumulators
    directly and replaces V accumulating `gsr_backward` calls: the parameters are read once and the gradient table is
    written once instead of V read-modify-write passes.  Result == sum over the V views of ∇rasterize, same layout as
    `GradientTable`."""

    def __init__(self, rast, n: int, K: int, cameras: list, group=None, scatter_only: bool = False):
        """scatter_only: every rank ends with the reduced rows of its own Gaussian slice `slice_rows()` only (reduce-
        scatter semantics, for a Gaussian-sharded optimizer) instead of the full replicated table (all-reduce)."""
        self.rast, self.n, self.K, self.cameras = rast, n, K, list(cameras)
        self.scatter_only = bool(scatter_only)
        self.V = len(self.cameras)
        assert 1 <= self.V <= 16, "1..16 views per batch"
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.multi = multi
        self.group = (group if group is not None else dist.group.WORLD) if multi else None
        self.world = dist.get_world_size(self.group) if multi else 1
        self.rank = dist.get_rank(self.group) if multi else 0
        self.mine = views_for_rank(self.V, self.rank, self.world)
        self.af = 16 if rast.channels <= 6 else 20   # csrc/common.cuh acc_floats: the handle's accumulator rows
        self.ef = 12 if rast.channels <= 6 else 16   # csrc/common.cuh exchange_floats: what crosses NVLink
        dev = rast.device
        slots = (self.V + self.world - 1) // self.world  # views per rank (same on every rank: symmetric)
        per = 4 + 3 + 3 + 1 + 3 * K
        self.acc = [torch.empty(n * self.af, dtype=torch.float32, device=dev) for _ in range(slots)]
        if multi:
            import torch.distributed._symmetric_memory as symm_mem
            self.rows = symm_mem.empty(slots * n * self.ef, dtype=torch.float32, device=dev)
            self.h_gacc = symm_mem.rendezvous(self.rows, self.group)
            self.table_flat = symm_mem.empty(n * per, dtype=torch.float32, device=dev)
            self.h_table = symm_mem.rendezvous(self.table_flat, self.group)
            bases, self.table_ptrs = list(self.h_gacc.buffer_ptrs), list(self.h_table.buffer_ptrs)
            stride = n * self.ef * 4
            self.view_ptrs = [bases[view_owner(v, self.world)[0]] + view_owner(v, self.world)[1] * stride
                              for v in range(self.V)]
            self.local_rows = [self.rows[j * n * self.ef:(j + 1) * n * self.ef] for j in range(slots)]
        else:
            self.table_flat = torch.empty(n * per, dtype=torch.float32, device=dev)
            self.h_gacc = self.h_table = None
            self.table_ptrs = [self.table_flat.data_ptr()]
            self.view_ptrs = [self.acc[v].data_ptr() for v in range(self.V)]
        if self.scatter_only:
            self.table_ptrs = [p if r == self.rank else 0 for r, p in enumerate(self.table_ptrs)]
        self.views, off = {}, 0
        for name, s_ in SEGMENTS:
            width = 3 * K if s_ is None else s_
            v = self.table_flat[off:off + n * width]
            self.views[name] = v.view(n, K, 3) if s_ is None else v.view(n, width)
            off += n * width

    def close(self):
        """Detach the rasterizer from this object's accumulators (it would otherwise keep a pointer into memory that is
        freed with this object)."""
        r, self.rast = self.rast, None
        if r is not None and getattr(r, "_h", None):
            try:
                r.set_accumulator(None)
            except Exception:
                pass

    def __del__(self):
        self.close()

    def slice_rows(self) -> tuple[int, int]:
        """[lo, hi): the Gaussians whose gradient rows this rank reduces (the kernel's slicing: 64-aligned chunks)."""
        chunk = (self.n + self.world - 1) // self.world
        chunk = (chunk + 63) // 64 * 64
        lo = min(self.rank * chunk, self.n)
        return lo, min(lo + chunk, self.n)

    def step(self, params: dict, vpixels: dict, sh_degree: int, background=(0.0, 0.0, 0.0), images: dict | None = None):
        """`vpixels[v]` = cotangent of view v (needed for this rank's views only).  Returns the table views."""
        r = self.rast
        for j, v in enumerate(self.mine):
            r.set_accumulator(self.acc[j])  # pointer swap: this view's accumulator
            img = r._forward(params["means"], params["shs"], params["opac"], params["scales"], params["rots"], None, None,
                             self.cameras[v], sh_degree, background, None, None)
            if images is not None:
                images[v] = img.clone()
            r.backward_render(vpixels[v], self.n, background)
            if self.multi:  # publish the finished accumulator as fp32 exchange rows in peer-mapped memory
                r.export_accumulator(self.n, self.local_rows[j])
        if self.h_gacc is not None:
            self.h_gacc.barrier(channel=0)   # every rank's rows are complete
        r.backward_gaussians_views(self.cameras, self.view_ptrs, self.world, self.rank, self.table_ptrs, params["means"],
                                   params["shs"], params["opac"], params["scales"], params["rots"], sh_degree,
                                   exchange_rows=self.multi)
        if self.h_table is not None:
            self.h_table.barrier(channel=0)  # every rank's table is complete (and nobody still reads my rows)
        return self.views


class PeerFusedBackward(ViewBatchBackward):
    """One view per rank (weak scaling): `ViewBatchBackward` with V == world.  `step` takes this rank's cotangent and
    returns (image of this rank's view, table views)."""

    def __init__(self, rast, n: int, K: int, cameras: list, group=None):
        super().__init__(rast, n, K, cameras, group)
        assert self.V == self.world, "one camera (view) per rank"

    def step(self, params: dict, vpixels, sh_degree: int, background=(0.0, 0.0, 0.0)):
        views = super().step(params, {self.rank: vpixels}, sh_degree, background)
        return self.rast.image, views
