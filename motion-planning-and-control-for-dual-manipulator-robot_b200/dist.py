"""Multi-GPU: plain data partitioning of the batch (every IK problem is independent) -- one
process per GPU, contiguous slab per rank, no collective on the data path.  When every rank needs the whole answer:

  * `all_gather_results`  -- ONE NCCL all-gather of (q, converged) after the kernel (the baseline), or
  * `SymmetricResults` + `solve_sharded_fused` -- solve + all-gather in ONE kernel launch: lanes store results into this
    rank's own array, pusher warps of the same kernel copy finished 1024-problem chunks into every peer's array through
    NVLink peer mappings (torch symmetric memory) while the other blocks keep solving, a tail kernel copies the chunks
    still in flight at the end; no collective follows, only a cross-rank barrier."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous slab [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_results(q_soa: torch.Tensor, conv: torch.Tensor, n_total: int, group=None):
    """q_soa [nq][n_local], conv u8 [n_local] of this rank's slab (shard_bounds) -> (q [nq][n_total] view,
    conv [n_total]) on every rank.  One all_gather_into_tensor per output; slabs are padded to the largest
    slab so the collective is a single fixed-size call."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nq = q_soa.shape[0]
    n_max = -(-n_total // world)
    lo, hi = shard_bounds(n_total, rank, world)
    assert hi - lo == q_soa.shape[1] == conv.shape[0], "local slab does not match shard_bounds"
    if q_soa.shape[1] != n_max:
        qp = q_soa.new_zeros((nq, n_max)); qp[:, : hi - lo] = q_soa
        cp = conv.new_zeros((n_max,)); cp[: hi - lo] = conv
    else:
        qp, cp = q_soa.contiguous(), conv.contiguous()
    qg = q_soa.new_empty((world, nq, n_max))
    cg = conv.new_empty((world, n_max))
    # flat views: the gloo backend (CPU tests) only accepts 1-D concatenation; NCCL takes either
    dist.all_gather_into_tensor(qg.view(-1), qp.reshape(-1), group=group)
    dist.all_gather_into_tensor(cg.view(-1), cp.reshape(-1), group=group)
    if n_total == n_max * world:
        return qg.permute(1, 0, 2).reshape(nq, n_total), cg.reshape(n_total)
    qs, cs = [], []
    for r in range(world):
        a, b = shard_bounds(n_total, r, world)
        qs.append(qg[r, :, : b - a]); cs.append(cg[r, : b - a])
    return torch.cat(qs, dim=1), torch.cat(cs)


def solve_sharded(solver, q_init_soa, pose_soa, n_total: int, *, gather=True, group=None, **kw):
    """Each rank solves its slab (`q_init_soa`, `pose_soa` hold ONLY this rank's columns) and, if `gather`,
    all ranks receive the full (q, converged)."""
    q, conv, iters, resid = solver.solve_soa(q_init_soa, pose_soa, **kw)
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return q, conv, iters, resid
    qg, cg = all_gather_results(q, conv, n_total, group)
    return qg, cg, iters, resid


class SymmetricResults:
    """Result arrays q [nq][n_total] and converged u8 [n_total] allocated as torch symmetric memory: every rank maps
    every other rank's copy, and the fused launch (gik_solve_scatter_*) fills all of them.
    Reuse: call `barrier()` again before the next launch overwrites arrays a peer may still be reading."""

    def __init__(self, nq: int, n_total: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.n_total, self.nq = int(n_total), int(nq)
        self.q = symm_mem.empty((nq, n_total), dtype=dtype, device=device)
        self.conv = symm_mem.empty((n_total,), dtype=torch.uint8, device=device)
        self._hq = symm_mem.rendezvous(self.q, group)
        self._hc = symm_mem.rendezvous(self.conv, group)
        self.q_ptrs = [int(p) for p in self._hq.buffer_ptrs]
        self.conv_ptrs = [int(p) for p in self._hc.buffer_ptrs]
        self.rank, self.world = self._hq.rank, self._hq.world_size

    def barrier(self):
        """Stream-ordered cross-rank barrier: once this stream has passed it, every rank's kernel launched before its
        own barrier has completed, i.e. all peer stores into this rank's arrays have landed."""
        self._hq.barrier()


def solve_sharded_fused(solver, q_init_soa, pose_soa, results: SymmetricResults, *, group=None, **kw):
    """Each rank solves its slab and the same launch copies the results into every rank's `results` arrays; returns
    (results.q, results.conv, iters_local, resid_local) after the barrier."""
    lo, hi = shard_bounds(results.n_total, results.rank, results.world)
    assert hi - lo == q_init_soa.shape[1], "local slab does not match shard_bounds"
    iters, resid = solver.solve_scatter_soa(q_init_soa, pose_soa, results.q_ptrs, results.conv_ptrs, results.n_total,
                                            lo, **kw)
    results.barrier()
    return results.q, results.conv, iters, resid
