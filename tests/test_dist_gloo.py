"""world_size-2 gloo test of the multi-GPU host logic (slab partition + result all-gather), run on CPU.  The
kernels are not involved: each rank fabricates its slab's `results` deterministically from the global index."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gik_b200.dist import all_gather_results, shard_bounds
    lo, hi = shard_bounds(n_total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float32)
    q = torch.stack([idx * 15 + c for c in range(15)])            # [15][n_local], value encodes (component, index)
    conv = (torch.arange(lo, hi) % 3 == 0).to(torch.uint8)
    qg, cg = all_gather_results(q, conv, n_total)
    full = torch.arange(n_total, dtype=torch.float32)
    ok = bool(torch.equal(qg, torch.stack([full * 15 + c for c in range(15)])) and
              torch.equal(cg, (torch.arange(n_total) % 3 == 0).to(torch.uint8)))
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])          # even split, and ragged slabs that need padding
def test_all_gather_results_world2(n_total):
    world = 2
    with mp.get_context("spawn").Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
