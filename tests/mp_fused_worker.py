"""torchrun worker (2+ GPUs): the fused solve + all-gather (kernel stores into every rank's symmetric-memory result
arrays over NVLink) must equal solve + NCCL all_gather_into_tensor, on every rank.  Launched by
tests/test_gpu_parity.py::test_fused_all_gather_two_gpus."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import gik_b200
    from gik_b200 import dist as gdist
    solver = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
    all_ok = 1
    for n_total, dtype in ((100003, torch.float32),       # ragged slabs, unaligned offsets: scalar chunk pushes
                           (1 << 21, torch.float32),      # aligned slabs: 16-byte chunk pushes, packed lane kernel
                           (40000, torch.float64)):
        lo, hi = gdist.shard_bounds(n_total, rank, world)
        g = torch.Generator(device=dev).manual_seed(5)
        pos = torch.tensor([0.2, -0.4, 0.93], device=dev) + torch.rand((n_total, 3), device=dev, generator=g) * \
            torch.tensor([0.4, 0.8, 0.47], device=dev)
        pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n_total, 9), pos], 1).t().contiguous().to(dtype)
        my_pose = pose[:, lo:hi].contiguous()
        q0 = torch.zeros((15, hi - lo), device=dev, dtype=dtype)
        # baseline: kernel + NCCL all-gather
        q, conv, _, _ = solver.solve_soa(q0, my_pose)
        qg, cg = gdist.all_gather_results(q, conv, n_total)
        # fused: finished chunks are pushed into every rank's arrays over NVLink by the solve kernel itself
        res = gdist.SymmetricResults(15, n_total, dtype, dev)
        res.q.fill_(-1); res.conv.fill_(7)
        res.barrier()
        qf, cf, _, _ = gdist.solve_sharded_fused(solver, q0, my_pose, res)
        torch.cuda.synchronize()
        ok = bool(torch.equal(qf, qg) and torch.equal(cf, cg))
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res.barrier()
        all_ok = min(all_ok, int(t.item()))
        if rank == 0:
            print(f"n_total={n_total} {dtype}: {'ok' if t.item() == 1 else 'MISMATCH'}", flush=True)
        del res
    t = torch.tensor([all_ok], device=dev)
    if rank == 0:
        print("FUSED_OK" if t.item() == 1 else "FUSED_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
