"""Aggregate device->host bandwidth with N ranks copying at the same time (run under torchrun, one rank per GPU):
the host-side ceiling of the Gymnasium face, which moves 7 MB per step and rank into pinned host memory.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw_ranks.py
"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 64 << 20
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
    for direction in ('d2h', 'h2d'):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank < active:
            for _ in range(3):
                (h.copy_(d, non_blocking=True) if direction == 'd2h' else d.copy_(h, non_blocking=True))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 40
            for _ in range(reps):
                (h.copy_(d, non_blocking=True) if direction == 'd2h' else d.copy_(h, non_blocking=True))
            torch.cuda.synchronize()
            gbs = reps * n / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbs], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        if rank == 0:
            print(f'{direction}: {active} rank(s) copying 64 MiB pinned buffers concurrently: aggregate {float(t):.1f} GB/s, {float(t) / active:.1f} GB/s per rank', flush=True)
if world > 1:
    dist.destroy_process_group()
