"""Host <-> device copy ceiling with N ranks copying at once (one rank per GPU, pinned buffers, no compute): the bound
of bench.py's e2e leg.  Run under torchrun like bench.py; rank 0 prints one JSON line.
  h2d_only     every rank copies 2.4 GB host -> device repeatedly
  both         every rank additionally copies 0.72 GB device -> host on a second stream (the e2e leg's ratio)"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
h_in = torch.empty(600_000_000, dtype=torch.float32).pin_memory()
h_out = torch.empty(180_000_000, dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ('h2d_only', 'both'):
    for rep in range(2):      # first repetition = warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 4
        for _ in range(n):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if mode == 'both':
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    gb = n * (h_in.numel() * 4 + (h_out.numel() * 4 if mode == 'both' else 0)) / 1e9
    res[mode] = {'per_rank_gbs': gb / dt, 'aggregate_gbs': world * gb / dt}
if int(os.environ.get('RANK', '0')) == 0:
    print(json.dumps({'n_gpus': world, **res}))
if world > 1:
    dist.destroy_process_group()
