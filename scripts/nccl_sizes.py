"""All-reduce time vs message size on this box (torchrun, one rank per GPU): the config-4 evaluation moves 220 MB per
all-reduce; profiles/r2_bench_8gpu_cfg4.json shows 11.7 ms between the end of rank 0's compute and the end of the
step.  This separates transfer time from waiting for the slowest rank, and tests chunked collectives."""
import json
import os
import sys
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
out = []
for mb in (8, 35, 64, 128, 220, 256):
    n = mb * 1000 * 1000 // 4
    t = torch.ones(n, dtype=torch.float32, device="cuda")
    for chunks in (1, 4):
        parts = list(t.chunk(chunks))
        for _ in range(3):
            for p in parts:
                dist.all_reduce(p)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            for p in parts:
                dist.all_reduce(p)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(tt.item())
            out.append(dict(MB=mb, chunks=chunks, ms=ms, busbw_GBps=mb / 1e3 * 2 * (world - 1) / world / (ms / 1e3)))
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
