"""Host-copy ceiling of the box: pinned H2D alone, D2H alone and both at once (two streams) on every visible GPU at the
same time.  The e2e leg of bench.py moves 1 GB in and 6 GB out per scene; this says what the link and the host allow.
python profiles/microbench/pcie_duplex.py [MB]      (under torchrun: one rank per GPU, reports the per-rank rates)"""
import json
import os
import sys

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = mb * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a)
    s2.wait_event(a)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return (int(h2d) + int(d2h)) * reps * n / ms / 1e6


run(True, True, 1)
res = {"rank": rank, "world": world, "mb": mb, "h2d_gbs": run(True, False), "d2h_gbs": run(False, True),
       "both_total_gbs": run(True, True)}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print(json.dumps({"per_rank": out, "sum_both_gbs": sum(r["both_total_gbs"] for r in out)}))
    dist.destroy_process_group()
else:
    print(json.dumps(res))
