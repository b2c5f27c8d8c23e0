#!/usr/bin/env python
"""tools/nccl_probe.py (under torchrun): all_gather_into_tensor bandwidth for the tile-row strips of an 8192^2 frame."""
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size()
n = 8192 * 8192 // w
strip = torch.ones(n, dtype=torch.int32, device="cuda"); full = torch.empty(n * w, dtype=torch.int32, device="cuda")
for _ in range(5): dist.all_gather_into_tensor(full, strip)
torch.cuda.synchronize(); dist.barrier()
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dist.all_gather_into_tensor(full, strip)
    b.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"all_gather {n*4/2**20:.0f} MiB/rank x{w}: {a.elapsed_time(b)/10:.3f} ms")
dist.destroy_process_group()
