#!/usr/bin/env python
"""tools/tiles_probe.py (under torchrun): per-rank, per-frame device and host times of the tile-row mode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from rusterizer_b200 import scenes, sharding
from rusterizer_b200.render import Renderer
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sc = scenes.sphere_scene(1001, 501, width=8192, height=8192)
r = Renderer(sc.width, sc.height, device=local)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); r.set_stream(st.cuda_stream)
r.uniforms().bind_texture(0, sc.texture)
m = r.upload(sc.draws[0].mesh)
b = r.uniforms().write_block(); b.projection = sc.projection; b.world = sc.draws[0].world; b.view = sc.view
mode = sys.argv[1] if len(sys.argv) > 1 else "peer"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pf = None
r0, r1 = sharding.row_range(rank, world, sc.height, 16)
if mode == "peer":
    pf = sharding.PeerFrame(r)
else:
    r.set_row_range(r0, r1)
    per = sharding.strip_rows(sc.height, world, 16)
    strip = torch.empty((per, sc.width), dtype=torch.int32, device="cuda")
    full = torch.empty((per * world, sc.width), dtype=torch.int32, device="cuda")
r.render(m, 0, 0); r.framebuffer_device()
def frame():
    r.render(m, 0, 0)
    if pf: pf.finish_frame(); pf.release()
    elif mode == "nccl": r.framebuffer_async(strip.data_ptr()); dist.all_gather_into_tensor(full, strip)
    else: r.framebuffer_async(strip.data_ptr())
for use_flush in (False, True):
    for _ in range(3): frame()
    r.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    host = []
    t_all = time.perf_counter()
    for a, e in ev:
        if use_flush: flush.zero_()
        t = time.perf_counter(); a.record(st); frame(); e.record(st); host.append((time.perf_counter() - t) * 1e3)
    r.sync(); torch.cuda.synchronize()
    t_all = (time.perf_counter() - t_all) * 1e3
    print(f"rank {rank} mode {mode} flush {use_flush}: dev ms {[round(a.elapsed_time(e), 3) for a, e in ev]} host ms {[round(h, 3) for h in host]} wall {t_all:.2f}", flush=True)
    dist.barrier()
if pf: pf.close()
dist.destroy_process_group()
