#!/usr/bin/env python
"""tools/overlap_probe.py: throughput of back-to-back C2 frames with 1..3 contexts (each on its own stream)
alternating on ONE GPU, mesh copies cycled so that the inputs exceed L2 (no flush inside the timed region)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rusterizer_b200 import scenes
from rusterizer_b200.render import Renderer

sc = scenes.sphere_scene(1001, 501)
mesh = sc.draws[0].mesh
K = 120
for nctx in (1, 2, 3, 4):
    rs, meshes = [], []
    for c in range(nctx):
        r = Renderer(sc.width, sc.height)
        r.uniforms().bind_texture(0, sc.texture)
        b = r.uniforms().write_block(); b.projection = sc.projection; b.view = sc.view; b.world = sc.draws[0].world
        rs.append(r)
        meshes.append([r.upload(mesh) for _ in range(6 // nctx + 1)])
    for r, ms in zip(rs, meshes):  # warm-up / sizing
        r.render(ms[0], 0, 0); r.framebuffer_device()
        for m in ms:
            r.render(m, 0, 0); r.framebuffer_async()
        r.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        r = rs[i % nctx]; ms = meshes[i % nctx]
        r.render(ms[(i // nctx) % len(ms)], 0, 0)
        r.framebuffer_async()
    for r in rs:
        r.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"ctxs {nctx}: {dt / K * 1e6:.1f} us/frame wall ({1e6 / (dt / K) / 1e6 * 1.0:.2f} Gtris/s)")
    for r in rs:
        r.close()
