#!/usr/bin/env python
"""tests/sanitize_workload.py: a few small frames through every kernel path (small + large + clipped triangles, all fragment
shaders, registry extension, scissor, streaming host path, interleaved bands), compared with the oracle.  Meant to be
run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tests/sanitize_workload.py
    compute-sanitizer --tool racecheck python tests/sanitize_workload.py
    compute-sanitizer --tool initcheck python tests/sanitize_workload.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ may use the oracle; tools/ may not
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import compare, gpu_render, oracle_render
from rusterizer_b200 import mathx, scenes

W, H = 160, 96
base = scenes.default_scene(1.0, width=W, height=H)
clip = scenes.clip_test_scene(0.4, width=W, height=H)
sph = scenes.sphere_scene(33, 17, width=W, height=H)
rng = np.random.default_rng(1)
t1 = scenes.Texture(rng.integers(0, 256, (16, 24, 3), dtype=np.uint8))
cases = [base, clip, sph,
         scenes.Scene("mixed", W, H, base.view, base.projection,
                      base.draws + [scenes.Draw(clip.draws[0].mesh, clip.draws[0].world, scenes.FS_COLOR),
                                    scenes.Draw(sph.draws[0].mesh, mathx.translate(1.0, -0.5, 1.0), scenes.fs_with_texture(scenes.FS_TEXTURE_BLEND, 1)),
                                    scenes.Draw(scenes.centered_quad(9.0), mathx.translate(0.0, 0.0, 3.0), scenes.FS_DEBUG)],
                      base.texture, [t1], scissor=(5, 3, 150, 90)),
         scenes.overdraw_scene(nx=40, ny=20, width=W, height=H),
         # wide target: the clipped quad's queue items split into column chunks and several slabs, interior tiles carry
         # ENTRY_FULL, and the frame's tile counters are zeroed by the dedicated kernel (tiny first vertex grid)
         scenes.fullscreen_quad_scene(1408, 272),
         scenes.near_clip_scene(40, 12, 640, 208)]
for sc in cases:
    for dbg in (True, False):
        msgs = compare(oracle_render(sc), gpu_render(sc, debug=dbg), check_samples=dbg)
        assert not msgs, (sc.name, msgs)
    print("ok", sc.name, flush=True)
# interleaved bands + row range
from rusterizer_b200.render import Renderer
o = oracle_render(sph)
full = np.zeros_like(o["fb"])
for rank in range(2):
    r = Renderer(W, H); r.uniforms().bind_texture(0, sph.texture); r.set_row_interleave(1, rank, 2)
    fb = gpu_render(sph, debug=False, renderer=r)["fb"]
    own = (np.arange(H) // 16) % 2 == rank
    full[own] = fb[own]; r.close()
assert np.array_equal(full, o["fb"]); print("ok interleave", flush=True)
# streaming host path
import torch
r = Renderer(W, H); r.uniforms().bind_texture(0, sph.texture)
b = r.uniforms().write_block(); b.projection, b.view, b.world = sph.projection, sph.view, sph.draws[0].world
m = sph.draws[0].mesh
pos, att, idx = (torch.from_numpy(a).pin_memory() for a in (m.vertices, m.attributes, m.indices.view(np.int32)))
outs = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
for k in range(4):
    r.render_arrays(pos.data_ptr(), att.data_ptr(), m.n_vertices, idx.data_ptr(), m.indices.size, 0, 0)
    r.framebuffer_host_async(outs[k % 2].data_ptr())
r.sync()
assert np.array_equal(outs[1].numpy().view(np.uint32), o["fb"]); r.close(); print("ok streaming", flush=True)
# runtime sample counts (generic N-sample tile kernel) and guard-band clipping
for n in (1, 2, 8):
    for sc in (scenes.default_scene(1.0, width=W, height=H), scenes.near_clip_scene(20, 10, W, H)):
        sc.msaa = n
        msgs = compare(oracle_render(sc), gpu_render(sc, debug=True))
        assert not msgs, (sc.name, n, msgs)
print("ok msaa 1/2/8", flush=True)
for sc in (scenes.clip_test_scene(0.4, width=W, height=H), scenes.near_clip_scene(20, 10, W, H)):
    sc.guard_band = 3.0
    msgs = compare(oracle_render(sc), gpu_render(sc, debug=True))
    assert not msgs, (sc.name, msgs)
print("ok guard band", flush=True)
# an async frame that overflows the initial tile bins: rz_sync grows and replays it
sc = scenes.overdraw_scene(8, 8, width=64, height=64)
sc.draws = [scenes.Draw(d.mesh, d.world, d.fs) for d in sc.draws] * 40
r = Renderer(sc.width, sc.height); r.uniforms().bind_texture(0, sc.texture)
out = torch.zeros((sc.height, sc.width), dtype=torch.int32).pin_memory()
scenes.render_scene(r, sc); r.framebuffer_host_async(out.data_ptr()); r.sync()
assert np.array_equal(out.numpy().view(np.uint32), oracle_render(sc)["fb"]); r.close(); print("ok async replay", flush=True)
