"""Shared test helpers: run a Scene through the CPU oracle and through the CUDA path."""
from __future__ import annotations

import numpy as np


def oracle_render(scene, fast=False):
    """Returns dict(fb, depth, color, owner, counters) from the CPU oracle."""
    from oracle.oracle import OracleRenderer, get_lib

    r = OracleRenderer(scene.width, scene.height, get_lib(fast=fast))
    if scene.texture is not None:
        r.bind_texture(0, scene.texture.texels)
        for k, t in enumerate(getattr(scene, "extra_textures", [])):
            r.bind_texture(k + 1, t.texels)
    r.write_block(view=scene.view, projection=scene.projection)
    if getattr(scene, "scissor", None):
        r.set_scissor(*scene.scissor)
    if getattr(scene, "msaa", 4) != 4:
        r.set_msaa(scene.msaa)
    if getattr(scene, "guard_band", 1.0) != 1.0:
        r.set_guard_band(scene.guard_band)
    for d in scene.draws:
        r.write_block(world=d.world)
        r.render(d.mesh.vertices, d.mesh.attributes, d.mesh.indices, 0, d.fs)
    out = dict(depth=r.depth_samples(), color=r.color_samples(), owner=r.owner_samples(), counters=r.counters())
    out["fb"] = r.framebuffer()
    r.close()
    return out


def gpu_render(scene, debug=True, device_resident=False, renderer=None):
    """Returns the same dict from the CUDA path (through the C ABI)."""
    from rusterizer_b200.render import Renderer
    from rusterizer_b200.scenes import render_scene

    r = renderer or Renderer(scene.width, scene.height)
    if renderer is None and scene.texture is not None:
        r.uniforms().bind_texture(0, scene.texture)
        for k, t in enumerate(getattr(scene, "extra_textures", [])):
            r.uniforms().bind_texture(k + 1, t)
    if debug:
        r.debug_capture(True)
    if getattr(scene, "scissor", None):
        r.set_scissor(*scene.scissor)
    if getattr(scene, "msaa", 4) != 4:
        r.set_msaa(scene.msaa)
    if getattr(scene, "guard_band", 1.0) != 1.0:
        r.set_guard_band(scene.guard_band)
    r.reset_counters()
    meshes = [r.upload(d.mesh) for d in scene.draws] if device_resident else None
    render_scene(r, scene, meshes)
    out = dict(fb=r.framebuffer())
    out["counters"] = r.counters()
    out["timings"] = r.timings()
    if debug:
        out["depth"], out["color"], out["owner"] = r.debug_read()
    if renderer is None:
        r.close()
    return out


def compare(o, g, check_samples=True):
    """Parity bar (BASELINE.json north_star): coverage + depth outcome bit-exact, colour within
    +-1 LSB per channel (we expect, and assert, exact equality; the tolerance is reported only)."""
    msgs = []
    if check_samples:
        if not np.array_equal(o["owner"], g["owner"]):
            bad = np.argwhere(o["owner"] != g["owner"])
            msgs.append(f"owner mismatch at {len(bad)} samples, first {bad[:5].tolist()}")
        od, gd = o["depth"].view(np.uint32), g["depth"].view(np.uint32)
        if not np.array_equal(od, gd):
            bad = np.argwhere(od != gd)
            msgs.append(f"depth bits mismatch at {len(bad)} samples, first {bad[:5].tolist()}")
        if not np.array_equal(o["color"], g["color"]):
            bad = np.argwhere(o["color"] != g["color"])
            msgs.append(f"sample colour mismatch at {len(bad)} samples, first {bad[:5].tolist()}")
    if not np.array_equal(o["fb"], g["fb"]):
        a = o["fb"].view(np.uint8).astype(np.int16)
        b = g["fb"].view(np.uint8).astype(np.int16)
        msgs.append(f"framebuffer mismatch at {int((o['fb'] != g['fb']).sum())} px, max channel diff {int(np.abs(a - b).max())}")
    for k, v in o["counters"].items():
        if g["counters"].get(k) != v:
            msgs.append(f"counter {k}: oracle {v} gpu {g['counters'].get(k)}")
    return msgs


def save_png(fb, path):
    from PIL import Image

    a = np.zeros(fb.shape + (3,), np.uint8)
    a[..., 0] = (fb >> 16) & 0xFF
    a[..., 1] = (fb >> 8) & 0xFF
    a[..., 2] = fb & 0xFF
    Image.fromarray(a).save(path)
