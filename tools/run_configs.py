#!/usr/bin/env python
"""Measure every BASELINE.json config on one GPU (device-resident meshes, CUDA-event stage times):
C1 default scene, C2 1M-triangle sphere, C2' overdraw, C3 near-clip field at 4K, C4 8192^2 (sphere and
full-screen quad).  Prints one JSON object per config; parity of each scene vs the oracle is checked
at reduced size by tests/test_gpu_parity.py."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from rusterizer_b200 import scenes
from rusterizer_b200.render import Renderer

CONFIGS = [
    ("C1 default scene 1280x720 (268 tris, 2 draws)", lambda: scenes.default_scene(1.0)),
    ("C2 1M-tri sphere 1920x1080", lambda: scenes.sphere_scene()),
    ("C2' overdraw 4x250K grids back-to-front 1920x1080", lambda: scenes.overdraw_scene(back_to_front=True)),
    ("C2' overdraw front-to-back", lambda: scenes.overdraw_scene(back_to_front=False)),
    ("C3 near-clip field 250K tris 3840x2160", lambda: scenes.near_clip_scene()),
    ("C4(i) 1M-tri sphere 8192x8192", lambda: scenes.sphere_scene(width=8192, height=8192)),
    ("C4(ii) full-screen quad 8192x8192", lambda: scenes.fullscreen_quad_scene(8192, 8192)),
]
for name, mk in CONFIGS:
    sc = mk()
    r = Renderer(sc.width, sc.height)
    r.uniforms().bind_texture(0, sc.texture)
    dms = [r.upload(d.mesh) for d in sc.draws]
    ts = []
    for i in range(25):  # 10 warm-up frames (the first config also ramps the clocks), median of the next 15
        if i == 10: r.reset_counters()
        scenes.render_scene(r, sc, dms); r.framebuffer_device()
        ts.append(r.timings())
    t = {k: float(np.median([x[k] for x in ts[10:]])) for k in ts[0]}
    c = {k: v / 15 for k, v in r.counters().items()}
    print(json.dumps({"config": name, "tris": sc.n_triangles, "ms": t, "Mtris_per_s": sc.n_triangles / t["total_ms"] / 1e3,
                      "Gsamples_per_s": c["n_samples_written"] / t["total_ms"] / 1e6,
                      "clipped_in": c["n_clipped_in"], "samples_written": c["n_samples_written"], "bbox_px": c["n_bbox_px"],
                      "hbm_floor_frac": sc.algorithmic_bytes() / (t["total_ms"] / 1e3) / 6539.2e9}), flush=True)
    r.close()
