#!/usr/bin/env python
"""tools/stage_times.py [LIB ...]: per-stage CUDA-event times of the C2 frame for the default library and
each variant (RZ_B200_LIB); no parity checks -- experimental variants may render wrong images."""
import os, subprocess, sys, json
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from rusterizer_b200 import scenes
    from rusterizer_b200.render import Renderer
    which = os.environ.get("RZ_SCENE", "c2")  # c2 | overdraw | c4i | c4ii | c3 | c1
    sc = {"c2": lambda: scenes.sphere_scene(1001, 501), "overdraw": lambda: scenes.overdraw_scene(),
          "c4i": lambda: scenes.sphere_scene(1001, 501, width=8192, height=8192), "c4ii": lambda: scenes.fullscreen_quad_scene(),
          "c3": lambda: scenes.near_clip_scene(), "c1": lambda: scenes.default_scene(1.0)}[which]()
    r = Renderer(sc.width, sc.height)
    r.uniforms().bind_texture(0, sc.texture)
    ms = [r.upload(d.mesh) for d in sc.draws]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    acc = {}
    for i in range(25):
        flush.zero_(); torch.cuda.synchronize()
        scenes.render_scene(r, sc, ms); r.framebuffer_device()
        if i >= 5:
            for k, v in r.timings().items(): acc.setdefault(k, []).append(v)
    print(json.dumps({k: round(1e3 * sorted(v)[len(v) // 2], 1) for k, v in acc.items()}))
else:
    for lib in ["default"] + sys.argv[1:]:
        env = dict(os.environ)
        lib, *kv = lib.split("@")  # LIB@ENV=VALUE@ENV=VALUE: tuning knobs for this run
        for e in kv:
            env[e.split("=")[0]] = e.split("=")[1]
        if lib != "default": env["RZ_B200_LIB"] = os.path.abspath(lib)
        lib = "@".join([lib] + kv)
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        print(lib, os.environ.get("RZ_SCENE", "c2"), "median us:", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
