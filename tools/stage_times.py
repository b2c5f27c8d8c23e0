#!/usr/bin/env python
"""tools/stage_times.py [LIB ...]: per-stage CUDA-event times of the C2 frame for the default library and
each variant (RZ_B200_LIB); no parity checks -- experimental variants may render wrong images."""
import os, subprocess, sys, json
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from rusterizer_b200 import scenes
    from rusterizer_b200.render import Renderer
    sc = scenes.sphere_scene(1001, 501)
    r = Renderer(sc.width, sc.height)
    r.uniforms().bind_texture(0, sc.texture)
    m = r.upload(sc.draws[0].mesh)
    b = r.uniforms().write_block(); b.projection = sc.projection; b.view = sc.view; b.world = sc.draws[0].world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    acc = {}
    for i in range(25):
        flush.zero_(); torch.cuda.synchronize()
        r.render(m, 0, 0); r.framebuffer_device()
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
        print(lib, "median us:", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
