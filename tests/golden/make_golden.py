#!/usr/bin/env python
"""Regenerates tests/golden/frames.json: sha256 of the resolved u32 framebuffer, of the per-sample
depth bits and of the per-sample owner keys, plus the work counters, of the CPU oracle
(oracle/rz_oracle.c, pinned to the reference by tests/test_oracle_kats.py) on a set of small scenes.

The reference crate is Rust and cannot run here, so these vectors come from the oracle; they guard
against drift of the oracle itself and let the GPU tests check against committed data.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def golden_scenes():
    from rusterizer_b200 import scenes

    return {
        "default_t1.0_texture": scenes.default_scene(1.0, fs=0),
        "default_t2.5_color": scenes.default_scene(2.5, fs=1),
        "default_t0.0_debug": scenes.default_scene(0.0, fs=2),
        "clip_test_t0.0": scenes.clip_test_scene(0.0),
        "clip_test_t3.0": scenes.clip_test_scene(3.0),
        "sphere_20k_640x360": scenes.sphere_scene(101, 101, width=640, height=360),
        "near_clip_6400_640x360": scenes.near_clip_scene(80, 40, width=640, height=360),
        "overdraw_b2f_480x270": scenes.overdraw_scene(60, 30, width=480, height=270, back_to_front=True),
        "fullscreen_quad_512": scenes.fullscreen_quad_scene(512, 512),
    }


def digest(result):
    h = lambda a: hashlib.sha256(a.tobytes()).hexdigest()
    return {"fb_sha256": h(result["fb"]), "depth_sha256": h(result["depth"]), "owner_sha256": h(result["owner"]),
            "color_sha256": h(result["color"]), "counters": result["counters"]}


if __name__ == "__main__":
    from helpers import oracle_render

    out = {name: digest(oracle_render(sc)) for name, sc in golden_scenes().items()}
    path = Path(__file__).with_name("frames.json")
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path, len(out), "scenes")
