#!/usr/bin/env python
"""tests/fuzz_parity.py [--seconds S] [--seed N]: a timed fuzz campaign of the CUDA path against the CPU oracle
(test infrastructure, run on a GPU box: `python tests/fuzz_parity.py --seconds 240 > gpurun_out/fuzz.txt`).

Every case is a seeded random frame: random resolution (odd sizes included), one to four draws of random triangle
soups or indexed grids, triangle sizes from sub-pixel slivers to several screens, depths that cross the near and far
planes, both windings, duplicated (coincident) triangles, any of the three reference shaders, a random sample count
(1/2/4/8), guard band and scissor rect.  The comparison is tests/helpers.compare: per-sample owner, depth
bits and colour, the resolved image and every work counter must be equal.  Prints one line per failing case with the
seed that reproduces it, and a summary line; exit status 1 if any case failed.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from helpers import compare, gpu_render, oracle_render  # noqa: E402
from rusterizer_b200 import mathx, scenes  # noqa: E402
from rusterizer_b200.mesh import Mesh  # noqa: E402


def soup(rng, nt, lo, hi, zlo, zhi):
    ctr = rng.uniform(-3, 3, (nt, 1, 3)).astype(np.float32)
    ctr[..., 2] = rng.uniform(zlo, zhi, (nt, 1)).astype(np.float32)
    size = (10 ** rng.uniform(lo, hi, (nt, 1, 1))).astype(np.float32)
    shape = rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32)
    sliver = rng.rand(nt) < 0.3  # a third are slivers: the third vertex nearly on the segment of the other two
    t = rng.uniform(0, 1, (nt, 1)).astype(np.float32)
    shape[sliver, 2] = (shape[sliver, 0] * (1 - t[sliver]) + shape[sliver, 1] * t[sliver]
                        + rng.normal(0, 1, (int(sliver.sum()), 3)).astype(np.float32) * np.float32(10.0 ** rng.uniform(-4, -1)))
    verts = (ctr + shape * size).reshape(-1, 3)
    idx = np.arange(nt * 3, dtype=np.uint32)
    if rng.rand() < 0.3:  # coincident duplicates, later in submission order
        dup = rng.randint(0, nt, max(1, nt // 10))
        idx = np.concatenate([idx, (dup[:, None] * 3 + np.arange(3)[None]).reshape(-1).astype(np.uint32)])
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    return Mesh(verts, idx, attrs)


def grid(rng, nx, ny, z, tilt):
    xs = np.linspace(-2.5, 2.5, nx + 1, dtype=np.float32)
    ys = np.linspace(-1.6, 1.6, ny + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    Z = (z + tilt * Y + rng.normal(0, 0.02, X.shape)).astype(np.float32)
    verts = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    a = (jj * (nx + 1) + ii).reshape(-1)
    quads = np.stack([a, a + 1, a + nx + 2, a, a + nx + 2, a + nx + 1], -1)
    if rng.rand() < 0.5:
        quads = quads[:, ::-1]
    attrs = rng.uniform(0, 1, (len(verts), 6)).astype(np.float32)
    return Mesh(verts, quads.reshape(-1).astype(np.uint32), attrs)


def make_case(seed):
    rng = np.random.RandomState(seed)
    W = int(rng.choice([64, 97, 160, 255, 320, 333, 512, 640, 1000]))
    H = int(rng.choice([48, 61, 120, 144, 180, 250, 288, 400]))
    draws = []
    for _ in range(rng.randint(1, 5)):
        fs = int(rng.choice([scenes.FS_COLOR, scenes.FS_TEXTURE, scenes.FS_DEBUG]))
        kind = rng.rand()
        if kind < 0.55:
            lo = rng.uniform(-3.0, -1.0)
            mesh = soup(rng, int(rng.randint(20, 2500)), lo, lo + rng.uniform(0.5, 3.0), rng.uniform(-4.6, -2), rng.uniform(-1, 60))
        elif kind < 0.85:
            mesh = grid(rng, int(rng.randint(2, 120)), int(rng.randint(2, 80)), rng.uniform(-4.5, 10), rng.uniform(-2, 2))
        else:  # a handful of huge triangles (several screens large)
            mesh = soup(rng, int(rng.randint(1, 12)), 0.5, 1.8, -4.9, 5)
        world = mathx.translate(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(-1, 2)))
        draws.append(scenes.Draw(mesh, world, fs))
    base = scenes.default_scene(0.0, width=W, height=H)
    s = scenes.Scene(f"fuzz{seed}", W, H, base.view, base.projection, draws, base.texture)
    s.msaa = int(rng.choice([4, 4, 4, 1, 2, 8]))
    s.guard_band = float(rng.choice([1.0, 1.0, 1.5, 4.0]))
    if rng.rand() < 0.25:
        x0, y0 = int(rng.randint(0, W)), int(rng.randint(0, H))
        s.scissor = (x0, y0, int(rng.randint(x0, W + 20)), int(rng.randint(y0, H + 20)))
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--only", type=int, default=None, help="run this one seed and print the mismatches")
    a = ap.parse_args()
    if a.only is not None:
        s = make_case(a.only)
        msgs = compare(oracle_render(s), gpu_render(s, debug=True))
        print(s.name, s.width, s.height, "msaa", s.msaa, "guard", s.guard_band, "scissor", getattr(s, "scissor", None), msgs or "ok")
        return 1 if msgs else 0
    t0, n, bad, tris, samples = time.time(), 0, 0, 0, 0
    seed = a.seed
    while time.time() - t0 < a.seconds:
        s = make_case(seed)
        o = oracle_render(s)
        g = gpu_render(s, debug=True, device_resident=bool(seed & 1))
        msgs = compare(o, g)
        n += 1
        tris += int(o["counters"]["n_tris_in"])
        samples += int(o["counters"]["n_samples_written"])
        if msgs:
            bad += 1
            print(f"FAIL seed {seed}: {s.width}x{s.height} msaa {s.msaa} guard {s.guard_band} scissor {getattr(s, 'scissor', None)}: "
                  + "; ".join(msgs), flush=True)
        seed += 1
    print(f"fuzz: {n} frames (seeds {a.seed}..{seed - 1}), {tris} triangles, {samples} samples written, {bad} mismatching frames, "
          f"{time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
