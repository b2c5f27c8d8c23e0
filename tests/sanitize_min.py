"""tests/sanitize_min.py: two small frames through the release tile kernel (multi-chunk tiles with sort, register
replay, dense chunks; small-triangle sphere) for a quick `compute-sanitizer --tool racecheck python tests/sanitize_min.py`
(5 s on a B200); tests/sanitize_workload.py is the full workload."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import gpu_render
from rusterizer_b200 import scenes
sc = scenes.overdraw_scene(nx=24, ny=12, width=96, height=64)   # multi-chunk tiles: sort + register replay + dense chunks
gpu_render(sc, debug=False); print("ok overdraw", flush=True)
sc = scenes.sphere_scene(33, 17, width=96, height=64)
gpu_render(sc, debug=False); print("ok sphere", flush=True)
