"""Committed golden vectors (tests/golden/frames.json, made by tests/golden/make_golden.py)."""
import importlib.util
import json
from pathlib import Path

import pytest

from helpers import gpu_render, oracle_render

GOLD = Path(__file__).parent / "golden"
spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
FRAMES = json.loads((GOLD / "frames.json").read_text())
SCENES = mg.golden_scenes()


@pytest.mark.parametrize("name", sorted(FRAMES))
def test_oracle_reproduces_golden(name):
    assert mg.digest(oracle_render(SCENES[name])) == FRAMES[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FRAMES))
def test_gpu_matches_golden(name):
    """The CUDA path against committed data only (no oracle in the loop)."""
    assert mg.digest(gpu_render(SCENES[name], debug=True)) == FRAMES[name]
