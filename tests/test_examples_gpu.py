"""The example drivers run end to end on the GPU (headless mirrors of the reference's tests/Raytracing.cpp)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["--forward", "--carve"]])
def test_raytracing_example(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "raytracing.py"), "--scene", "S64"] + extra,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Voxel::view set on" in r.stdout and "Resolution Single Dim: 16" in r.stdout
    if extra:
        assert "seen free" in r.stdout
    else:
        assert "found = True" in r.stdout
