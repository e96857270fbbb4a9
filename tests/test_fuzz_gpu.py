"""GPU tier: randomised differential fuzz.  2000 random poses x 4 volumes (incl. the anisotropic off-origin
one): k_forward_line, k_forward_skip and carve-on-line against the brute-force march, the line-first reverse march against the
bit-grid one -- depth, voxel ids, points, visibility, counters, observed grids.  Run twice: against the production library and
against libdmf_b200_checked.so (-DDMF_CHECKED), whose own bounds checks on every computed grid index must count 0 violations
(the substitute for compute-sanitizer, which is closed on this GPU pool)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "depth-map-fusion-utils_b200")


def _run(lib, poses):
    env = dict(os.environ)
    env["DMF_FUZZ_POSES"] = str(poses)
    if lib:
        env["DMF_B200_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_driver.py")], capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_fuzz_production_build():
    rep = _run(None, 2000)
    assert rep["lib_version"] % 2 == 0
    assert rep["mismatches"] == 0, rep
    assert all(v["hit_pixels_seen"] > 10000 for v in rep["volumes"].values()), rep


def test_fuzz_checked_build_counts_no_bounds_violation():
    lib = os.path.join(PKG, "libdmf_b200_checked.so")
    if not os.path.exists(lib):
        pytest.fail("libdmf_b200_checked.so is missing: __graft_entry__.build() builds it (build.py build_checked)")
    rep = _run(lib, 1000)
    assert rep["lib_version"] % 2 == 1, "the checked library was not the one loaded"
    assert rep["bounds_violations"] == 0 and rep["mismatches"] == 0, rep
