"""CPU tier: the relink proof (BASELINE north_star "the tests/ drivers relink unchanged").

Where the reference tree exists (/root/reference: the authoring container, not the GPU box) its UNMODIFIED drivers and
include/Algorithms.hpp are compiled with the drop-in Camera.hpp / Volume.hpp / RayTracingEngine.hpp FIRST on the include path,
so every call the reference makes into the three hot-path classes type-checks against the drop-in signatures and defaults.
Eigen, PCL and the PCL/VTK viewer resolve to the stand-ins in oracle/ref_shim (test scaffolding; the real libraries are not in
this image).  `-include numeric ...`: the reference relies on <numeric>/<algorithm>/<iterator>/<climits> arriving through the real
Eigen/PCL headers.  tests/CameraPathGen.cpp is additionally compiled to an object and LINKED against libdmf_b200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "depth-map-fusion-utils_b200")
REF = os.environ.get("DMF_REFERENCE", "/root/reference")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-std=c++17", "-w", "-include", "numeric", "-include", "algorithm", "-include", "iterator", "-include", "climits",
         "-I", os.path.join(PKG, "dropin"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-I", os.path.join(REF, "include")]
DRIVERS = ["CameraPathGen", "SetCover", "Raytracing", "CameraMotionPlanner", "CameraMotionTSP", "CameraPlacement"]

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "include", "RayTracingEngine.hpp")), reason="no reference tree here (GPU box)")


@needs_ref
@pytest.mark.parametrize("driver", DRIVERS)
def test_reference_driver_compiles_against_dropin_headers(driver):
    r = subprocess.run([GXX, "-fsyntax-only"] + FLAGS + [os.path.join(REF, "tests", driver + ".cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@needs_ref
def test_dropin_headers_are_the_ones_found_first():
    """the include path really resolves the three hot-path headers to the drop-ins, not to the reference's"""
    r = subprocess.run([GXX, "-fsyntax-only", "-H"] + FLAGS + [os.path.join(REF, "tests", "SetCover.cpp")], capture_output=True, text=True)
    assert r.returncode == 0
    for h in ("Camera.hpp", "Volume.hpp", "RayTracingEngine.hpp"):
        hits = [ln.strip(". ") for ln in r.stderr.splitlines() if ln.rstrip().endswith("/" + h)]
        assert hits and all(os.path.join(PKG, "dropin") in x for x in hits), (h, hits)
    assert any(ln.rstrip().endswith(os.path.join(REF, "include", "Algorithms.hpp")) for ln in r.stderr.splitlines()), "the reference's own Algorithms.hpp should be the one compiled"


@needs_ref
def test_reference_driver_links_against_the_library(tmp_path):
    """tests/CameraPathGen.cpp, unchanged: compile to an executable against the drop-in headers and link libdmf_b200.so"""
    exe = str(tmp_path / "CameraPathGen")
    r = subprocess.run([GXX, "-O1"] + FLAGS + [os.path.join(REF, "tests", "CameraPathGen.cpp"), "-L", PKG, "-ldmf_b200", f"-Wl,-rpath,{PKG}", "-pthread", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    assert "dmf_forward" in syms and "dmf_reverse" in syms, "the driver's RayTracingEngine calls should resolve to the C ABI"
