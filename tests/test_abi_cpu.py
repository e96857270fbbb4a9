"""CPU tier: the C-ABI library loads, exports every symbol include/dmf_b200.h declares, refuses to compute without
a GPU (no fallback), and the product never links or imports the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmf_b200.h")
PKG = os.path.join(ROOT, "depth-map-fusion-utils_b200")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(dmf):
    lib = C.CDLL(dmf.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmf_b200.h but not exported"
    assert set(names) == set(dmf.SYMBOLS), set(names) ^ set(dmf.SYMBOLS)


def test_library_contains_sm100a_code(dmf):
    out = subprocess.run(["cuobjdump", "-lelf", dmf.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout, out.stdout


def test_no_cpu_fallback_without_gpu(dmf):
    lib = dmf.load()
    if lib.dmf_device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU tier only")
    with pytest.raises(dmf.DmfError, match="no CPU fallback"):
        dmf.Context(0)


def test_product_does_not_reference_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "dmf_oracle.hpp\"" not in txt.replace("oracle/dmf_oracle.hpp", ""), f
    ldd = subprocess.run(["ldd", os.path.join(PKG, "libdmf_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_host_angle_threshold_matches_host_libm(dmf, oracle):
    """the kernels' `dot_min <= d <= 1` is the reference's degree(acos(d)) in [0,90] on this host's libm"""
    out = np.zeros(3, np.float32)
    assert dmf.load().dmf_host_angle_test(out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    dot_min, lo, hi = (float(v) for v in out)
    assert -0.018 < dot_min < -0.017
    assert not (lo < hi), "host acosf non-monotonic around the 90-degree threshold; ties would be counted"
    d = np.float32(dot_min)
    for _ in range(2000):
        d = np.nextafter(d, np.float32(-1))
        assert not (0 <= oracle.degree_acosf(float(d)) <= 90)
    d = np.float32(dot_min)
    for _ in range(2000):
        assert 0 <= oracle.degree_acosf(float(d)) <= 90
        d = np.nextafter(d, np.float32(1))
    rng = np.random.default_rng(0)
    for d in rng.uniform(-1.2, 1.2, 5000).astype(np.float32):
        assert (0 <= oracle.degree_acosf(float(d)) <= 90) == (dot_min <= float(d) <= 1.0)


def test_scenes_are_deterministic(dmf):
    a, b = dmf.scenes.scene("S64"), dmf.scenes.scene("S64")
    assert np.array_equal(a.points, b.points) and np.array_equal(a.normals, b.normals)
    assert dmf.scenes.scene("S128").zdelta == 8 and dmf.scenes.scene("S512").zdelta == 2
    assert len(dmf.scenes.scene("S512").points) == 142296
    p = dmf.scenes.bench_poses(1.0, 1024)
    assert p.shape == (1024, 12) and p.dtype == np.float32 and len(np.unique(p, axis=0)) == 1024
    bits = np.array([0b101, 1 << 63], np.uint64)
    assert list(dmf.bits_to_indices(bits)) == [0, 2, 127]


def test_pose_file_wire_format(dmf, tmp_path):
    """FileRoutines.hpp:69-112: count line, then 3 comma-separated rows per pose, %g formatting"""
    from dmf_b200.posefile import read_camera_locations, write_camera_locations
    poses = dmf.scenes.poses_sphere_lookat(1.0, 5)
    path = str(tmp_path / "cams.txt")
    write_camera_locations(path, poses)
    lines = open(path).read().splitlines()
    assert lines[0] == "5" and len(lines) == 16 and all(len(l.split(",")) == 4 for l in lines[1:])
    assert lines[2].split(",")[0] in ("-1", "1") or "." in lines[2] or "e" in lines[2]
    back = read_camera_locations(path)
    assert back.shape == (5, 12)
    np.testing.assert_allclose(back, poses, rtol=1e-5, atol=1e-6)       # 6 significant digits survive
    write_camera_locations(path, back)                                   # idempotent from the second write on
    assert np.array_equal(read_camera_locations(path), back)
    open(path, "w").write("1\n1,0,0\n0,1,0,0\n0,0,1,0\n")
    with pytest.raises(ValueError):
        read_camera_locations(path)


def test_multi_gpu_entry_points_refuse_without_gpu(dmf):
    """CPU tier: the group entry points exist, and without a GPU they fail with the library's message (no fallback, no crash)"""
    lib = dmf.load()
    if lib.dmf_device_count() > 0:
        pytest.skip("a GPU is visible; the group is exercised by tests/test_multi_gpu.py and bench.py")
    h = C.c_void_p()
    assert lib.dmf_comm_init_all(C.byref(h), 0) != 0 and not h.value
    assert b"no CPU fallback" in lib.dmf_last_error()
    assert lib.dmf_comm_init_rank(C.byref(h), None, None, 0, 1) != 0
    for bad in (lib.dmf_sweep_forward(None, None, None, 0, None), lib.dmf_sweep_reverse(None, 1, None, 0, None), lib.dmf_comm_fuse_observed(None), lib.dmf_comm_fuse_marks(None, 1)):
        assert bad != 0


def test_view_dealing_of_the_group_matches_the_python_sharding():
    """view g on GPU g mod N (dmf_comm.cuh my_view_count / row0 + j * row_step) is sweep.shard_indices(..., "strided")"""
    from dmf_b200.sweep import shard_indices
    for n in (0, 1, 7, 37, 128, 1024):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                idx = shard_indices(n, rank, world, "strided")
                assert len(idx) == ((n - rank + world - 1) // world if rank < n else 0)      # my_view_count
                assert all(g % world == rank for g in idx)
                seen.extend(idx.tolist())
            assert sorted(seen) == list(range(n))
