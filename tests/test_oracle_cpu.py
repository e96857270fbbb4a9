"""CPU tier: the oracle against the committed golden vectors, and its self-consistency.

The reference has no golden vectors of its own.  tests/golden/golden_v1.npz was generated from this oracle
(tests/golden/make_golden.py) and guards against accidental change; the oracle itself is pinned against the
reference's own headers compiled here in tests/test_reference_build_cpu.py (PARITY UNPINNED only for the op order
inside the un-vendored Eigen).
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
CASES = ["S64", "S128-odd", "S128-clutter"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("flat", [True, False])
def test_oracle_matches_golden(oracle, dmf, golden, name, flat):
    """both storages (reference pointer grid, flat index grid) reproduce every golden vector"""
    g = golden
    K = g["K"]; H, W = (int(v) for v in g["HW"])
    sc = dmf.scenes.scene(name)
    vol = oracle.volume_from_scene(sc, flat=flat)
    assert np.array_equal(vol.occupied(), g[f"{name}/occupied"])
    poses = g[f"{name}/poses"]
    zd = int(g[f"{name}/zdelta"])
    for i, p in enumerate(poses):
        for mode, tag in ((oracle.MODE_POINTS, "points"), (oracle.MODE_GOOD_POINTS, "good")):
            for sparse in (0, 1):
                r = oracle.forward(vol, K, H, W, p, mode, zd, bool(sparse))
                assert np.array_equal(r["ids"], g[f"{name}/{i}/{tag}/s{sparse}/ids"])
                if mode == oracle.MODE_POINTS:
                    assert np.array_equal(r["depth"], g[f"{name}/{i}/depth/s{sparse}"].astype(np.int32))
                    assert [r["counters"][k] for k in ("samples", "inbounds", "hits")] == g[f"{name}/{i}/counters/s{sparse}"].tolist()
        assert oracle.forward(vol, K, H, W, p, oracle.MODE_MINIMUM, 1, True, want_pixels=False)["min_depth"] == int(g[f"{name}/{i}/min"])
        rv = oracle.reverse(vol, K, H, W, p, fast=True)
        assert np.array_equal(rv["ids"], g[f"{name}/{i}/reverse_fast/ids"]) and np.array_equal(rv["flags"], g[f"{name}/{i}/reverse_fast/flags"])
        assert np.array_equal(oracle.reverse(vol, K, H, W, p, fast=False)["ids"], g[f"{name}/{i}/reverse_slow/ids"])
        zb, n = oracle.zbuffer(vol, K, H, W, p)
        assert np.array_equal(zb, g[f"{name}/{i}/zbuffer"]) and n == int(g[f"{name}/{i}/zbuffer_n"])
    vol.clear_marks()
    for i, p in enumerate(poses):
        oracle.forward(vol, K, H, W, p, oracle.MODE_CLASSIFY, zd, False, view=1 + i, want_pixels=False)
    view, good = vol.marks()
    assert np.array_equal(view, g[f"{name}/classify/view"]) and np.array_equal(good, g[f"{name}/classify/good"])
    sets = [np.sort(g[f"{name}/{i}/reverse_fast/ids"]) for i in range(len(poses))]
    assert np.array_equal(oracle.greedy_set_cover(sets), g[f"{name}/setcover"])


def test_golden_is_not_trivial(golden):
    g = golden
    for name in CASES:
        n = len(g[f"{name}/poses"])
        assert sum(len(g[f"{name}/{i}/points/s0/ids"]) for i in range(n)) > 100
        assert sum(len(g[f"{name}/{i}/reverse_fast/ids"]) for i in range(n)) > 100
        assert g[f"{name}/classify/good"].sum() > 0


def test_eigen_order_switch_dyadic_and_ties(oracle, dmf):
    """Rule E1/E2 alternatives: identical on identity-rotation poses (provably), counted on general poses."""
    sc = dmf.scenes.scene("S64")
    vol = oracle.volume_from_scene(sc)
    K = dmf.scenes.REFERENCE_K
    p1 = dmf.scenes.pose_p1(1.024)[0]
    try:
        a = oracle.forward(vol, K, 120, 160, p1, 0, 8, False)
        oracle.set_eigen_order(1)
        b = oracle.forward(vol, K, 120, 160, p1, 0, 8, False)
        assert np.array_equal(a["depth"], b["depth"]) and np.array_equal(a["ids"], b["ids"])
        # general rotation: the alternative order may move a few samples across a voxel face ("tie cases")
        p = dmf.scenes.poses_sphere_lookat(1.024, 50)[17]
        oracle.set_eigen_order(0); a = oracle.forward(vol, K, 120, 160, p, 0, 8, False)
        oracle.set_eigen_order(1); b = oracle.forward(vol, K, 120, 160, p, 0, 8, False)
        ties = int((a["depth"] != b["depth"]).sum())
        assert ties <= 0.01 * a["depth"].size, ties   # a handful at most; reported in DESIGN.md
    finally:
        oracle.set_eigen_order(0)


def test_volume_restatement_details(oracle):
    # constructVolume recomputes dim by truncation (Volume.hpp:121-123): 93 requested -> 92 on a unit cube
    v = oracle.Volume([0, 1, 0, 1, 0, 1], [93, 93, 93])
    assert list(v.dims) == [92, 92, 92]
    v = oracle.Volume([0, 1, 0, 1, 0, 1], [128, 512, 64])
    assert list(v.dims) == [128, 512, 64]
    # strict AABB test and first-insertion order of occupied_cells_
    v = oracle.Volume([0, 1, 0, 1, 0, 1], [4, 4, 4])
    pts = np.array([[0.9, 0.1, 0.1], [0.0, 0.5, 0.5], [1.0, 0.5, 0.5], [0.1, 0.1, 0.1], [0.9, 0.1, 0.1]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (5, 1))
    assert v.integrate(pts, nrm) == 3
    ids = v.occupied()
    assert [(int(i) >> 40, (int(i) >> 20) & 0xFFFFF, int(i) & 0xFFFFF) for i in ids] == [(3, 0, 0), (0, 0, 0)]
    off, n = v.normals_csr()
    assert list(off) == [0, 2, 3]


def test_affine_inverse_restated(oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = rng.normal(size=(3, 4)).astype(np.float32)
        inv = oracle.affine_inverse(a.reshape(12)).reshape(3, 4).astype(np.float64)
        full = np.vstack([a.astype(np.float64), [0, 0, 0, 1]])
        np.testing.assert_allclose(np.vstack([inv, [0, 0, 0, 1]]) @ full, np.eye(4), atol=5e-4)


def test_degree_acos_semantics(oracle):
    assert oracle.degree_acosf(1.0) == 0
    assert oracle.degree_acosf(0.0) == 90          # acos(0)=pi/2 -> 90.00007.. -> 90
    assert oracle.degree_acosf(-0.02) == 91
    assert oracle.degree_acosf(1.0000001) == -2147483648   # NaN -> INT_MIN (x86)
    assert oracle.degree_acosf(-1.0) == 180


def test_set_cover_known_answer(oracle):
    sets = [np.arange(0, 10), np.arange(5, 30), np.arange(28, 40), np.arange(0, 3)]
    sel = oracle.greedy_set_cover([s.astype(np.uint64) for s in sets])
    assert list(sel) == [1, 2, 0]    # gains 25, 10 (30..39), 5 (0..4); set 3 adds nothing


# ---- carve mode (PixelOut::observed): an extension defined by the oracle, see oracle/dmf_oracle.hpp ------------------
GOLDEN_CARVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_carve_v1.npz")


def _bits(words):
    return np.unpackbits(words.view(np.uint8), bitorder="little")


@pytest.mark.parametrize("name", CASES)
def test_oracle_observed_matches_golden_and_is_consistent(oracle, dmf, golden, name):
    """the observed grid of carve mode: golden vectors, both storages, and what it must mean -- observed & occupied is the
    set of first-hit voxels, observed popcount <= in-bounds samples, accumulation over views is a plain OR"""
    g, gc = golden, np.load(GOLDEN_CARVE)
    K = g["K"]; H, W = (int(v) for v in g["HW"])
    sc = dmf.scenes.scene(name)
    poses = g[f"{name}/poses"]; zd = int(g[f"{name}/zdelta"])
    dims = np.asarray(sc.dims, np.int64) + 1
    for flat in (True, False):
        vol = oracle.volume_from_scene(sc, flat=flat)
        occ = vol.occupied()
        occ_idx = ((occ >> np.uint64(40)).astype(np.int64) * dims[1] + ((occ >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64)) * dims[2] + (occ & np.uint64(0xFFFFF)).astype(np.int64)
        for sparse in (0, 1):
            obs, inb, hit_ids, singles = None, 0, set(), []
            for i, p in enumerate(poses):
                obs, c = oracle.forward_observed(vol, K, H, W, p, oracle.MODE_POINTS, zd, bool(sparse), observed=obs)
                one, c1 = oracle.forward_observed(vol, K, H, W, p, oracle.MODE_POINTS, zd, bool(sparse))
                assert c1 == c and c["inbounds"] == int(g[f"{name}/{i}/counters/s{sparse}"][1])
                assert int(_bits(one).sum()) <= c["inbounds"]
                singles.append(one)
                inb += c["inbounds"]
                hit_ids.update(int(v) for v in g[f"{name}/{i}/points/s{sparse}/ids"])
                if i == 0:
                    assert np.array_equal(obs, gc[f"{name}/s{sparse}/view0"])
            assert np.array_equal(obs, gc[f"{name}/s{sparse}/all_views"]) and inb == int(gc[f"{name}/s{sparse}/inbounds"])
            assert np.array_equal(obs, np.bitwise_or.reduce(np.stack(singles), axis=0))
            b = _bits(obs)
            seen_occ = occ[b[occ_idx] == 1]
            assert set(int(v) for v in seen_occ) == hit_ids
            assert b.sum() > len(hit_ids)           # free voxels were recorded too
