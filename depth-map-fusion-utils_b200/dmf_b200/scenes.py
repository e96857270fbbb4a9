"""Deterministic synthetic scenes and pose sets.

No files, no RNG state: everything is a pure function of its arguments, so the CPU oracle, the CUDA
path and the committed golden fixtures all see the same inputs.  Point clouds are (n,3) float32 with
(n,3) float32 normals, in the order they would be fed to VoxelVolume::integratePointCloud
(reference include/Volume.hpp:199-228); poses are (n,12) float32 row-major 3x4 camera->world affines
(the Eigen::Affine3f the reference passes to RayTracingEngine, include/RayTracingEngine.hpp:447).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# Camera intrinsics repeated verbatim in six reference drivers (tests/Raytracing.cpp:61).
REFERENCE_K = np.array(
    [602.39306640625, 0.0, 314.6370849609375, 0.0, 602.39306640625, 245.04962158203125, 0.0, 0.0, 1.0],
    dtype=np.float32,
)


def scaled_K(factor: float) -> np.ndarray:
    """Reference K scaled for a factor-times larger image (1080x1920 = 3x would be 1440x1920; we follow
    SURVEY 8d: x3 K with a 1080x1920 sensor)."""
    K = REFERENCE_K.astype(np.float64).copy()
    K[[0, 2, 4, 5]] *= factor
    return K.astype(np.float32)


@dataclass
class Scene:
    name: str
    bounds: np.ndarray  # (6,) float64: xmin,xmax,ymin,ymax,zmin,zmax  (setDimensions)
    dims: np.ndarray  # (3,) int32 requested by setVolumeSize
    points: np.ndarray  # (n,3) float32
    normals: np.ndarray  # (n,3) float32

    @property
    def voxel_size(self) -> float:
        d = (self.bounds[1::2] - self.bounds[0::2]) / self.dims
        return float(d[0] * d[1] * d[2])

    @property
    def zdelta(self) -> int:
        """resolution_single_dimension of tests/Raytracing.cpp:84-85."""
        return int(round(float(np.cbrt(self.voxel_size * 1e9))))


def _shell_voxels(lo: int, hi: int):
    """Indices (x-major, then y, then z) of the one-voxel-thick shell of the cube [lo,hi]^3 and, per voxel,
    the list of outward face normals it carries (order -x,+x,-y,+y,-z,+z)."""
    n = hi - lo + 1
    r = np.arange(lo, hi + 1, dtype=np.int32)
    chunks = []
    for x in r:
        if x == lo or x == hi:
            yy, zz = np.meshgrid(r, r, indexing="ij")
            yz = np.stack([yy.ravel(), zz.ravel()], 1)
        else:
            yy, zz = np.meshgrid(r, r, indexing="ij")
            m = (yy == lo) | (yy == hi) | (zz == lo) | (zz == hi)
            yz = np.stack([yy[m], zz[m]], 1)  # boolean mask keeps row-major (y, z) order
        chunks.append(np.concatenate([np.full((len(yz), 1), x, np.int32), yz], 1))
    vox = np.concatenate(chunks, 0)
    assert len(vox) == n**3 - max(n - 2, 0) ** 3
    return vox


def box_shell(L: float, N: int, name: str | None = None, frac=(0.35, 0.65)) -> Scene:
    """'Synthetic box': hollow shell spanning [0.35L,0.65L]^3, one voxel thick; one point per (voxel, incident
    face) at the voxel centre with that face's outward normal."""
    lo, hi = int(frac[0] * N), int(frac[1] * N)
    vox = _shell_voxels(lo, hi)
    delta = L / N
    faces = [(0, lo, -1.0), (0, hi, 1.0), (1, lo, -1.0), (1, hi, 1.0), (2, lo, -1.0), (2, hi, 1.0)]
    # per voxel, per face flag -> repeat voxel once per incident face, keeping voxel order
    flags = np.stack([vox[:, ax] == v for ax, v, _ in faces], 1)  # (n,6)
    rep = flags.sum(1)
    centres = ((vox.astype(np.float64) + 0.5) * delta).astype(np.float32)
    pts = np.repeat(centres, rep, axis=0)
    vi, fi = np.nonzero(flags)  # row-major => voxel order, then face order
    nrm = np.zeros((len(vi), 3), np.float32)
    for k, (ax, _, s) in enumerate(faces):
        nrm[fi == k, ax] = s
    b = np.array([0, L, 0, L, 0, L], np.float64)
    return Scene(name or f"box{N}", b, np.array([N, N, N], np.int32), pts, nrm)


_MASK64 = (1 << 64) - 1


def splitmix64(seed: int):
    """Generator of uniform doubles in [0,1) from splitmix64 (public-domain constants)."""
    state = seed & _MASK64
    while True:
        state = (state + 0x9E3779B97F4A7C15) & _MASK64
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
        z ^= z >> 31
        yield (z >> 11) / float(1 << 53)


def clutter(L: float, N: int, n_spheres: int = 64, seed: int = 0xD3F7, name: str | None = None) -> Scene:
    """Box shell + solid spheres with radial normals (S512-clutter)."""
    base = box_shell(L, N)
    g = splitmix64(seed)
    delta = L / N
    pts, nrm = [base.points], [base.normals]
    for _ in range(n_spheres):
        c = np.array([0.2 * L + 0.6 * L * next(g) for _ in range(3)])
        rad = 0.01 + 0.03 * next(g)
        lo = np.maximum(np.floor((c - rad) / delta).astype(int), 0)
        hi = np.minimum(np.floor((c + rad) / delta).astype(int), N - 1)
        ax = [np.arange(lo[i], hi[i] + 1) for i in range(3)]
        gx, gy, gz = np.meshgrid(*ax, indexing="ij")
        cen = (np.stack([gx, gy, gz], -1).reshape(-1, 3) + 0.5) * delta
        d = cen - c
        dist = np.linalg.norm(d, axis=1)
        m = (dist <= rad) & (dist > 0)
        pts.append(cen[m].astype(np.float32))
        nrm.append((d[m] / dist[m, None]).astype(np.float32))
    b = np.array([0, L, 0, L, 0, L], np.float64)
    return Scene(name or f"clutter{N}", b, np.array([N, N, N], np.int32), np.concatenate(pts), np.concatenate(nrm))


def scene(name: str) -> Scene:
    """Named scenes of SURVEY 8d."""
    if name == "S32":  # tiny, for pure-Python cross checks
        return box_shell(1.024, 32, "S32")
    if name == "S64":
        return box_shell(1.024, 64, "S64")
    if name == "S128":
        return box_shell(1.024, 128, "S128")
    if name == "S128d":  # dyadic 128^3 (voxel = 2^-7 m): the float-accumulating whole-grid loops are exact here
        return box_shell(1.0, 128, "S128d")
    if name == "S256":
        return box_shell(1.0, 256, "S256")
    if name == "S512":
        return box_shell(1.0, 512, "S512")
    if name == "S1024":
        return box_shell(1.0, 1024, "S1024")
    if name == "S128-odd":  # non-dyadic bounds, for tie statistics
        return box_shell(0.937, 128, "S128-odd")
    if name == "S512-odd":
        return box_shell(0.937, 512, "S512-odd")
    if name == "S128-clutter":
        return clutter(1.024, 128, 24, name="S128-clutter")
    if name == "S512-clutter":
        return clutter(1.0, 512, 64, name="S512-clutter")
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------ poses


def _pose(x, y, z, t) -> np.ndarray:
    m = np.zeros((3, 4), np.float64)
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = x, y, z, t
    return m.astype(np.float32).reshape(12)


def pose_p1(L: float) -> np.ndarray:
    """P1: identity linear part, camera at (0.5L, 0.5L, 0.02) looking along +z at the box."""
    return _pose([1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5 * L, 0.5 * L, 0.02])[None, :]


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    eye, target, up = (np.asarray(a, np.float64) for a in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    if np.linalg.norm(x) < 1e-9:
        x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return _pose(x, y, z, eye)


def sphere_directions(factor: float = 24.0) -> np.ndarray:
    """Unit directions in the (phi, theta) order of Algorithms::generateSphere (Algorithms.hpp:88-112)."""
    out = []
    phi = 0.0
    while phi <= 2 * math.pi:
        theta = 0.0
        while theta <= math.pi:
            out.append((math.cos(phi) * math.sin(theta), math.sin(phi) * math.sin(theta), math.cos(theta)))
            theta += math.pi / factor
        phi += math.pi / factor
    return np.array(out, np.float64)


def poses_sphere_lookat(L: float, n: int, radius: float = 0.45, factor: float = 24.0) -> np.ndarray:
    """P1024 (look-at variant): cameras on a sphere around the cube centre looking at it."""
    c = np.array([0.5 * L] * 3)
    d = sphere_directions(factor)
    d = d[np.linalg.norm(d[:, :2], axis=1) > 1e-6]  # drop the poles (repeated 2*factor+1 times each)
    assert len(d) >= n, (len(d), n)
    return np.stack([look_at(c + radius * v, c) for v in d[:n]])


def poses_position_camera(L: float, n: int, standoff: float = 0.5, factor: float = 24.0, box_half: float = 0.15) -> np.ndarray:
    """P1024 (reference variant): Algorithms::positionCamera poses (Algorithms.hpp:190-236) for surface samples
    of the box along sphere directions: fixed x=(0,-1,0), y=(1,0,0) columns, z = -normal (after the flip of
    positionCameras :286-292), translation = point + normal*standoff.  The linear part is not a rotation."""
    c = np.array([0.5 * L] * 3)
    d = sphere_directions(factor)
    d = d[np.linalg.norm(d[:, :2], axis=1) > 1e-6]
    out = []
    for v in d[:n]:
        nrm = v.astype(np.float32).astype(np.float64)
        p = c + v * (box_half * L / np.max(np.abs(v)))  # point on the box surface along v
        if nrm[2] <= 0:
            nrm = -nrm
        t = p + nrm * standoff
        out.append(_pose([0, -1, 0], [1, 0, 0], -nrm, t))
    return np.stack(out)


def position_cameras(points, normals, distance: int = 300) -> np.ndarray:
    """Algorithms::positionCameras(locations, distance) (Algorithms.hpp:282-298 -> positionCamera :190-236), float for float:
    normals with z <= 0 are flipped (:286-292); x = (0,-1,0), y = (1,0,0) (the fixed columns of :223-225), z = -normal;
    translation = point + Vector3f(normal) * float(distance / 1000.0) evaluated like movePointAway (:114-122): the product in
    float, the sum in double, stored to float.  points, normals: (n, 3) -> (n, 12) float32 poses."""
    f32 = np.float32
    pts = np.ascontiguousarray(points, f32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, f32).reshape(-1, 3).copy()
    nrm[nrm[:, 2] <= 0] *= f32(-1)
    d = f32(float(distance) / 1000.0)                                  # double(distance)/1000.0 narrowed to float by Vector3f * double
    moved = (nrm * d).astype(np.float64) + pts.astype(np.float64)      # n(i) + pi[i]: float product, double sum
    out = np.zeros((len(pts), 3, 4), f32)
    out[:, :, 0] = (0, -1, 0)
    out[:, :, 1] = (1, 0, 0)
    out[:, :, 2] = nrm * f32(-1)
    out[:, :, 3] = moved.astype(f32)
    return out.reshape(-1, 12)


def poses_helix(L: float, n: int, radius: float = 0.48, seed: int = 0xC0FFEE) -> np.ndarray:
    """P10k: helix around the cube centre with +-1 mm / +-0.2 deg jitter."""
    g = splitmix64(seed)
    c = np.array([0.5 * L] * 3)
    out = []
    for i in range(n):
        th = 2 * math.pi * i / 500.0
        eye = np.array([c[0] + radius * math.cos(th), c[1] + radius * math.sin(th), 0.1 * L + 0.8 * L * i / max(n, 1)])
        eye += (np.array([next(g), next(g), next(g)]) - 0.5) * 0.002
        tgt = c + (np.array([next(g), next(g), next(g)]) - 0.5) * 2 * radius * math.tan(math.radians(0.2))
        out.append(look_at(eye, tgt))
    return np.stack(out)


def poses_fibonacci(L: float, n: int, radius: float = 0.48) -> np.ndarray:
    """P4096: Fibonacci sphere, look-at centre."""
    c = np.array([0.5 * L] * 3)
    ga = math.pi * (3.0 - math.sqrt(5.0))
    out = []
    for i in range(n):
        zc = 1 - 2 * (i + 0.5) / n
        rr = math.sqrt(max(0.0, 1 - zc * zc))
        v = np.array([rr * math.cos(ga * i), rr * math.sin(ga * i), zc])
        out.append(look_at(c + radius * v, c))
    return np.stack(out)


def bench_poses(L: float, n: int) -> np.ndarray:
    """Pose batch used by bench.py: the sphere look-at sweep, repeated cyclically to n views."""
    base = poses_sphere_lookat(L, min(n, 1024))
    reps = -(-n // len(base))
    return np.tile(base, (reps, 1))[:n].copy()
