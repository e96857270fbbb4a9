"""Camera-pose text files in the reference's wire format (include/FileRoutines.hpp:69-112): first line = number of poses,
then three lines per pose, each "a,b,c,d" = one row of the 3x4 camera->world affine, written with the default C++
ostream float formatting (6 significant digits, "%g").  This is how the reference's drivers exchange the views chosen
by set cover / the TSP path (tests/SetCover.cpp:316-318); real trajectories can be fed to dmf_forward / dmf_reverse
through it."""
from __future__ import annotations

import numpy as np


def _fmt(x: float) -> str:
    return "%g" % float(np.float32(x))          # std::ostream << float: precision 6, shortest of %e/%f


def write_camera_locations(filename: str, poses) -> None:
    """writeCameraLocations(filename, transformations) (FileRoutines.hpp:98-112); poses: (n,12) or (n,3,4) or (n,4,4)."""
    a = np.asarray(poses, np.float32)
    if a.ndim == 3 and a.shape[1:] == (4, 4):
        a = a[:, :3, :]
    a = a.reshape(-1, 3, 4)
    with open(filename, "w") as f:
        f.write(f"{len(a)}\n")
        for T in a:
            for row in T:
                f.write(",".join(_fmt(v) for v in row) + "\n")


def read_camera_locations(filename: str) -> np.ndarray:
    """readCameraLocations(filename) (FileRoutines.hpp:69-96): returns (n,12) float32 (std::stof per field)."""
    with open(filename) as f:
        lines = f.read().splitlines()
    n = int(lines[0])
    out = np.zeros((n, 3, 4), np.float32)
    for i in range(n):
        for j in range(3):
            fields = lines[1 + 3 * i + j].split(",")
            if len(fields) != 4:
                raise ValueError(f"{filename}: pose {i} row {j}: expected 4 comma-separated numbers")   # assert(numbers.size()==4)
            out[i, j] = [np.float32(float(x)) for x in fields]
    return out.reshape(n, 12)
