"""dmf_b200 -- B200-native RayTracingEngine hot path.  Compute lives in libdmf_b200.so."""
from . import scenes
from .comm import Comm
from ._lib import DmfError, LIB_PATH, SYMBOLS, load
from .engine import (FWD_CARVE, FWD_NO_COUNTERS, FWD_NO_SKIP, FWD_TWO_PROBE, GRID_AUTO, GRID_BIT, GRID_BYTE, MODE_CLASSIFY, MODE_GOOD_POINTS, MODE_MARK, MODE_MINIMUM, MODE_POINTS, NO_VOXEL,
                     Camera, Context, RayTracingEngine, VoxelVolume, bits_to_indices, greedySetCover, optimizeCameraPosition, repositionCamerasSampled, reposition_from_minimum, willCollide)

__all__ = ["scenes", "Comm", "DmfError", "LIB_PATH", "SYMBOLS", "load", "Camera", "Context", "RayTracingEngine", "VoxelVolume",
           "greedySetCover", "willCollide", "optimizeCameraPosition", "repositionCamerasSampled", "reposition_from_minimum", "bits_to_indices", "MODE_POINTS", "MODE_GOOD_POINTS", "MODE_CLASSIFY", "MODE_MARK", "MODE_MINIMUM",
           "GRID_BIT", "GRID_BYTE", "GRID_AUTO", "NO_VOXEL", "FWD_NO_SKIP", "FWD_TWO_PROBE", "FWD_CARVE", "FWD_NO_COUNTERS"]
