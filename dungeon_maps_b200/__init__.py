"""dungeon_maps_b200 — B200-native (sm_100a) drop-in for the depth → top-down hot path of
Ending2015a/dungeon_maps: same public API (`MapProjector`, `MapBuilder`, `TopdownMap`, the
functional projection / transform functions, `utils`), bodies replaced by fused CUDA kernels
behind a C ABI (include/dungeon_maps_b200.h).  No CPU fallback.

    import dungeon_maps_b200 as dmap
    proj = dmap.MapProjector(width=640, height=480, hfov=1.22, ...)
    topdown, mask, height = proj.orth_project(depth, value_map=semantics, get_height_map=True)
"""
from . import utils
from . import maps
from . import synth

from .maps import *  # noqa: F401,F403

__version__ = '0.1.0'
