"""hagrid_b200 — B200-native irregular-grid build + traversal behind the
reference's (cg-saarland/hagrid) build_grid()/traverse_grid() interface."""
from .api import (HIT_PRIM_ID, HIT_STEPS, CELL_DTYPE, HIT_DTYPE, RAY_DTYPE, SMALL_CELL_DTYPE, TRI_DTYPE,
                  GridInfo, HagridError, Library, Scene, library, make_camera, parse_obj)

__all__ = ["HIT_PRIM_ID", "HIT_STEPS", "CELL_DTYPE", "HIT_DTYPE", "RAY_DTYPE", "SMALL_CELL_DTYPE", "TRI_DTYPE",
           "GridInfo", "HagridError", "Library", "Scene", "library", "make_camera", "parse_obj"]
