"""ucnerf_b200 - B200-native (sm_100a) implementation of UC-NeRF's forward-render hot path.

Public surface (mirrors the reference):
    ucnerf_b200.gridencoder.GridEncoder      <- nerf/gridencoder/grid.py::GridEncoder
    ucnerf_b200.render.render_image          <- nerf/internal/models.py::render_image
    ucnerf_b200.render.HotPathModel.forward  <- nerf/internal/models.py::Model.forward (eval path)
    ucnerf_b200/dropin/_gridencoder.py       <- the compiled `_gridencoder` extension module
All compute runs in libucnerf_b200.so (hand-written CUDA, C ABI in include/ucnerf_b200.h)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
