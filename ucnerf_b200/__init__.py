"""ucnerf_b200 - B200-native (sm_100a) implementation of UC-NeRF's forward-render hot path.

Public surface (mirrors the reference):
    ucnerf_b200.gridencoder.GridEncoder      <- nerf/gridencoder/grid.py::GridEncoder
    ucnerf_b200.render.render_image          <- nerf/internal/models.py::render_image
    ucnerf_b200.render.HotPathModel.forward  <- nerf/internal/models.py::Model.forward (eval path)
    ucnerf_b200.models.Model / NerfMLP / PropMLP <- the reference classes (same attributes / parameter names; forward
                                                dispatches training -> train_forward.level_loop, eval -> the fused path)
    ucnerf_b200/dropin/_gridencoder.py       <- the compiled `_gridencoder` extension module
Ops for the reference's training step (INTEGRATION.md seams 5-10):
    ucnerf_b200.train_forward.level_loop           <- the level loop of Model.forward(rand=True) (models.py:L126-311)
    ucnerf_b200.gridencoder.pooled.pooled_encode   <- MLP.predict_density front end (models.py:L485-496), fwd + bwd
    ucnerf_b200.stepfun.resample_level             <- max_dilate_weights + sample_intervals (models.py:L156-205)
    ucnerf_b200.render_train.cast_rays             <- render.cast_rays (render.py:L94-152), rand on / off
    ucnerf_b200.render_train.composite             <- compute_alpha_weights + acc / rgb (render.py:L155-205), fwd + bwd
    ucnerf_b200.gridencoder.optim.GridAdam         <- hash-decay + Adam + zero_grad for the tables
All compute runs in libucnerf_b200.so (hand-written CUDA, C ABI in include/ucnerf_b200.h)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
