"""Drop-in for the reference's compiled `_gridencoder` extension.

Put this directory on sys.path *before* importing the reference's `gridencoder` package and its
`import _gridencoder as _backend` (gridencoder/grid.py:L9-12) resolves to the B200 kernels; the reference's
grid.py, models.py, train.py and eval.py then run unchanged.  See INTEGRATION.md."""
from ucnerf_b200.gridencoder.backend import (grad_total_variation, grid_encode_backward,  # noqa: F401
                                             grid_encode_forward)
