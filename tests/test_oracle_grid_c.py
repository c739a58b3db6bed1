"""CPU: the C + OpenMP restatement of kernel_grid (oracle/grid_cpu.c, the grid backend of the CPU reference arm) is
bit-identical to the numpy restatement (oracle/ucnerf_oracle.py::grid_encode_forward), which is pinned to the reference
kernel's own outputs (tests/golden/gridref_*.npz)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import build_c, ucnerf_oracle as O

CASES = [
    # B, D, C, L, log2T, desired, gridtype, align_corners, interp
    (3001, 3, 4, 10, 19, 8192, 0, False, 0),
    (3001, 3, 4, 6, 19, 512, 0, False, 0),
    (1000, 3, 2, 8, 15, 1024, 0, False, 0),
    (1000, 3, 8, 4, 14, 128, 0, False, 1),
    (1000, 3, 1, 4, 14, 128, 1, False, 0),
    (1000, 3, 4, 4, 12, 128, 0, True, 0),
    (777, 2, 2, 8, 12, 2048, 0, False, 0),
    (513, 4, 2, 4, 12, 64, 0, False, 0),
    (300, 5, 1, 3, 10, 32, 0, False, 1),
]


@pytest.mark.parametrize("B,D,C,L,T,desired,gridtype,ac,interp", CASES)
def test_c_restatement_equals_numpy_restatement(B, D, C, L, T, desired, gridtype, ac, interp):
    lay = O.grid_layout(L, C, 16, desired, T, input_dim=D, align_corners=ac)
    S = float(np.log2(lay["per_level_scale"]))
    rng = np.random.default_rng(B + D + C)
    x = rng.random((B, D), dtype=np.float32)
    x[0] = 0.0
    x[1] = 1.0
    x[2, 0] = -0.25
    x[3, D - 1] = 1.5
    emb = (rng.random((int(lay["offsets"][-1]), C), dtype=np.float32) * 2 - 1)
    ref, _ = O.grid_encode_forward(x, emb, lay["offsets"], B, D, C, L, S, 16, False, gridtype, ac, interp)
    got = build_c.grid_encode_forward(x, emb, lay["offsets"], B, D, C, L, S, 16, gridtype, ac, interp)
    if float(S).is_integer():
        assert np.array_equal(got, ref)
    else:   # exp2f(level * S): glibc and numpy may differ by one ulp of the level scale (so does CUDA's exp2f)
        assert np.abs(got - ref).max() < 2e-4


def test_c_restatement_against_reference_kernel_vectors():
    """Vectors produced by the reference's own kernel_grid on a B200 (oracle/make_gridref_golden.py)."""
    import torch
    g = load_golden("gridref_hash_d3c4_waymo")
    B, D, C, L = (int(g[k]) for k in ("B", "D", "C", "L"))
    gen = torch.Generator().manual_seed(int(g["emb_seed"]))     # same generator call sequence as the golden script
    torch.rand((B, D), generator=gen)
    emb = (torch.rand((int(g["offsets"][-1]), C), generator=gen) * 2 - 1).numpy()
    got = build_c.grid_encode_forward(g["inputs"], emb, g["offsets"], B, D, C, L, float(g["S"]), int(g["H"]))
    assert np.abs(got - g["outputs"]).max() < 2e-6
    assert (got == g["outputs"]).mean() > 0.99
