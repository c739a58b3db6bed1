"""TEST INFRASTRUCTURE - vectors from the REFERENCE's own CUDA kernel.

Run on the GPU box (needs oracle/_ref/_gridencoder_ref.so, built by oracle/build_ref.py from
/root/reference/nerf/gridencoder/src/*.cu where they lie): executes the reference `grid_encode_forward`
(gridencoder.cu:L448-471) on seeded inputs and writes gpurun_out/gridref_<case>.npz.  The files are then
committed under tests/golden/ and pin the oracle's restatement of kernel_grid on CPU
(tests/test_oracle_grid.py::test_oracle_matches_reference_cuda_kernel_vectors).

    python oracle/make_gridref_golden.py [outdir]"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ucnerf_oracle as O  # noqa: E402

CASES = {
    # name: B, D, C, L, log2T, desired, gridtype, align_corners, interp, dy_dx
    "hash_d3c4_waymo": (384, 3, 4, 10, 21, 8192, 0, False, 0, False),
    "hash_d3c4_T15_dydx": (256, 3, 4, 4, 15, 128, 0, False, 0, True),
    "hash_d3c2_smooth": (256, 3, 2, 5, 12, 256, 0, False, 1, True),
    "tiled_d2c8": (256, 2, 8, 4, 10, 128, 1, False, 0, False),
    "align_d3c1": (256, 3, 1, 4, 12, 128, 0, True, 0, False),
}


def main(outdir):
    so = os.path.join(ROOT, "oracle", "_ref", "_gridencoder_ref.so")
    spec = importlib.util.spec_from_file_location("_gridencoder_ref", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    os.makedirs(outdir, exist_ok=True)
    for name, (B, D, C, L, T, desired, gridtype, ac, interp, dy) in CASES.items():
        lay = O.grid_layout(L, C, 16, desired, T, input_dim=D, align_corners=ac)
        offsets = torch.from_numpy(lay["offsets"])
        S = float(np.log2(lay["per_level_scale"]))
        g = torch.Generator().manual_seed(len(name) * 1000 + B)
        x = torch.rand((B, D), generator=g)
        x[0] = 0.0
        x[1] = 1.0
        x[2, 0] = -0.5
        x[3, D - 1] = 1.25
        n_emb = int(offsets[-1])
        # keep the fixture small: only the entries the kernel can touch matter, so store the table sparsely
        emb = torch.rand((n_emb, C), generator=g) * 2 - 1
        out = torch.empty(L, B, C, device="cuda")
        dd = torch.empty(B, L * D * C, device="cuda") if dy else None
        ref.grid_encode_forward(x.cuda(), emb.cuda(), offsets.cuda(), out, B, D, C, L, S, 16, dd, gridtype, ac, interp)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(outdir, f"gridref_{name}.npz"), inputs=x.numpy(),
                            emb_seed=np.int64(len(name) * 1000 + B), offsets=offsets.numpy(), outputs=out.cpu().numpy(),
                            dy_dx=(dd.cpu().numpy() if dy else np.zeros(0, np.float32)), B=B, D=D, C=C, L=L, S=S, H=16,
                            gridtype=gridtype, align_corners=ac, interp=interp, has_dy_dx=dy)
        print("wrote", name, out.shape)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out"))
