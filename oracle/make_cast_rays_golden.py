"""TEST INFRASTRUCTURE - golden vectors for the stand-alone cast_rays op from the REFERENCE's own render.cast_rays
(internal/render.py:L94-152) on CPU, rand=False and rand=True; torch.rand_like / torch.randn_like are patched to
return the stored draws (flip mask draw, rotation draw, rand_vec - in the reference's call order).

    python oracle/make_cast_rays_golden.py        # writes tests/golden/cast_rays.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, ucnerf_oracle as O  # noqa: E402


def main():
    R = ref_shim.load_reference()
    N, S = 24, 16
    g = torch.Generator().manual_seed(31)
    batch = O.synthetic_rays(N, seed=9)
    batch["radii"] = batch["radii"] * (0.5 + torch.rand((N, 1), generator=g))
    batch["directions"] = batch["directions"] * (0.7 + torch.rand((N, 1), generator=g))      # not unit norm
    tdist = torch.sort(torch.rand((N, S + 1), generator=g) ** 2 * 7.9 + 0.01, dim=-1).values
    flip01, rot01 = torch.rand((N, S), generator=g), torch.rand((N, S), generator=g)
    rand_vec = torch.randn((N, 3), generator=g)
    out = {"tdist": tdist.numpy(), "flip01": flip01.numpy(), "rot01": rot01.numpy(), "rand_vec": rand_vec.numpy()}
    for k in ("origins", "directions", "cam_dirs", "radii"):
        out[k] = batch[k].numpy()
    o_rand, o_randn = torch.rand_like, torch.randn_like
    for tag, rand in (("det", False), ("rand", True)):
        queue = [flip01, rot01]
        torch.rand_like = lambda t, *a, **k: queue.pop(0).clone()
        torch.randn_like = lambda t, *a, **k: rand_vec.clone()
        try:
            means, stds, ts = R.render.cast_rays(tdist, batch["origins"], batch["directions"], batch["cam_dirs"],
                                                 batch["radii"], rand, std_scale=0.5)
        finally:
            torch.rand_like, torch.randn_like = o_rand, o_randn
        out[f"means_{tag}"], out[f"stds_{tag}"], out[f"ts_{tag}"] = means.numpy(), stds.numpy(), ts.numpy()
        print(tag, tuple(means.shape), tuple(stds.shape), tuple(ts.shape), "draws left:", len(queue))
    path = os.path.join(ROOT, "tests", "golden", "cast_rays.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
