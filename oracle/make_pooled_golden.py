"""TEST INFRASTRUCTURE - golden vectors for the pooled hash-grid encode (training front end of MLP.predict_density)
from the REFERENCE's own modules running on CPU (oracle/ref_shim.py: unmodified internal/models.py, coord.py,
render.py, gridencoder/grid.py + autograd; the native `_gridencoder` is the pinned CPU stand-in).

For the proposal MLP (L=6) and the NeRF MLP (L=10) of the waymo.gin model:
  means, stds      <- reference render.cast_rays on seeded rays / sampled intervals (some far outside the unit ball)
  features         <- input of `density_layer` inside reference MLP.predict_density (forward pre-hook)
  coord            <- third return value of predict_density
  grad_features    <- d loss / d features for loss = sum(x * G), G seeded
  grad_embeddings  <- embeddings.grad after loss.backward(), stored sparsely (touched rows only)

    python oracle/make_pooled_golden.py        # writes tests/golden/pooled_encode.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_shim, ucnerf_oracle as O  # noqa: E402

N_RAYS, N_INTERVALS = 6, 8    # 48 intervals x 6 points per MLP


def main():
    cfg, params, batch = cases.make_case("waymo", N_RAYS)
    model, _ = ref_shim.build_reference_model(cfg, params)
    R = ref_shim.load_reference()
    g = torch.Generator().manual_seed(11)
    # interval fenceposts in metric distance: dense near the camera, sparse far away (contraction on both sides of |x| = 1)
    t = torch.sort(torch.rand((N_RAYS, N_INTERVALS + 1), generator=g) ** 2 * 7.5 + 0.02, dim=-1).values
    with ref_shim.inject_rand_vec(batch["rand_vec"]):
        means, stds, _ = R.render.cast_rays(t, batch["origins"], batch["directions"], batch["cam_dirs"], batch["radii"],
                                            False, std_scale=0.5)
    out = {"means": means.numpy(), "stds": stds.numpy()}
    for tag, mlp in (("prop", model.prop_mlp_0), ("nerf", model.nerf_mlp)):
        captured = {}

        def hook(_mod, args):
            captured["features"] = args[0]
            args[0].retain_grad()

        h = mlp.density_layer.register_forward_pre_hook(hook)
        mlp.encoder.embeddings.grad = None
        raw, x, coord = mlp.predict_density(means, stds)
        h.remove()
        G = torch.randn(x.shape, generator=g)
        (x * G).sum().backward()
        feats = captured["features"]
        ge = mlp.encoder.embeddings.grad
        rows = torch.nonzero(ge.abs().sum(-1) > 0).reshape(-1)
        out.update({f"{tag}_features": feats.detach().numpy(), f"{tag}_coord": coord.detach().numpy(),
                    f"{tag}_grad_features": feats.grad.numpy(), f"{tag}_grad_rows": rows.numpy().astype(np.int64),
                    f"{tag}_grad_vals": ge[rows].numpy()})
        print(tag, "features", tuple(feats.shape), "touched rows", rows.numel(), "|grad|max", float(ge.abs().max()))
    path = os.path.join(ROOT, "tests", "golden", "pooled_encode.npz")
    np.savez_compressed(path, **out, n_rays=N_RAYS, weight_seed=np.int64(cases.CASES["waymo"][2]),
                        ray_seed=np.int64(cases.CASES["waymo"][3]))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
