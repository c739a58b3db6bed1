"""TEST INFRASTRUCTURE - pins the oracle's restatement of the sky head + brightness correction (sky_render_rays,
combine_heads) against the UNMODIFIED reference `Model.forward` with `config.model_sky = True` and
`config.brightness_correction = True` (models.py:L84-95, L326-363, L743-904), i.e. the configuration of the shipped
scripts/train_waymo.sh, and writes tests/golden/heads.npz.  Build container only:

    python -m oracle.make_heads_golden"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import cases, ref_shim, ucnerf_oracle as O  # noqa: E402

N = 48


def main():
    assert ref_shim.available(), "needs /root/reference"
    cfg, params, batch = cases.make_case("waymo", N)
    R = ref_shim.load_reference()
    ref_shim.build_reference_model(cfg, params)      # sets the class attributes of Model / MLPs for this config
    conf = R.configs.Config()
    conf.brightness_correction, conf.model_sky, conf.training_views = True, True, 9
    m = R.models.Model(config=conf)
    sd = m.state_dict()
    heads = cases.make_heads(seed=3, n_views=9)
    new = {}
    for k, v in sd.items():
        src = params.get(k, heads.get(k))
        if src is None:
            assert k.endswith('.idx'), k
            src = v
        assert tuple(src.shape) == tuple(v.shape), (k, src.shape, v.shape)
        new[k] = src.to(v.dtype)
    m.load_state_dict(new)
    m.eval()
    cam = 4
    b = {k: v for k, v in batch.items() if k != "rand_vec"}
    with torch.no_grad(), ref_shim.inject_rand_vec(batch["rand_vec"]):
        rr, rh = m(False, b, train_frac=1.0, compute_extras=True, zero_glo=True, eval_camidx=torch.tensor(cam))
    orr, oh = O.model_forward(params, cfg, batch)
    sky = O.sky_render_rays(heads, batch["origins"], batch["directions"], batch["far"], batch["cam_dirs"])
    assert torch.equal(sky, rr[-1]["sky_rgbs"]), (sky - rr[-1]["sky_rgbs"]).abs().max()
    aff = O.brightness_affine(heads, cam, n_rays=N)
    aff_sky = O.brightness_affine({k.replace("sky_latent_code", "latent_code"): v for k, v in heads.items()
                                   if "brightness_corr.latent_code" not in k}, cam, n_rays=N)
    assert torch.equal(aff, rr[-1]["affine_trans"][0]) and torch.equal(aff_sky, rr[-1]["affine_trans_sky"][0])
    rgb = O.combine_heads(orr[-1]["rgb"], oh[-1]["weights"], aff, sky, aff_sky)
    assert torch.equal(rgb, rr[-1]["rgb"]), (rgb - rr[-1]["rgb"]).abs().max()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "heads.npz")
    np.savez_compressed(path, n_rays=np.int64(N), cam=np.int64(cam), affine=aff.numpy(), affine_sky=aff_sky.numpy(),
                        sky_rgbs=sky.numpy(), rgb_plain=orr[-1]["rgb"].numpy(), rgb_final=rr[-1]["rgb"].numpy(),
                        weights=oh[-1]["weights"].numpy())
    print("oracle == reference bit-for-bit (sky_rgbs, both affines, combined rgb); wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
