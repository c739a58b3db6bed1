"""TEST INFRASTRUCTURE - pins the oracle's brightness-correction restatement (apply_brightness / brightness_affine)
against the UNMODIFIED reference `Model.forward` with `config.brightness_correction = True` (models.py:L94-95,
L339-363; extrinsic_optimizer.py:L4-49) and writes tests/golden/brightness.npz.  Build container only:

    python -m oracle.make_brightness_golden"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import cases, ref_shim, ucnerf_oracle as O  # noqa: E402


def main():
    assert ref_shim.available(), "needs /root/reference"
    cfg, params, batch = cases.make_case("waymo", 96)
    R = ref_shim.load_reference()
    # reference Model with the brightness head: Config fields set before construction (models.py:L94-95)
    orig = R.configs.Config
    conf_patch = dict(brightness_correction=True, training_views=12, model_sky=False)

    class Conf(orig):
        pass

    model, conf = ref_shim.build_reference_model(cfg, params)  # plain model (no head) for the state_dict layout
    conf2 = orig()
    for k, v in conf_patch.items():
        setattr(conf2, k, v)
    torch.manual_seed(5)
    m2 = R.models.Model(config=conf2)
    sd = m2.state_dict()
    head = {k: v.clone() for k, v in sd.items() if k.startswith("brightness_corr.")}
    head["brightness_corr.latent_code"] = torch.randn_like(head["brightness_corr.latent_code"]) * 0.5
    new = {}
    for k, v in sd.items():
        if k in params:
            new[k] = params[k].to(v.dtype)
        elif k in head:
            new[k] = head[k]
        else:
            assert k.endswith(".idx"), k
            new[k] = v
    m2.load_state_dict(new)
    m2.eval()
    cam = 7
    b = {k: v for k, v in batch.items() if k != "rand_vec"}
    with torch.no_grad(), ref_shim.inject_rand_vec(batch["rand_vec"]):
        rr, _ = m2(False, b, train_frac=1.0, compute_extras=True, zero_glo=True, eval_camidx=torch.tensor(cam))
    ref_rgb = rr[-1]["rgb"]
    ref_aff = rr[-1]["affine_trans"]
    orr, _ = O.model_forward(params, cfg, batch)
    aff = O.brightness_affine(head, cam, n_rays=96)
    assert torch.equal(aff, ref_aff[0]), (aff - ref_aff[0]).abs().max()
    mine = O.apply_brightness(orr[-1]["rgb"], aff)
    assert torch.equal(mine, ref_rgb), (mine - ref_rgb).abs().max()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "brightness.npz")
    np.savez_compressed(path, affine=aff.numpy(), rgb_plain=orr[-1]["rgb"].numpy(), rgb_corrected=ref_rgb.numpy(),
                        n_rays=np.int64(96), cam=np.int64(cam))
    print("oracle == reference (affine and corrected rgb, bit-for-bit); wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
