"""TEST INFRASTRUCTURE - the named parity cases (config + seeds) shared by oracle/make_golden.py, tests/ and
bench.py.  Config labels follow SURVEY.md section 8(d)."""
from oracle import ucnerf_oracle as O


def three_level_config():
    """The reference's *class defaults* for sampling (models.py:L33-35,L55): 2 proposal levels x 64 samples with
    512 / 2048 grids (L=6 / L=8) + 32 NeRF samples; MLP shapes as waymo.gin."""
    return O.HotPathConfig(num_prop_samples=64, num_nerf_samples=32, prop_grids=[O.GridSpec(512), O.GridSpec(2048)])


CASES = {
    # name: (config factory, n_rays, weight seed, ray seed)
    "config1": (O.config1, 512, 0, 0),
    "waymo": (O.waymo_config, 128, 0, 1),
    "three_level": (three_level_config, 64, 2, 3),
    "target1024": (O.target_config, 16, 0, 4),
}


def make_case(name, n_rays=None):
    factory, n, wseed, rseed = CASES[name]
    cfg = factory()
    params = O.init_params(cfg, seed=wseed)
    batch = O.synthetic_rays(n_rays or n, seed=rseed)
    return cfg, params, batch


def param_checksums(params):
    import torch
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in params.items()
            if v.dtype == torch.float32}
