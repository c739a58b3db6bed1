"""Synthetic workloads for bench.py / smoke(): random-init weights under the reference state_dict names and
pinhole-camera ray batches with the keys of the reference's ray dict (internal/datasets.py:L386-476,
camera_utils.py:L611-632).  There is no dataset or checkpoint in the container, so benchmarks use these."""
import math
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import torch


@dataclass
class Workload:
    name: str
    num_prop_samples: int = 128
    num_nerf_samples: int = 32
    prop_desired: List[int] = field(default_factory=lambda: [512])
    nerf_desired: int = 8192
    log2_hashmap_size: int = 21
    bottleneck_width: int = 256
    net_width_viewdirs: int = 256
    height: int = 600
    width: int = 800

    @property
    def samples_per_ray(self):
        return self.num_prop_samples * len(self.prop_desired) + self.num_nerf_samples

    @staticmethod
    def grid_levels(desired, base=16):
        return int(np.log(desired / base) / np.log(2)) + 1

    def gather_bytes_per_ray(self):
        """6 points x L levels x 8 corners x C=4 x 4 B per ray-sample."""
        b = sum(self.num_prop_samples * 768 * self.grid_levels(d) for d in self.prop_desired)
        return b + self.num_nerf_samples * 768 * self.grid_levels(self.nerf_desired)

    def mlp_flops_per_ray(self):
        f = 0
        for d in self.prop_desired:
            lc = 4 * self.grid_levels(d)
            f += self.num_prop_samples * 2 * (lc * 64 + 64)
        lc = 4 * self.grid_levels(self.nerf_desired)
        bw, w, nd = self.bottleneck_width, self.net_width_viewdirs, 27
        f += self.num_nerf_samples * 2 * (lc * 64 + 64 * bw + (bw + nd) * w + (w + bw + nd) * w + w * 3)
        return f


WORKLOADS = {
    # BASELINE.json configs[1] as SURVEY.md section 8(d) restates it: one 800x600 eval frame, waymo.gin shapes
    "eval_800x600_waymo_gin": Workload("eval_800x600_waymo_gin"),
    # BASELINE.json configs[2]/[3] image size: one full-resolution Waymo frame (datasets.py:L896-897), hot path only
    "eval_1920x1280_waymo_gin": Workload("eval_1920x1280_waymo_gin", height=1280, width=1920),
    # north_star "1024 samples/ray synthetic rays": 512 prop + 512 fine, 65,536 rays (256x256)
    "target_1024spp": Workload("target_1024spp", num_prop_samples=512, num_nerf_samples=512, height=256, width=256),
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    "config1_cpu_case": Workload("config1_cpu_case", num_prop_samples=32, num_nerf_samples=32, prop_desired=[128],
                                 nerf_desired=128, log2_hashmap_size=15, bottleneck_width=32, net_width_viewdirs=32,
                                 height=64, width=64),
}


def _layout(levels, desired, log2_T, base=16, C=4):
    pls = np.exp2(np.log2(desired / base) / (levels - 1))
    sizes, res = [], []
    for i in range(levels):
        r = int(np.ceil(base * pls ** i)) + 1
        sizes.append(int(np.ceil(min(2 ** log2_T, r ** 3) / 8) * 8))
        res.append(r)
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), np.array(res, dtype=np.int32)


def synthetic_state_dict(wl: Workload, seed=0, emb_range=0.5) -> Dict[str, torch.Tensor]:
    """nn.Linear default init / kaiming_uniform_ for the view layers (models.py:L478) / U(+-0.5) embeddings (the
    reference's 1e-4 init renders a featureless scene), on CPU; move to the GPU with .cuda()."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def linear(prefix, fin, fout, kaiming=False):
        sd[prefix + '.weight'] = uni((fout, fin), math.sqrt(6.0 / fin) if kaiming else 1.0 / math.sqrt(fin))
        sd[prefix + '.bias'] = uni((fout,), 1.0 / math.sqrt(fin))

    def encoder(prefix, desired):
        L = Workload.grid_levels(desired)
        offsets, sizes = _layout(L, desired, wl.log2_hashmap_size)
        sd[prefix + '.offsets'] = torch.from_numpy(offsets)
        sd[prefix + '.grid_sizes'] = torch.from_numpy(sizes)
        sd[prefix + '.embeddings'] = uni((int(offsets[-1]), 4), emb_range)
        return 4 * L

    for i, d in enumerate(wl.prop_desired):
        fin = encoder(f'prop_mlp_{i}.encoder', d)
        linear(f'prop_mlp_{i}.density_layer.0', fin, 64)
        linear(f'prop_mlp_{i}.density_layer.2', 64, 1)
    fin = encoder('nerf_mlp.encoder', wl.nerf_desired)
    linear('nerf_mlp.density_layer.0', fin, 64)
    linear('nerf_mlp.density_layer.2', 64, wl.bottleneck_width)
    d_in = wl.bottleneck_width + 27
    linear('nerf_mlp.lin_second_stage_0', d_in, wl.net_width_viewdirs, kaiming=True)
    linear('nerf_mlp.lin_second_stage_1', wl.net_width_viewdirs + d_in, wl.net_width_viewdirs, kaiming=True)
    linear('nerf_mlp.rgb_layer', wl.net_width_viewdirs, 3)
    return sd


def pinhole_camera(height, width, seed=0, near=0.0, far=8.0, focal=None):
    """(pixtocam, camtoworld, width, height, near, far) of the synthetic camera `pinhole_rays(seed)` uses, as float32
    values like the reference's Waymo loader holds them (datasets.py:L672,L855-857)."""
    rng = np.random.default_rng(seed)
    focal = focal or 2000.0 * width / 1920.0
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    origin = rng.uniform(-0.1, 0.1, 3)
    K = np.array([[focal, 0, width * 0.5], [0, focal, height * 0.5], [0, 0, 1]], dtype=np.float32)
    pose = np.concatenate([R, origin[:, None]], 1).astype(np.float32)
    return np.linalg.inv(K), pose, width, height, near, far


def pinhole_rays(height, width, seed=0, near=0.0, far=8.0, focal=None) -> Dict[str, torch.Tensor]:
    """One pinhole camera inside the scene (OpenGL convention, looks along -z of a random pose): flat [H*W, .] CPU
    tensors with the reference's ray-dict keys + the cone-basis `rand_vec` (render.py:L140)."""
    rng = np.random.default_rng(seed)
    focal = focal or 2000.0 * width / 1920.0
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    origin = rng.uniform(-0.1, 0.1, 3)
    px, py = np.meshgrid(np.arange(width, dtype=np.float64) + 0.5, np.arange(height, dtype=np.float64) + 0.5)

    def cam_dirs_at(u, v):
        return np.stack([(u - width / 2) / focal, -(v - height / 2) / focal, -np.ones_like(u)], -1) @ R.T

    d = cam_dirs_at(px, py)
    dx = cam_dirs_at(px + 1, py)
    viewdirs = d / np.linalg.norm(d, axis=-1, keepdims=True)
    vx = dx / np.linalg.norm(dx, axis=-1, keepdims=True)
    radii = np.linalg.norm(vx - viewdirs, axis=-1, keepdims=True) * 2 / np.sqrt(12)
    n = height * width
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(n, -1), dtype=np.float32))
    g = torch.Generator().manual_seed(seed + 999)
    return dict(origins=t(np.broadcast_to(origin, d.shape)), directions=t(d), viewdirs=t(viewdirs),
                cam_dirs=t(np.broadcast_to(-R[:, 2], d.shape)), radii=t(radii),
                near=torch.full((n, 1), float(near)), far=torch.full((n, 1), float(far)),
                rand_vec=torch.randn((n, 3), generator=g))


def synthetic_heads(seed=0, n_views=9) -> Dict[str, torch.Tensor]:
    """Random-init weights of the sky head (`skynerf.*`, internal/models.py:L84-92,L743-795) and of the brightness
    affines of one camera, under the reference state_dict names (nn.Linear-style init), on CPU."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def linear(name, fin, fout):
        b = 1.0 / math.sqrt(fin)
        sd[name + '.weight'] = (torch.rand((fout, fin), generator=g) * 2 - 1) * b
        sd[name + '.bias'] = (torch.rand((fout,), generator=g) * 2 - 1) * b

    for i in range(8):
        linear(f'skynerf.pts_linears.{i}', 3 if i == 0 else (259 if i == 5 else 256), 256)
    linear('skynerf.views_linears.0', 283, 128)
    linear('skynerf.feature_linear', 256, 256)
    linear('skynerf.alpha_linear', 256, 1)
    linear('skynerf.rgb_linear', 128, 3)
    eye = torch.cat([torch.eye(3), torch.zeros(3, 1)], 1)
    sd['affine'] = eye + 0.05 * torch.randn((3, 4), generator=g)
    sd['affine_sky'] = eye + 0.05 * torch.randn((3, 4), generator=g)
    return sd


def make_renderer(wl: Workload, state_dict, device=None):
    from .render import HotPathModel
    dev = device or f"cuda:{torch.cuda.current_device()}"
    sd = {k: v.to(dev) for k, v in state_dict.items()}
    return HotPathModel(sd, num_prop_samples=wl.num_prop_samples, num_nerf_samples=wl.num_nerf_samples,
                        num_prop_levels=len(wl.prop_desired), bottleneck_width=wl.bottleneck_width,
                        net_width_viewdirs=wl.net_width_viewdirs, device=dev)
