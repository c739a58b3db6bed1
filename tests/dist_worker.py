"""Worker for the world_size-2 tests of render_image's tile shard + single all-gather.
--backend nccl : real renderer on 2 GPUs, compared with a single-GPU render of the whole image.
--backend gloo : CPU stand-in renderer (a deterministic function of the rays) so the host-side shard / pad /
                 gather / reshape logic is exercised without a GPU."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ucnerf_b200 import render as R  # noqa: E402


class FakeRenderer:
    """Same surface as HotPathModel.render_rays, on CPU: outputs are simple functions of the ray origin."""
    num_levels, samples, vis_num_rays = 2, [8, 4], 4
    device = torch.device("cpu")
    options = {}

    def set_option(self, key, value):          # render_image announces the row length of an image tile and takes it back
        FakeRenderer.options[key] = int(value)

    def render_rays(self, batch, train_frac, rand_vec, want):
        o = batch["origins"]
        n = o.shape[0]
        out = {}
        packed = torch.zeros(n, R.PACKED_WIDTH)
        packed[:, 0:3] = o
        packed[:, 3] = o.sum(-1)
        packed[:, 4] = o[:, 0] * 2
        for i in range(5, 10):
            packed[:, i] = o[:, 1] + i
        out["packed"] = packed
        for l, S in enumerate(self.samples):
            out[f"sdist_{l}"] = torch.linspace(0, 1, S + 1).expand(n, S + 1).contiguous()
            out[f"weights_{l}"] = o[:, :1].expand(n, S).contiguous()
        out["sample_rgb"] = o[:, None, :].expand(n, self.samples[-1], 3).contiguous()
        out["sample_coord"] = (o[:, None, :] * 0.5).expand(n, self.samples[-1], 3).contiguous()
        return {k: out[k] for k in want}


def render_rays(ray_batch, network_fn, **_):
    """Stand-in for the reference's module-level render_rays (models.py:L849-904): any per-ray function will do here."""
    return {"rgb_map": network_fn(ray_batch[:, 0:3], ray_batch[:, 3:6]) + ray_batch[:, 6:7]}


class _Sky(torch.nn.Module):
    def forward(self, o, d):
        return torch.sin(o * 3.0) + 0.5 * d


class _Bright(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, indices):
        n = indices.reshape(-1).shape[0]
        a = torch.tensor([[1.1, 0.0, 0.1, 0.01], [0.0, 0.9, 0.0, -0.02], [0.05, 0.0, 1.0, 0.0]])
        b = torch.tensor([[0.8, 0.1, 0.0, 0.03], [0.0, 1.2, 0.0, 0.0], [0.0, 0.0, 0.7, 0.05]])
        return a.expand(n, 3, 4), b.expand(n, 3, 4)


class HeadsModel(torch.nn.Module):   # `render_rays` above is resolved through this class's module
    def __init__(self):
        super().__init__()
        self.skynerf, self.brightness_corr = _Sky(), _Bright()


class Acc:
    def __init__(self):
        self.process_index = dist.get_rank()
        self.num_processes = dist.get_world_size()
        self.is_main_process = self.process_index == 0


class Cfg:
    render_chunk_size, vis_num_rays = 15000, 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    a = ap.parse_args()
    dist.init_process_group(a.backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W = 37, 53  # 1961 rays: not divisible by the world size -> padding path
    g = torch.Generator().manual_seed(0)
    if a.backend == "gloo":
        origins = torch.rand((H, W, 3), generator=g)
        batch = {k: origins for k in ("origins", "directions", "viewdirs", "cam_dirs")}
        batch.update({k: origins[..., :1] for k in ("radii", "near", "far")})
        img = R.render_image(None, Acc(), batch, False, 1.0, Cfg(), renderer=FakeRenderer(), return_weights=True,
                             rand_vec=torch.zeros(H * W, 3))
        assert torch.equal(img["rgb"], origins)
        assert torch.equal(img["depth"], origins.sum(-1)) and torch.equal(img["acc"], origins[..., 0] * 2)
        assert torch.equal(img["distance_percentile_95"], origins[..., 1] + 8)
        assert img["weights"].shape == (H, W, 4) and torch.equal(img["weights"][..., 0], origins[..., 0])
        assert img["coord"].shape == (H, W, 4, 3) and torch.equal(img["coord"][:, :, 2, :], origins * 0.5)
        assert len(img["ray_sdist"]) == 2 and img["ray_rgbs"][0].shape == (4, 8, 3)
        assert "ray_tile_width" not in FakeRenderer.options        # 53 pixels per row: no patch mapping announced
        # an image whose per-rank tiles start on rows (8 x 16 pixels over 2 ranks): the row length is announced for the tile's
        # render and taken back before the vis-ray render
        seen = []

        class Tiled(FakeRenderer):
            def render_rays(self, batch, train_frac, rand_vec, want):
                seen.append((batch["origins"].shape[0], FakeRenderer.options.get("ray_tile_width", 0)))
                return super().render_rays(batch, train_frac, rand_vec, want)

        o2 = torch.rand((8, 16, 3), generator=g)
        b2 = {k: o2 for k in ("origins", "directions", "viewdirs", "cam_dirs")}
        b2.update({k: o2[..., :1] for k in ("radii", "near", "far")})
        img2 = R.render_image(None, Acc(), b2, False, 1.0, Cfg(), renderer=Tiled(), rand_vec=torch.zeros(128, 3))
        assert torch.equal(img2["rgb"], o2)
        if 128 % world == 0 and (128 // world) % 16 == 0:
            assert seen[0] == (128 // world, 16) and seen[1][1] == 0 and FakeRenderer.options["ray_tile_width"] == 0, seen
        # heads of the shipped config on top (sky through the reference-module path, brightness affines): the sharded
        # result must equal the single-process one
        class HCfg(Cfg):
            model_sky, brightness_correction, ucnerf_reference_sky = True, True, True

        class FakeAffine(FakeRenderer):
            affine = None

            def set_rgb_affine(self, a):
                FakeAffine.affine = a

            def render_rays(self, batch, train_frac, rand_vec, want):
                out = super().render_rays(batch, train_frac, rand_vec, want)
                if FakeAffine.affine is not None and "packed" in out:
                    a = FakeAffine.affine
                    out["packed"][:, 0:3] = out["packed"][:, 0:3] @ a[:3, :3].T + a[:3, 3]
                return out

        hm = HeadsModel()
        img_h = R.render_image(hm, Acc(), batch, False, 1.0, HCfg(), renderer=FakeAffine(), eval_camidx=torch.tensor(2),
                               rand_vec=torch.zeros(H * W, 3))

        class One:
            process_index, num_processes, is_main_process = 0, 1, True

        ref_h = R.render_image(hm, One(), batch, False, 1.0, HCfg(), renderer=FakeAffine(), eval_camidx=torch.tensor(2),
                               rand_vec=torch.zeros(H * W, 3))
        for k in ("rgb", "sky_rgbs", "acc", "depth"):
            assert torch.allclose(img_h[k], ref_h[k], atol=1e-6), k
        assert img_h["sky_rgbs"].shape == (H, W, 3) and img_h["affine_trans_sky"].shape == (H, W, 3, 4)
        assert not torch.allclose(img_h["rgb"], img["rgb"])
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import cases, ucnerf_oracle as O
        from test_gpu_render import build_renderer
        cfg, params, _ = cases.make_case("waymo", 1)
        r = build_renderer(cfg, params)
        rays = O.synthetic_rays(H * W, seed=77)
        b2d = {k: v.reshape(H, W, -1).cuda() for k, v in rays.items()}
        rv = b2d["rand_vec"].reshape(-1, 3)
        img = R.render_image(None, Acc(), b2d, False, 1.0, Cfg(), renderer=r, rand_vec=rv)

        class One:
            process_index, num_processes, is_main_process = 0, 1, True

        ref = R.render_image(None, One(), b2d, False, 1.0, Cfg(), renderer=r, rand_vec=rv)
        for k in ("rgb", "depth", "acc", "distance_mean", "distance_median"):
            assert torch.equal(img[k], ref[k]), k
        # `img` came through the fused NVLink tile exchange (peer.PeerImage, the default on NCCL); the plain all-gather
        # path must give the same image, and so must several frames in a row (the two image buffers alternate)
        assert getattr(r, "_peer_image", None) is not None and r._peer_image[1] is not None, "peer exchange not active"

        class NoPeer(Cfg):
            ucnerf_peer_exchange = False

        img_nccl = R.render_image(None, Acc(), b2d, False, 1.0, NoPeer(), renderer=r, rand_vec=rv, return_weights=True)
        img_peer = R.render_image(None, Acc(), b2d, False, 1.0, Cfg(), renderer=r, rand_vec=rv, return_weights=True)
        for k in ("rgb", "depth", "acc", "distance_percentile_95", "weights", "coord"):
            assert torch.equal(img_peer[k], img_nccl[k]), k
        for k in ("rgb", "depth", "acc", "distance_percentile_95"):
            assert torch.equal(img_nccl[k], ref[k]), k
        rays2 = O.synthetic_rays(H * W, seed=78)
        b2 = {k: v.reshape(H, W, -1).cuda() for k, v in rays2.items()}
        for frame in range(3):
            got = R.render_image(None, Acc(), b2 if frame % 2 else b2d, False, 1.0, Cfg(), renderer=r,
                                 rand_vec=(b2 if frame % 2 else b2d)["rand_vec"].reshape(-1, 3))
            want = R.render_image(None, One(), b2 if frame % 2 else b2d, False, 1.0, Cfg(), renderer=r,
                                  rand_vec=(b2 if frame % 2 else b2d)["rand_vec"].reshape(-1, 3))
            assert torch.equal(got["rgb"], want["rgb"]) and torch.equal(got["acc"], want["acc"]), frame
        r._peer_image[1].close()
    dist.barrier()
    if rank == 0:
        print("DIST_OK", a.backend, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
