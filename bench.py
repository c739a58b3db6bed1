#!/usr/bin/env python
"""bench.py - UC-NeRF forward-render hot path on B200: ray-samples/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload eval_800x600_waymo_gin|target_1024spp]
    python bench.py --impl reference ...      # the CPU arm: oracle port of the reference path on host cores

A "step" renders one image of the workload per GPU (weak scaling: every rank renders its own frame of the same
size, then the ranks exchange the packed pixels with ONE NCCL all-gather, the multi-GPU eval design of
SURVEY.md section 8e).  value = ray-samples of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e` is the same metric through the host-buffer C-ABI entry (ucnerf_render_rays_host): pinned host rays in,
packed pixels back to the host, copies inside the timed region.  One JSON line is printed by rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eval_800x600_waymo_gin")
    ap.add_argument("--chunk-rays", type=int, default=0)
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="rays in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encode-runs", type=int, default=-1, help="A/B: cell-run reuse in the encode kernel (bit0 prop, bit1 NeRF); -1 = library default")
    ap.add_argument("--heads", action="store_true", help="also time the frame with the sky + brightness heads of the "
                    "shipped Waymo configuration (BASELINE.json configs[2]); reported under `with_heads`")
    ap.add_argument("--tc-debug", type=int, default=0, help="profiling experiment flags for the TC kernel (invalid results)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))), "measured"
    return 6650.0, 1400.0, "fallback"


def oracle_config(wl):
    from oracle import ucnerf_oracle as O  # CPU baseline only
    return O.HotPathConfig(num_prop_samples=wl.num_prop_samples, num_nerf_samples=wl.num_nerf_samples,
                           prop_grids=[O.GridSpec(d, log2_hashmap_size=wl.log2_hashmap_size) for d in wl.prop_desired],
                           nerf_grid=O.GridSpec(wl.nerf_desired, log2_hashmap_size=wl.log2_hashmap_size),
                           bottleneck_width=wl.bottleneck_width, net_width_viewdirs=wl.net_width_viewdirs)


def time_cpu_port(wl, sd, rays, n_sample, repeats=1, chunk=1024):
    """The oracle port of the reference path on the host cores (torch CPU, all threads).  The O(S^2) pairwise
    resampling temporaries of the reference bound the chunk size."""
    from oracle import ucnerf_oracle as O
    cfg = oracle_config(wl)
    torch.set_num_threads(os.cpu_count() or 1)
    n_sample = min(n_sample, rays["origins"].shape[0])
    if wl.num_prop_samples >= 512:
        chunk = min(chunk, 64)
    sub = {k: v[:n_sample] for k, v in rays.items()}
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for a in range(0, n_sample, chunk):
            O.model_forward(sd, cfg, {k: v[a:a + chunk] for k, v in sub.items()})
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_sample * wl.samples_per_ray / best, best, n_sample


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port; the reference is Python and does not travel)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from ucnerf_b200 import synthetic
    wl = synthetic.WORKLOADS[args.workload]
    sd = synthetic.synthetic_state_dict(wl, seed=0)
    rays = synthetic.pinhole_rays(wl.height, wl.width, seed=0)
    n_sample = args.cpu_sample_rays or (256 if wl.num_prop_samples >= 512 else 2048)
    for _ in range(max(args.warmup, 0) and 1):
        time_cpu_port(wl, sd, rays, min(n_sample, 256))
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        _, dt, n = time_cpu_port(wl, sd, rays, n_sample)
        tot += n
    el = time.perf_counter() - t0
    value = tot * wl.samples_per_ray / el
    sample = f"first {n_sample} rays of the frame per step, oracle port (torch CPU fp32), {os.cpu_count()} threads"
    line = {"impl": "reference", "metric": "ray_samples_per_sec", "value": value, "unit": "ray-samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_step": n_sample, "samples_per_ray": wl.samples_per_ray},
            "rays_per_sec": value / wl.samples_per_ray,
            "cpu_baseline": {"value": value, "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch.distributed as dist
    from ucnerf_b200 import _lib, synthetic
    from ucnerf_b200.render import PACKED_WIDTH

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        sys.exit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    wl = synthetic.WORKLOADS[args.workload]
    sd = synthetic.synthetic_state_dict(wl, seed=0)            # same weights on every rank (replicated model)
    r = synthetic.make_renderer(wl, sd, dev)
    if args.chunk_rays:
        r.set_option("chunk_rays", args.chunk_rays)
    if args.tc_debug:
        r.set_option("tc_debug", args.tc_debug)
    if args.encode_runs >= 0:
        r.set_option("encode_runs", args.encode_runs)
    rays_h = synthetic.pinhole_rays(wl.height, wl.width, seed=rank)   # every rank renders its own camera
    n = rays_h["origins"].shape[0]
    rays_d = {k: v.to(dev) for k, v in rays_h.items()}
    want = ("packed",)
    gathered = torch.empty((world * n, PACKED_WIDTH), device=dev) if world > 1 else None

    def step():
        out = r.render_rays(rays_d, 1.0, rays_d["rand_vec"], want)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out["packed"])
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    # ---- timed region: device-resident inputs -------------------------------------------------
    r.set_option("timing", 1)
    r.timing(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    fam = r.timing(reset=True)
    r.set_option("timing", 0)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # ---- end-to-end: host buffers through the C-ABI host entry --------------------------------
    pin = {k: v.reshape(-1).contiguous().pin_memory() if k in ("radii", "near", "far") else v.contiguous().pin_memory()
           for k, v in rays_h.items()}
    out_h = {"packed": torch.empty((n, PACKED_WIDTH), dtype=torch.float32).pin_memory()}
    for _ in range(max(1, min(args.warmup, 2))):
        r.render_rays_host(pin, 1.0, want=("packed",), out=out_h)
    sync_all()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        r.render_rays_host(pin, 1.0, want=("packed",), out=out_h)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_h["packed"].to(dev, non_blocking=True))
    e3.record()
    sync_all()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    # ---- end-to-end from camera parameters: rays generated on the GPU (ucnerf_render_camera_host) -----------
    cam = synthetic.pinhole_camera(wl.height, wl.width, seed=rank)
    for _ in range(max(1, min(args.warmup, 2))):
        r.render_camera(*cam, want=("packed",), host_out=out_h)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.render_camera(*cam, want=("packed",), host_out=out_h)
    ms_cam = (time.perf_counter() - t0) * 1e3     # the call synchronises: wall clock == device + copy time
    t = torch.tensor([ms_cam], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_cam = float(t.item())
    # ---- optional: the frame with the heads of scripts/train_waymo.sh (sky head kernel + brightness affine) --------
    heads_info = None
    if args.heads:
        from ucnerf_b200.render import SkyHead
        hsd = synthetic.synthetic_heads(seed=0)
        sky = SkyHead(hsd, device=dev)
        aff_sky = hsd["affine_sky"].to(dev)

        def step_heads():
            r.set_rgb_affine(hsd["affine"])
            out = r.render_rays(rays_d, 1.0, rays_d["rand_vec"], ("packed", f"weights_{r.num_levels - 1}"))
            srgb = sky.render(rays_d["origins"], rays_d["directions"], rays_d["far"], rays_d["cam_dirs"])
            opac = 1 - out[f"weights_{r.num_levels - 1}"].sum(-1, keepdim=True)      # models.py:L351-354
            out["packed"][:, 0:3] += opac * (srgb @ aff_sky[:3, :3].T + aff_sky[:3, 3])
            if world > 1:
                dist.all_gather_into_tensor(gathered, out["packed"])
            return out

        for _ in range(2):
            step_heads()
        sync_all()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for _ in range(args.steps):
            step_heads()
        e5.record()
        sync_all()
        r.set_rgb_affine(None)
        t = torch.tensor([e4.elapsed_time(e5)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_h = float(t.item()) / args.steps
        heads_info = {"ms_per_step": ms_h, "rays_per_sec": world * n / (ms_h * 1e-3),
                      "sky_head_ms": ms_h - ms / args.steps, "sky_samples_per_ray": 120,
                      "note": "fused foreground path + tensor-core sky head (8x256 MLP x 120 samples per ray) + both "
                              "brightness affines, models.py:L326-363; device-resident rays"}
    clocks = sampler.stop() if sampler else None
    checksum = float(out_h["packed"][:, :3].double().mean())

    if rank == 0:
        spr = wl.samples_per_ray
        total_samples = world * n * spr * args.steps
        value = total_samples / (ms * 1e-3)
        hbm_peak, tf_peak, how = measured_peaks()
        # roofline of the dominant kernel family (per launch, live CUDA-event durations from the timed region)
        alg_bytes = {"encode_prop": sum(wl.num_prop_samples * 768 * wl.grid_levels(d) for d in wl.prop_desired),
                     "encode_nerf": wl.num_nerf_samples * 768 * wl.grid_levels(wl.nerf_desired)}
        fam_ms = {k: v[0] for k, v in fam.items()}
        dom = max(fam_ms, key=fam_ms.get)
        shares = {k: round(v / max(sum(fam_ms.values()), 1e-9), 4) for k, v in fam_ms.items()}
        roofline = None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp):   # per-launch DRAM bytes of that kernel from the committed ncu --set full capture
            tj = json.load(open(tp))
            if tj.get("workload") == wl.name and dom in tj and not args.chunk_rays:
                traffic = tj[dom]["dram_bytes_read"] + tj[dom]["dram_bytes_write"]
        if dom in alg_bytes:
            bytes_total = alg_bytes[dom] * n * args.steps       # all launches of that family on this rank
            achieved = bytes_total / (fam_ms[dom] * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": f"sample_encode_kernel ({dom})", "achieved": achieved,
                        "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                        "note": "achieved = ALGORITHMIC gather bytes (768*L B per ray-sample) / time; the gathers are "
                                "served mostly from L1/L2 (the 101 MB proposal table is L2-resident, neighbouring "
                                "rays share cells), so measured DRAM `traffic` per launch is far below the algorithmic "
                                "bytes and frac can exceed 1; the kernel is bound by L1 wavefronts / issue, see "
                                "profiles/r1_summary.md",
                        "peak_source": how, "launches": fam[dom][1], "avg_launch_ms": fam_ms[dom] / max(fam[dom][1], 1),
                        "algorithmic_bytes_per_launch": bytes_total / max(fam[dom][1], 1)}
        else:
            flops = wl.mlp_flops_per_ray() * n * args.steps
            achieved = flops / (fam_ms[dom] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf_peak, "traffic": traffic, "peak_source": how, "launches": fam[dom][1],
                        "avg_launch_ms": fam_ms[dom] / max(fam[dom][1], 1)}
        # all hash-gather kernels together (the BASELINE.md convention: whole-frame gather bytes / frame time)
        gather_gbs = wl.gather_bytes_per_ray() * n * args.steps / ((fam_ms["encode_prop"] + fam_ms["encode_nerf"]) * 1e-3) / 1e9
        line = {
            "metric": "ray_samples_per_sec", "value": value, "unit": "ray-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_gpu_per_step": n, "image": f"{wl.width}x{wl.height}",
                       "samples_per_ray": spr, "prop_samples": wl.num_prop_samples, "nerf_samples": wl.num_nerf_samples,
                       "grid_levels": [wl.grid_levels(d) for d in wl.prop_desired] + [wl.grid_levels(wl.nerf_desired)],
                       "log2_hashmap_size": wl.log2_hashmap_size, "parallelism": f"ray-tile x{world}, replicated model",
                       "collective": "1 all_gather of packed [rays,12] per step" if world > 1 else "none",
                       "l2": "working set (330 MB hash tables + >0.5 GB per-chunk workspace) exceeds the 126 MB L2; "
                             "no explicit flush"},
            "rays_per_sec": value / spr,
            "e2e": {"value": total_samples / (ms_e2e * 1e-3), "unit": "ray-samples/s",
                    "h2d_bytes_per_step": int(n * 18 * 4), "d2h_bytes_per_step": int(n * PACKED_WIDTH * 4),
                    "ms_per_step": ms_e2e / args.steps, "api": "ucnerf_render_rays_host (C ABI, pinned host buffers)"},
            "e2e_camera": {"value": total_samples / (ms_cam * 1e-3), "unit": "ray-samples/s", "h2d_bytes_per_step": 200,
                           "d2h_bytes_per_step": int(n * PACKED_WIDTH * 4), "ms_per_step": ms_cam / args.steps,
                           "api": "ucnerf_render_camera_host (rays generated on the GPU from pose + intrinsics; "
                                  "replaces the loader's numpy pixels_to_rays)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernel_time_shares": shares,
            "kernel_ms_per_step": {k: v / args.steps for k, v in fam_ms.items()},
            "hash_gather_all_levels": {"achieved": gather_gbs, "unit": "GB/s", "frac": gather_gbs / hbm_peak,
                                       "bytes_per_ray": wl.gather_bytes_per_ray()},
            "clocks": clocks, "checksum_mean_rgb": checksum,
        }
        if heads_info:
            line["with_heads"] = heads_info
        if world == 1 and not args.no_cpu_baseline:
            ns = args.cpu_sample_rays or (128 if wl.num_prop_samples >= 512 else 8192)
            v, dt, ns = time_cpu_port(wl, sd, rays_h, ns)
            line["cpu_baseline"] = {"value": v, "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"first {ns} rays of the same frame, oracle port (torch CPU fp32, all "
                                              f"threads), {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
