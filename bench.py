#!/usr/bin/env python
"""bench.py - UC-NeRF forward-render hot path on B200: ray-samples/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload eval_800x600_waymo_gin|target_1024spp] [--extras LIST]
    python bench.py --impl reference ...      # the CPU arm: the reference's own Python on the host cores

A "step" renders one image of the workload per GPU (weak scaling: every rank renders its own frame-sized tile of a
world x larger image, then the ranks exchange the packed pixels with ONE NCCL all-gather - the tile shard + gather code
of ucnerf_b200.render that `render_image` uses).  value = ray-samples of all ranks /
max-over-ranks device time, inputs resident in HBM.  `e2e` is the same metric through the host-buffer C-ABI entry
(ucnerf_render_rays_host): pinned host rays in, packed pixels back to the host, copies inside the timed region.

One JSON line is printed by rank 0.  Besides the contract keys it carries (each with its own clocks sample):
  parity                    GPU pixels of the bench frame vs the CPU reference leg (and vs the reference on the same GPU)
  roofline / rooflines      per kernel family: algorithmic fraction, measured-DRAM fraction, the ncu limiter
  e2e_camera, e2e_render_image   rays generated on the GPU; the `render_image()` surface itself, host batch in
  target_1024spp            north_star target config (512 + 512 samples per ray)
  with_heads                BASELINE config 3: one full-resolution frame with the sky + brightness heads
  strong_scaling_full_res   BASELINE config 4: that frame row-tiled over the N GPUs
  train_step                BASELINE config 5: one optimisation step, 8,192 rays per GPU
  gpu_reference             the reference's own Python + its own CUDA kernel (sm_100a build) on this same GPU
  cpu_baseline              the reference's own Python + C/OpenMP grid kernel on the host cores (bounded sample)."""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

EXTRAS_1 = ("parity", "e2e_camera", "e2e_render_image", "target_1024spp", "with_heads", "train_step", "gpu_reference",
            "cpu_baseline")          # (the default N = 1 run takes about 40 s)
EXTRAS_N = ("strong_scaling_full_res", "train_step")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eval_800x600_waymo_gin")
    ap.add_argument("--extras", default="default", help="'default', 'none', 'all' or a comma list of: "
                    + ",".join(sorted(set(EXTRAS_1 + EXTRAS_N))))
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="N > 1: fused NVLink tile exchange (default) "
                    "or render + NCCL all-gather")
    ap.add_argument("--cameras", default="same", choices=["same", "distinct"], help="N > 1: every rank renders the same "
                    "camera (identical work per rank, default) or rank r renders camera r")
    ap.add_argument("--chunk-rays", type=int, default=0)
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="rays in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encode-runs", type=int, default=-1, help="A/B: cell-run reuse in the encode kernel (bit0 prop, bit1 NeRF); -1 = library default")
    ap.add_argument("--heads", action="store_true", help="alias of --extras ...,with_heads")
    ap.add_argument("--option", action="append", default=[], help="key=int library option (A/B experiments)")
    ap.add_argument("--ray-tile-width", default="image", help="'image' (default): tell the library that the ray batch is a row-major "
                    "image of the workload's width (option ray_tile_width: 4x8-pixel patches per warp, bit-identical results); 0 = off")
    ap.add_argument("--tc-debug", type=int, default=0, help="profiling experiment flags for the TC kernel (invalid results)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  Read through NVML
    inside this process (what nvidia-smi itself queries) every 50 ms: spawning an nvidia-smi process five times a second
    next to the timed steps occasionally cost the steps tens of milliseconds on a fresh box.  NVML is initialised and
    queried once BEFORE the region (`prime`); nvidia-smi remains the fallback where pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    _nvml = None        # (module, {gpu index: handle}) once primed; False if unavailable

    @classmethod
    def prime(cls, gpu_index=0):
        """Initialise NVML and resolve the device handle (by PCI bus id of the CUDA device, so CUDA_VISIBLE_DEVICES does not
        matter); one query, so that the first-call cost is paid outside every timed region."""
        if cls._nvml is False:
            return None
        try:
            if cls._nvml is None:
                import pynvml
                pynvml.nvmlInit()
                cls._nvml = (pynvml, {})
            nv, handles = cls._nvml
            if gpu_index not in handles:
                try:
                    pr = torch.cuda.get_device_properties(gpu_index)
                    bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                    handles[gpu_index] = nv.nvmlDeviceGetHandleByPciBusId(bus.encode())
                except Exception:
                    handles[gpu_index] = nv.nvmlDeviceGetHandleByIndex(gpu_index)
                h = handles[gpu_index]          # first calls of every query used later (they are the slow ones)
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            return handles[gpu_index]
        except Exception:
            cls._nvml = False
            return None

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []          # (sm MHz, max sm MHz, set of reasons)
        self.handle = self.prime(gpu_index)
        self._stop_evt = threading.Event()
        self.first_done = threading.Event()

    def _sample_nvml(self):
        nv = self._nvml[0]
        sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.rows.append((float(sm), float(mx), {n for n, b in self.REASONS if mask & b}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                              str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [c.strip() for c in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.rows.append((float(r[1]), float(r[2]), {n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")}))

    def run(self):
        first = True
        while not self._stop_evt.is_set():
            try:
                if self.handle is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            if first:       # taken before the caller's timed region starts (Clocks.__enter__ waits for it); dropped in stop()
                first = False
                self.first_done.set()
            self._stop_evt.wait(0.05 if self.handle is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        if len(self.rows) > 1:
            self.rows = self.rows[1:]       # the priming sample was taken on an idle GPU
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.handle is not None else "nvidia-smi"}


class Clocks:
    """with Clocks(rank, gpu) as c: ...  -> c.result (None on ranks other than 0)."""

    def __init__(self, rank, gpu):
        self.s = ClockSampler(gpu) if rank == 0 else None
        self.result = None

    def __enter__(self):
        # everything that can stall the launching thread happens here, before the caller synchronises and starts its clock:
        # thread start, the sampler's first driver query, a full garbage collection (then none until __exit__)
        if self.s:
            self.s.start()
            self.s.first_done.wait(2.0)
        gc.collect()
        gc.disable()
        return self

    def __exit__(self, *a):
        gc.enable()
        if self.s:
            self.result = self.s.stop()


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own Python on the host cores
# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU path for the hot path: its own, unmodified Python (internal/models.py `Model.forward`, render.py,
    stepfun.py, coord.py, gridencoder/grid.py - from /root/reference or the staged copy baseline/_ref) through
    oracle/ref_shim.py, with the C + OpenMP restatement of its CUDA-only grid kernel behind `_gridencoder`
    (oracle/grid_cpu.c; the reference has no CPU kernel).  kind = "reference".  Where the reference tree is not staged it
    falls back to the oracle port (kind = "port")."""

    def __init__(self, wl, sd):
        torch.set_num_threads(os.cpu_count() or 1)
        self.wl, self.sd = wl, sd
        self.cfg = oracle_config(wl)
        from oracle import ref_shim
        self.kind = "port"
        self.model = None
        if ref_shim.available():
            try:
                ref_shim.load_reference()
                ref_shim.use_grid_backend("oracle_c")
                self.model, _ = ref_shim.build_reference_model(self.cfg, {k: v.cpu() for k, v in sd.items()})
                self.kind = "reference"
            except Exception as e:  # a broken staging must not kill the bench: report the port instead
                print(f"[bench] reference tree not usable on CPU ({e}); timing the oracle port", file=sys.stderr)
        self.chunk = 64 if wl.num_prop_samples >= 512 else 2048      # the O(S^2) pairwise temporaries bound it

    def describe(self, n, dt=None):
        what = ("reference Python (internal/models.py Model.forward, unmodified) + C/OpenMP restatement of kernel_grid"
                if self.kind == "reference" else "oracle port (torch CPU fp32)")
        s = f"first {n} rays of the frame, {what}, {os.cpu_count()} threads, chunks of {self.chunk} rays"
        return s + (f", {dt:.1f} s" if dt is not None else "")

    @torch.no_grad()
    def render(self, rays, n):
        """-> dict(rgb, acc, depth_raw-or-None) for rays[:n], seconds."""
        from oracle import ref_shim, ucnerf_oracle as O
        n = min(n, rays["origins"].shape[0])
        outs = {"rgb": [], "acc": [], "depth": []}
        t0 = time.perf_counter()
        for a in range(0, n, self.chunk):
            sub = {k: v[a:min(a + self.chunk, n)] for k, v in rays.items()}
            if self.model is not None:
                ref_shim.use_grid_backend("oracle_c")
                with ref_shim.inject_rand_vec(sub["rand_vec"]):
                    rr, _ = self.model(False, {k: v for k, v in sub.items() if k != "rand_vec"}, train_frac=1.0,
                                       compute_extras=True, zero_glo=True)
            else:
                rr, _ = O.model_forward(self.sd, self.cfg, sub)
            for k in outs:
                outs[k].append(rr[-1][k])
        dt = time.perf_counter() - t0
        return {k: torch.cat(v) for k, v in outs.items()}, dt, n


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from ucnerf_b200 import synthetic
    wl = synthetic.WORKLOADS[args.workload]
    sd = synthetic.synthetic_state_dict(wl, seed=0)
    rays = synthetic.pinhole_rays(wl.height, wl.width, seed=0)
    cpu = CpuReference(wl, sd)
    n_sample = args.cpu_sample_rays or (128 if wl.num_prop_samples >= 512 else 16384)
    if args.warmup > 0:
        cpu.render(rays, min(n_sample, cpu.chunk))
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        _, dt, n = cpu.render(rays, n_sample)
        tot += n
    el = time.perf_counter() - t0
    value = tot * wl.samples_per_ray / el
    line = {"impl": "reference", "metric": "ray_samples_per_sec", "value": value, "unit": "ray-samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_step": n_sample, "samples_per_ray": wl.samples_per_ray},
            "rays_per_sec": value / wl.samples_per_ray,
            "cpu_baseline": {"value": value, "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": cpu.kind,
                             "sample": cpu.describe(n_sample) + " per step"},
            "e2e": {"value": value, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def sync_all(cx):
    torch.cuda.synchronize()
    if cx.world > 1:
        cx.dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(cx, ms):
    if cx.world == 1:
        return float(ms)
    t = torch.tensor([ms], device=cx.dev, dtype=torch.float64)
    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    return float(t.item())


def timed(cx, fn, steps, warmup):
    """warm-up, barrier + sync, CUDA events around `steps` calls on the current stream, barrier + sync, max over ranks.
    Returns ms for all steps (device time; the wall clock is taken too and the larger of the two is used, so a call that
    blocks on the host - D2H copies - is not under-reported)."""
    for _ in range(warmup):
        fn()
    sync_all(cx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), wall if cx.host_blocking else 0.0)
    sync_all(cx)
    return max_over_ranks(cx, ms)


def load_traffic():
    """Per-launch DRAM bytes and limiter of each kernel family from the committed ncu --set full capture of this build
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py from gpurun_out/*.ncu-rep; falls back to the r1 file)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            d["_file"] = "profiles/" + name
            return d
    return {}


def rooflines(wl, n, steps, fam, chunk_rays_opt):
    """One entry per kernel family: what the spec's recipe gives (algorithmic bytes or flops / live event time / measured
    peak), what DRAM really moved (ncu capture of this build), and the unit that limits the kernel."""
    hbm_peak, tf_peak, how = measured_peaks()
    tj = load_traffic()
    usable = tj.get("workload") == wl.name and not chunk_rays_opt
    alg_bytes = {"encode_prop": sum(wl.num_prop_samples * 768 * wl.grid_levels(d) for d in wl.prop_desired),
                 "encode_nerf": wl.num_nerf_samples * 768 * wl.grid_levels(wl.nerf_desired)}
    lc = 4 * wl.grid_levels(wl.nerf_desired)
    bw, w, nd = wl.bottleneck_width, wl.net_width_viewdirs, 27
    alg_flops = {"color_mlp": wl.num_nerf_samples * 2 * (64 * bw + (bw + nd) * w + (w + bw + nd) * w + w * 3)}
    kernels = {"encode_prop": f"sample_encode_kernel<{wl.grid_levels(wl.prop_desired[0])},prop>",
               "encode_nerf": f"sample_encode_kernel<{wl.grid_levels(wl.nerf_desired)},nerf>",
               "color_mlp": "color_mlp_tc_kernel", "resample": "resample_kernel", "composite": "composite_kernel"}
    out = {}
    for k, (ms, launches) in fam.items():
        if launches == 0 or ms <= 0:
            continue
        e = {"kernel": kernels.get(k, k), "launches": launches, "avg_launch_ms": ms / launches, "ms_per_step": ms / steps}
        t = tj.get(k) if usable else None
        if k in alg_bytes:
            total = alg_bytes[k] * n * steps
            ach = total / (ms * 1e-3) / 1e9
            e.update({"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "roofline_kind": "hbm",
                      "algorithmic_bytes_per_launch": total / launches})
        elif k in alg_flops:
            total = alg_flops[k] * n * steps
            ach = total / (ms * 1e-3) / 1e12
            e.update({"achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "roofline_kind": "tensor",
                      "algorithmic_flops_per_launch": total / launches,
                      "note": "reference arithmetic (fp32 MACs x 2) / time / measured dense-bf16 peak; the kernel issues 3 "
                              "FP16 tensor passes per product to keep fp32 accuracy, after folding the bottleneck layer"})
        if t and k not in ("encode_prop", "encode_nerf", "color_mlp"):
            # two launches of different size per chunk (one per level): quote the limiter, not a per-launch traffic figure
            e["traffic"] = None
            e["bound"] = t.get("bound", "unknown")
            e["limiter_pct"] = t.get("limiter_pct")
        elif t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
            # per-launch DRAM bytes of the captured launch, scaled to this run's average launch size
            scale = (n * steps / launches) / t.get("rays_per_launch", tj.get("chunk_rays", 131072))
            e["traffic"] = traffic * scale
            e["dram_frac"] = traffic * scale / (e["avg_launch_ms"] * 1e-3) / 1e9 / hbm_peak
            e["bound"] = t.get("bound", "unknown")
            if "limiter_pct" in t:
                e["limiter_pct"] = t["limiter_pct"]
        else:
            e["traffic"] = None
            e["bound"] = {"encode_prop": "l1_lsu_wavefronts+issue", "encode_nerf": "l1_lsu_wavefronts",
                          "color_mlp": "tensor", "resample": "issue", "composite": "issue"}.get(k, "unknown")
        e["peak_source"] = how
        e["traffic_source"] = tj.get("_file") if t else None
        out[k] = e
    return out


def run_ours(args):
    import torch.distributed as dist
    from ucnerf_b200 import _lib, synthetic
    from ucnerf_b200 import render as R
    from ucnerf_b200.render import PACKED_WIDTH

    cx = Ctx()
    cx.dist = dist
    cx.world = world = int(os.environ.get("WORLD_SIZE", 1))
    cx.rank = rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cx.host_blocking = False
    if not torch.cuda.is_available():
        sys.exit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cx.dev = dev = torch.device(f"cuda:{local}")
    if rank == 0:
        ClockSampler.prime(local)        # NVML initialised and queried once, well before any timed region
    t_start = time.perf_counter()
    if args.extras == "default":
        extras = set(EXTRAS_1 if world == 1 else EXTRAS_N)
    elif args.extras == "all":
        extras = set(EXTRAS_1 + EXTRAS_N)
    elif args.extras == "none":
        extras = set()
    else:
        extras = set(x for x in args.extras.split(",") if x)
    if args.heads:
        extras.add("with_heads")
    if args.no_cpu_baseline:
        extras.discard("cpu_baseline")
    if world > 1:
        extras -= {"cpu_baseline", "gpu_reference", "parity"}

    wl = synthetic.WORKLOADS[args.workload]
    sd = synthetic.synthetic_state_dict(wl, seed=0)            # same weights on every rank (replicated model)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    r = synthetic.make_renderer(wl, sd_dev, dev)
    if args.chunk_rays:
        r.set_option("chunk_rays", args.chunk_rays)
    if args.tc_debug:
        r.set_option("tc_debug", args.tc_debug)
    if args.encode_runs >= 0:
        r.set_option("encode_runs", args.encode_runs)
    tile_w = wl.width if args.ray_tile_width == "image" else int(args.ray_tile_width)
    if tile_w and wl.width % 4 == 0:
        r.set_option("ray_tile_width", tile_w)     # the batch below is that frame, row-major (as render_image passes it)
    for kv in args.option:
        k, v = kv.split("=")
        r.set_option(k, int(v))
    # every rank renders one frame of the workload: the SAME camera by default, so that the work per rank is identical and
    # the N-GPU number measures the system (exchange, clocks, launch path), not the scene variance between camera poses
    # (--cameras distinct: rank r renders camera r; frames then differ by a few per cent and max-over-ranks pays for it)
    cam_seed = rank if args.cameras == "distinct" else 0
    rays_h = synthetic.pinhole_rays(wl.height, wl.width, seed=cam_seed)
    n = rays_h["origins"].shape[0]
    rays_d = {k: v.to(dev) for k, v in rays_h.items()}
    # weak scaling = one image of world * n rays, rank `rank` owns tile [start, stop) (render.shard_bounds), the tiles
    # are exchanged by render.gather_tiles - the code path of render_image
    per, start, stop = R.shard_bounds(world * n, world, rank)
    assert per == n and stop - start == n

    peer = None
    if world > 1 and args.gather == "peer":
        from ucnerf_b200.peer import PeerImage
        try:
            peer = PeerImage(world * n, device=dev)          # collective; every rank falls back together
        except _lib.UcnerfError as e:
            if rank == 0:
                print(f"[bench] peer exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)

    def step_nccl():
        out = r.render_rays(rays_d, 1.0, rays_d["rand_vec"], ("packed",))
        img = R.gather_tiles(out["packed"], world, world * n)
        return out, img

    def step_peer():    # the tile exchange fused into the compositing kernel (NVLink peer stores) + one 4-byte all-reduce
        img, _ = peer.render(r, rays_d, 1.0, rays_d["rand_vec"], start)
        return {"packed": img[start:stop]}, img

    step = step_peer if peer is not None else step_nccl

    # ---- timed region: device-resident inputs -------------------------------------------------
    for _ in range(args.warmup):
        step()
    sync_all(cx)
    r.set_option("timing", 1)
    r.timing(reset=True)
    launches0 = _lib.launch_count()
    with Clocks(rank, local) as ck:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all(cx)
        e0.record()
        for _ in range(args.steps):
            out_dev, _ = step()
        e1.record()
        sync_all(cx)
    ms = max_over_ranks(cx, e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    fam = r.timing(reset=True)
    r.set_option("timing", 0)
    clocks = ck.result
    packed_dev = out_dev["packed"].clone()
    # N > 1: the same step with the other exchange, for the record (not the headline)
    other_ms = None
    if world > 1 and peer is not None:
        other_ms = timed(cx, step_nccl, args.steps, 2) / args.steps

    # ---- end-to-end: host buffers through the C-ABI host entry --------------------------------
    pin = {k: v.reshape(-1).contiguous().pin_memory() if k in ("radii", "near", "far") else v.contiguous().pin_memory()
           for k, v in rays_h.items()}
    out_h = {"packed": torch.empty((n, PACKED_WIDTH), dtype=torch.float32).pin_memory()}
    cx.host_blocking = True
    if world == 1:
        def e2e_step():
            r.render_rays_host(pin, 1.0, want=("packed",), out=out_h)
    else:
        pin2 = {k: (v if v.dim() == 2 else v.reshape(-1, 1)) for k, v in pin.items()}
        dev_in = {k: torch.empty_like(v, device=dev) for k, v in pin2.items()}

        def e2e_step():     # host rays in -> device render + tile exchange -> this rank's rows of the image back
            if peer is None:
                for k in dev_in:
                    dev_in[k].copy_(pin2[k], non_blocking=True)
            if peer is not None:
                # H2D of chunk c + 1 and D2H of this rank's rows of chunk c - 1 under chunk c's kernels; the tiles are
                # written to every rank's image by the compositing kernel; the all-reduce closes the frame
                peer.render_host(r, pin, 1.0, start, want=("packed",), out=out_h)
            else:
                o = r.render_rays(dev_in, 1.0, dev_in["rand_vec"], ("packed",))
                img = R.gather_tiles(o["packed"], world, world * n)
                out_h["packed"].copy_(img[start:stop], non_blocking=True)
            torch.cuda.current_stream().synchronize()
    with Clocks(rank, local) as ck:
        ms_e2e = timed(cx, e2e_step, args.steps, max(1, min(args.warmup, 2)))
    e2e = {"value": world * n * wl.samples_per_ray * args.steps / (ms_e2e * 1e-3), "unit": "ray-samples/s",
           "h2d_bytes_per_step": int(n * 18 * 4), "d2h_bytes_per_step": int(n * PACKED_WIDTH * 4),
           "ms_per_step": ms_e2e / args.steps, "clocks": ck.result,
           "api": "ucnerf_render_rays_host (C ABI, pinned host buffers; H2D of chunk c+1 and D2H of chunk c-1 overlap chunk c)"
                  if world == 1 else ("peer.PeerImage.render_host: ucnerf_render_rays_host with the tile exchange fused into the compositing "
                                     "kernel; own rows back to pinned host memory chunk by chunk" if peer is not None else
                                     "pinned host rays -> render_rays -> render.gather_tiles (1 all-gather) -> own tile D2H")}
    checksum = float(out_h["packed"][:, :3].double().mean())

    line_extra = {}
    # ---- end-to-end from camera parameters: rays generated on the GPU (ucnerf_render_camera_host) -----------
    if "e2e_camera" in extras:
        cam = synthetic.pinhole_camera(wl.height, wl.width, seed=cam_seed)
        with Clocks(rank, local) as ck:
            ms_cam = timed(cx, lambda: r.render_camera(*cam, want=("packed",), host_out=out_h), args.steps, 2)
        line_extra["e2e_camera"] = {
            "value": world * n * wl.samples_per_ray * args.steps / (ms_cam * 1e-3), "unit": "ray-samples/s",
            "h2d_bytes_per_step": 200, "d2h_bytes_per_step": int(n * PACKED_WIDTH * 4), "ms_per_step": ms_cam / args.steps,
            "clocks": ck.result, "api": "ucnerf_render_camera_host (rays generated on the GPU from pose + intrinsics; "
                                        "replaces the loader's numpy pixels_to_rays)"}
    # ---- the surface north_star names: render_image(), host batch in, [H,W,...] dict out, pixels read back --------
    if "e2e_render_image" in extras:
        b2d = {k: (v if v.dim() == 2 else v.reshape(-1, 1)).reshape(wl.height, wl.width, -1) for k, v in pin.items()}
        rv = b2d.pop("rand_vec").reshape(-1, 3)
        conf = types.SimpleNamespace(vis_num_rays=16, render_chunk_size=15000, model_sky=False, brightness_correction=False)
        host_img = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
                    (("rgb", (wl.height, wl.width, 3)), ("depth", (wl.height, wl.width)), ("acc", (wl.height, wl.width)))}

        def ri_step():
            img = R.render_image(None, None, b2d, False, 1.0, conf, verbose=False, renderer=r, rand_vec=rv)
            for k, v in host_img.items():
                v.copy_(img[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return img
        with Clocks(rank, local) as ck:
            ms_ri = timed(cx, ri_step, args.steps, 2)
        img = ri_step()
        line_extra["e2e_render_image"] = {
            "value": n * wl.samples_per_ray * args.steps / (ms_ri * 1e-3), "unit": "ray-samples/s",
            "ms_per_step": ms_ri / args.steps, "h2d_bytes_per_step": int(n * 18 * 4), "d2h_bytes_per_step": int(n * 5 * 4),
            "keys": sorted(img.keys()), "clocks": ck.result,
            "api": "ucnerf_b200.render.render_image(model=None, renderer=..., batch of pinned [H,W,.] host tensors) -> dict of "
                   "[H,W,.] device tensors incl. weights + ray bundles; rgb/depth/acc copied to pinned host memory"}
        del img
    cx.host_blocking = False

    # ---- north_star target: 1024 samples per ray -----------------------------------------------------------------
    if "target_1024spp" in extras and wl.name != "target_1024spp":
        wt = synthetic.WORKLOADS["target_1024spp"]
        rt = synthetic.make_renderer(wt, sd_dev, dev)
        rays_t = {k: v.to(dev) for k, v in synthetic.pinhole_rays(wt.height, wt.width, seed=rank).items()}
        nt = rays_t["origins"].shape[0]
        with Clocks(rank, local) as ck:
            ms_t = timed(cx, lambda: rt.render_rays(rays_t, 1.0, rays_t["rand_vec"], ("packed",)), 3, 2)
        line_extra["target_1024spp"] = {"value": nt * wt.samples_per_ray * 3 / (ms_t * 1e-3), "unit": "ray-samples/s",
                                        "rays_per_sec": nt * 3 / (ms_t * 1e-3), "ms_per_step": ms_t / 3, "rays": nt,
                                        "samples_per_ray": wt.samples_per_ray, "target": 1e8, "clocks": ck.result}
        rt.close()
        del rt, rays_t
    # ---- BASELINE configs 3 / 4: one full-resolution frame with the heads, row-tiled over the ranks -----------------
    if extras & {"with_heads", "strong_scaling_full_res"}:
        info = full_res_with_heads(cx, r, synthetic, R, local)
        if "with_heads" in extras and world == 1:
            line_extra["with_heads"] = info
        if "strong_scaling_full_res" in extras or world > 1:
            line_extra["strong_scaling_full_res"] = dict(info, scaling="strong")
    # ---- BASELINE config 5: one optimisation step ----------------------------------------------------------------------
    if "train_step" in extras:
        try:
            line_extra["train_step"] = train_step(cx, local)
        except Exception as e:  # the training extra must never take the headline line down
            line_extra["train_step"] = {"error": repr(e)[:300]}
    # ---- the reference itself on this GPU ---------------------------------------------------------------------------------
    if "gpu_reference" in extras and world == 1:
        try:
            line_extra["gpu_reference"] = gpu_reference(cx, wl, sd_dev, rays_d, packed_dev, ms / args.steps, local)
        except Exception as e:
            line_extra["gpu_reference"] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        spr = wl.samples_per_ray
        total_samples = world * n * spr * args.steps
        value = total_samples / (ms * 1e-3)
        hbm_peak, tf_peak, how = measured_peaks()
        rl = rooflines(wl, n, args.steps, fam, args.chunk_rays)
        fam_ms = {k: v[0] for k, v in fam.items()}
        dom = max(fam_ms, key=fam_ms.get)
        shares = {k: round(v / max(sum(fam_ms.values()), 1e-9), 4) for k, v in fam_ms.items()}
        roofline = dict(rl.get(dom, {}))
        roofline["family"] = dom
        roofline["note"] = ("frac = ALGORITHMIC gather bytes (768*L B per ray-sample, SURVEY 8d) / live event time / HBM copy "
                            "peak. It is NOT a DRAM statement: the gathers are served by L1/L2 (101 MB proposal table is "
                            "L2-resident, neighbouring rays share cells), so frac can exceed 1; `dram_frac` is measured DRAM "
                            "bytes / time / peak and `bound` names the unit ncu shows saturated.")
        gather_ms = fam_ms.get("encode_prop", 0) + fam_ms.get("encode_nerf", 0)
        gather_gbs = wl.gather_bytes_per_ray() * n * args.steps / (gather_ms * 1e-3) / 1e9
        line = {
            "metric": "ray_samples_per_sec", "value": value, "unit": "ray-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_gpu_per_step": n, "image": f"{wl.width}x{wl.height}",
                       "samples_per_ray": spr, "prop_samples": wl.num_prop_samples, "nerf_samples": wl.num_nerf_samples,
                       "grid_levels": [wl.grid_levels(d) for d in wl.prop_desired] + [wl.grid_levels(wl.nerf_desired)],
                       "log2_hashmap_size": wl.log2_hashmap_size, "parallelism": f"ray-tile x{world}, replicated model",
                       "cameras": args.cameras, "ray_tile_width": tile_w if wl.width % 4 == 0 else 0,
                       "collective": ("none" if world == 1 else
                                      "fused: compositing kernel stores the packed tiles into every rank's image over NVLink "
                                      "peer memory (peer.PeerImage) + one 4-byte all_reduce per step" if peer is not None else
                                      "render.gather_tiles: 1 all_gather of packed [rays,12] per step"),
                       "l2": "working set (330 MB hash tables + >0.5 GB per-chunk workspace) exceeds the 126 MB L2; "
                             "no explicit flush"},
            "rays_per_sec": value / spr,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "rooflines": rl,
            "kernel_time_shares": shares,
            "kernel_ms_per_step": {k: v / args.steps for k, v in fam_ms.items()},
            "hash_gather_all_levels": {"achieved": gather_gbs, "unit": "GB/s", "frac": gather_gbs / hbm_peak,
                                       "bytes_per_ray": wl.gather_bytes_per_ray(),
                                       "note": "algorithmic bytes of both gather kernels / their summed time (BASELINE.md section 4)"},
            "clocks": clocks, "checksum_mean_rgb": checksum,
        }
        if other_ms is not None:
            line["nccl_all_gather_variant"] = {"ms_per_step": other_ms, "value": world * n * spr / (other_ms * 1e-3),
                                               "note": "same step with render + render.gather_tiles (NCCL all-gather of 48 B/ray)"}
        line.update(line_extra)
        if world == 1 and (extras & {"cpu_baseline", "parity"}):
            cpu = CpuReference(wl, sd)
            _, dt0, n0 = cpu.render(rays_h, cpu.chunk)   # warm-up (thread pools, first touch) and speed probe
            # bounded sample: about 15 s of CPU work on this box's cores (whole chunks, at most the frame)
            ns = args.cpu_sample_rays or max(cpu.chunk, min(n, int(15.0 / max(dt0 / n0, 1e-9)) // cpu.chunk * cpu.chunk))
            ref_out, dt, ns = cpu.render(rays_h, ns)
            v = ns * spr / dt
            line["cpu_baseline"] = {"value": v, "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": cpu.kind,
                                    "sample": cpu.describe(ns, dt)}
            # parity of the very frame the number is quoted on: GPU pixels vs the CPU reference leg's pixels
            g = packed_dev[:ns].cpu()
            par = line.setdefault("parity", {})
            par["vs_cpu_" + cpu.kind] = {
                "rays": ns, "rgb_linf": float((g[:, 0:3] - ref_out["rgb"]).abs().max()),
                "acc_linf": float((g[:, 4] - ref_out["acc"]).abs().max()),
                "depth_linf_clear_rays": _depth_err(g[:, 3], ref_out["depth"], ref_out["acc"]),
                "tolerance": 1e-4}
        if "gpu_reference" in line and "parity" in line["gpu_reference"]:
            line.setdefault("parity", {})["vs_gpu_reference"] = line["gpu_reference"].pop("parity")
        line["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(line), flush=True)
    if world > 1:
        if peer is not None:
            peer.close()
        dist.barrier()
        dist.destroy_process_group()


def _depth_err(d, d_ref, acc_ref):
    """depth is overridden to 300 where acc < 0.6 (render.py:L208,L213): compare on rays away from that threshold,
    relative to the depth range."""
    clear = (acc_ref - 0.6).abs() > 1e-3
    if not bool(clear.any()):
        return None
    return float(((d - d_ref).abs() / d_ref.abs().clamp_min(1.0))[clear].max())


def full_res_with_heads(cx, r, synthetic, R, local):
    """BASELINE.json configs[2] (N=1) / configs[3] (N>1): ONE 1920x1280 Waymo-sized frame with the sky head and both
    brightness affines of scripts/train_waymo.sh, image rows tiled over the ranks (strong scaling), one all-gather of the
    finished packed tiles (render.gather_tiles)."""
    from ucnerf_b200.render import SkyHead, generate_rays
    wf = synthetic.WORKLOADS["eval_1920x1280_waymo_gin"]
    H, W = wf.height, wf.width
    hsd = synthetic.synthetic_heads(seed=0)
    sky = SkyHead(hsd, device=cx.dev)
    aff_sky = hsd["affine_sky"].to(cx.dev)
    cam = synthetic.pinhole_camera(H, W, seed=0)
    rows_per = (H + cx.world - 1) // cx.world
    row0 = min(cx.rank * rows_per, H)
    nrows = min(rows_per, H - row0)
    nl = r.num_levels

    def frame():
        rays = generate_rays(*cam, rows=(row0, nrows), device=cx.dev)
        r.set_rgb_affine(hsd["affine"])
        out = r.render_rays(rays, 1.0, rays["rand_vec"], ("packed", f"weights_{nl - 1}"))
        r.set_rgb_affine(None)
        srgb = sky.render(rays["origins"], rays["directions"], rays["far"], rays["cam_dirs"])
        opac = 1 - out[f"weights_{nl - 1}"].sum(-1, keepdim=True)      # models.py:L351-354
        out["packed"][:, 0:3] += opac * (srgb @ aff_sky[:3, :3].T + aff_sky[:3, 3])
        tile = out["packed"]
        if nrows < rows_per:   # last tile padded to the common size
            tile = torch.cat([tile, tile.new_zeros(((rows_per - nrows) * W, tile.shape[1]))])
        return R.gather_tiles(tile, cx.world, H * W)

    def hot_only():
        rays = generate_rays(*cam, rows=(row0, nrows), device=cx.dev)
        return r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))

    steps = 2
    with Clocks(cx.rank, local) as ck:
        ms_f = timed(cx, frame, steps, 1) / steps
        ms_hot = timed(cx, hot_only, steps, 1) / steps
    sky.close()
    torch.cuda.empty_cache()
    return {"ms_per_frame": ms_f, "rays_per_sec": H * W / (ms_f * 1e-3), "image": f"{W}x{H}", "n_gpus": cx.world,
            "ray_samples_per_sec": H * W * wf.samples_per_ray / (ms_f * 1e-3),
            "hot_path_ms": ms_hot, "heads_ms": ms_f - ms_hot, "sky_samples_per_ray": 120, "five_cam_s": 5 * ms_f * 1e-3,
            "clocks": ck.result,
            "note": "rays generated on the GPU per row tile + fused foreground path + tensor-core sky head (8x256 MLP x 120 "
                    "samples per ray) + both brightness affines (models.py:L326-363) + 1 all-gather; 5-camera rig = 5 "
                    "such frames"}


def train_step(cx, local):
    """BASELINE.json configs[4]: one optimisation step (forward + backward + exchange + step), 8,192 rays per GPU, through
    bench_train.py's recipe; `with_sky_head` = the same step with `Config.model_sky` (scripts/train_waymo.sh)."""
    import bench_train as BT
    with Clocks(cx.rank, local) as ck:
        res = BT.run(cx.dev, cx.world, cx.rank, steps=8, warmup=3, rays=8192)
    res["clocks"] = ck.result
    with Clocks(cx.rank, local) as ck:
        sky = BT.run(cx.dev, cx.world, cx.rank, steps=4, warmup=2, rays=8192, sky=True)
    res["with_sky_head"] = {k: sky[k] for k in ("value", "ms_per_step", "phase_ms_per_step", "gradient_allreduce_bytes_per_step")}
    res["with_sky_head"]["clocks"] = ck.result
    return res


def gpu_reference(cx, wl, sd_dev, rays_d, packed_ours, ours_ms_per_frame, local, chunk=15000):
    """SURVEY 8(d) "GPU-side reference (kernel to beat)": the reference's own Python (internal/models.py `Model.forward`,
    unmodified, from the staged tree) on this same B200, CUDA events around Model.forward in 15,000-ray chunks
    (configs/waymo.gin render_chunk_size) over the bench frame, with (a) the reference's own gridencoder.cu compiled for
    sm_100a and (b) this package's drop-in `_gridencoder` kernels behind `import _gridencoder`."""
    from oracle import ref_shim
    if not ref_shim.available():
        return {"unavailable": "reference tree not staged (baseline/stage_ref.py)"}
    cfg = oracle_config(wl)
    ref_shim.load_reference()
    model, _ = ref_shim.build_reference_model(cfg, {k: v.cpu() for k, v in sd_dev.items()})
    model = model.to(cx.dev).eval()
    n = rays_d["origins"].shape[0]
    b = {k: (v if v.dim() == 2 else v.reshape(-1, 1)) for k, v in rays_d.items() if k != "rand_vec"}
    rv = rays_d["rand_vec"]
    out = {"chunk_rays": chunk, "rays": n}
    backends = [("reference_kernel", "ref_cuda")] if os.path.exists(ref_shim.REF_CUDA_SO) else []
    backends.append(("dropin_kernels", "dropin"))

    @torch.no_grad()
    def frame(limit=None, keep=None):
        for a in range(0, n if limit is None else min(n, limit), chunk):
            cb = {k: v[a:a + chunk] for k, v in b.items()}
            with ref_shim.inject_rand_vec(rv[a:a + chunk]):
                rr, _ = model(False, cb, train_frac=1.0, compute_extras=True, zero_glo=True)
            if keep is not None:
                keep.append(torch.cat([rr[-1]["rgb"], rr[-1]["depth"][:, None], rr[-1]["acc"][:, None]], -1))

    for name, kind in backends:
        ref_shim.use_grid_backend(kind)
        frame(limit=chunk)                                   # warm-up: one chunk
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        keep = []
        with Clocks(cx.rank, local) as ck:
            e0.record()
            frame(keep=keep)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"ms_per_frame": ms, "ray_samples_per_sec": n * wl.samples_per_ray / (ms * 1e-3),
                     "speedup_of_this_package": ms / ours_ms_per_frame, "clocks": ck.result}
        ref_px = torch.cat(keep)
        if name == backends[0][0]:
            clear = (ref_px[:, 4] - 0.6).abs() > 1e-3
            d_err = ((packed_ours[:, 3] - ref_px[:, 3]).abs() / ref_px[:, 3].abs().clamp_min(1.0))[clear]
            out["parity"] = {"rays": n, "against": name, "rgb_linf": float((packed_ours[:, 0:3] - ref_px[:, 0:3]).abs().max()),
                             "acc_linf": float((packed_ours[:, 4] - ref_px[:, 4]).abs().max()),
                             "depth_linf_clear_rays": float(d_err.max()) if d_err.numel() else None, "tolerance": 1e-4}
        del keep, ref_px
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    out["note"] = ("reference Python + reference CUDA kernel vs this package's fused path, same GPU, same frame, same weights; "
                   "speedup = reference ms_per_frame / this line's ms_per_step")
    del model
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
