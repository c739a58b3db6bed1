#!/usr/bin/env python
"""Turn an `ncu --set full` capture of bench.py into profiles/r2_traffic.json: per kernel family the per-launch DRAM
bytes (dram__bytes_read.sum + dram__bytes_write.sum), the launch duration under ncu and the unit ncu shows busiest
(`bound`) - what bench.py's `roofline` / `rooflines` blocks quote as `traffic`, `dram_frac` and `bound`.

    gpurun -- 'ncu --set full --clock-control none --import-source on -k regex:"sample_encode|color_mlp_tc|resample_kernel|composite_kernel" \
               -c 10 -o gpurun_out/r2_full python bench.py --extras none --steps 1 --warmup 1'
    python tools/ncu_traffic.py gpurun_out/r2_full.ncu-rep [build tag] [rays per chunk]

The first launch of every family is the first chunk of the eval_800x600_waymo_gin frame: 131,072 rays, or - with bench.py's
default ray_tile_width = 800 - 128,000 (a whole number of 8-row bands)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {  # candidate limiters: metric -> name used in the bench line
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_wavefronts",
    "sm__inst_executed_realtime.avg.pct_of_peak_sustained_elapsed": "issue_slots",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed": "alu_pipe",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_pipe",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_heavy_pipe",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram",
}
EXTRA = ["sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
         "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers"]


def family(name):
    if "sample_encode_kernel<" in name:
        args = name.split("<", 1)[1].split(">", 1)[0].replace(" ", "").split(",")
        return "encode_nerf" if args[1] == "1" else "encode_prop"
    for key, fam in (("color_mlp_tc", "color_mlp"), ("resample_kernel", "resample"), ("composite_kernel", "composite")):
        if key in name:
            return fam
    return None


def main():
    rep = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, m):
        hits = [i for i, h in enumerate(hdr) if h == m] or [i for i, h in enumerate(hdr) if h.endswith("." + m)]
        if not hits:
            return None
        try:
            v = float(r[hits[0]].replace(",", ""))
        except ValueError:
            return None
        u = units[hits[0]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}.get(u, 1)

    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
    out = {"_comment": f"per-launch numbers from `ncu --set full --clock-control none` ({os.path.basename(rep)}, build {tag}): first "
                       f"{chunk:,}-ray chunk of eval_800x600_waymo_gin; written by tools/ncu_traffic.py; bench.py copies them into "
                       "roofline.traffic / dram_frac / bound",
           "workload": "eval_800x600_waymo_gin", "chunk_rays": chunk, "build": tag}
    # per family the longest launch of the capture (e.g. the level-1 resample, not the trivial level-0 one)
    body = sorted(rows[2:], key=lambda r: -(val(r, "gpu__time_duration.sum") or 0))
    for r in body:
        name = r[hdr.index("Kernel Name")]
        fam = family(name)
        if fam is None or fam in out:
            continue
        busy = {nm: val(r, m) for m, nm in UNITS.items() if val(r, m) is not None}
        bound = max(busy, key=busy.get)
        e = {"kernel": name.replace("void ", "").split("(")[0], "dram_bytes_read": val(r, "dram__bytes_read.sum"),
             "dram_bytes_write": val(r, "dram__bytes_write.sum"), "duration_ms_under_ncu": val(r, "gpu__time_duration.sum"),
             "bound": bound, "limiter_pct": {k: round(v, 1) for k, v in sorted(busy.items(), key=lambda kv: -kv[1])}}
        for m in EXTRA:
            v = val(r, m)
            if v is not None:
                e[m] = v
        out[fam] = e
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
