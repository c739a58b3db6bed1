#!/bin/bash
# A/B of colour-MLP experiments (dev tool)
mkdir -p gpurun_out
run() {
  name=$1; shift
  python bench.py --no-cpu-baseline "$@" > gpurun_out/ab_$name.log 2>&1
  tail -1 gpurun_out/ab_$name.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
    print('$name', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, d['checksum_mean_rgb'], 'e2e', round(d['e2e']['ms_per_step'],2), 'cam', round(d['e2e_camera']['ms_per_step'],2))
except Exception as e: print('$name', 'FAILED', e)
"
}
run smembias --steps 4 --warmup 3
run globalbias --steps 4 --warmup 3 --tc-debug 16
run smembias_1024 --steps 3 --warmup 3 --workload target_1024spp
