#!/bin/bash
# colour-MLP experiments (dev tool)
mkdir -p gpurun_out
run() {
  name=$1; shift
  python bench.py --no-cpu-baseline "$@" > gpurun_out/ab_$name.log 2>&1
  tail -1 gpurun_out/ab_$name.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
    print('$name', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, d['checksum_mean_rgb'])
except Exception as e: print('$name', 'FAILED', e)
"
}
run joint --steps 4 --warmup 3
run joint_1024 --steps 3 --warmup 3 --workload target_1024spp
