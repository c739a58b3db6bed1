#!/bin/bash
# A/B of the encode kernel forms (dev tool): one compact line per variant; full JSON in gpurun_out/ab_<name>.log
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
unset UCNERF_B200_LIB
run runs0 --encode-runs 0 "$@"
run runs1 --encode-runs 1 "$@"
run runs2 --encode-runs 2 "$@"
run runs3 --encode-runs 3 "$@"
export UCNERF_B200_LIB=$PWD/ucnerf_b200/csrc/build/var_minb4/lib.so
run minb4_runs1 --encode-runs 1 "$@"
unset UCNERF_B200_LIB
