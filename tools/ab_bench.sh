#!/bin/bash
# A/B bench of library variants built under ucnerf_b200/csrc/build/var_*/lib.so (dev tool; UCNERF_B200_LIB selects the .so)
# usage: tools/ab_bench.sh [bench args...]   -> one compact line per variant, full JSON in gpurun_out/ab_<name>.log
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
run main "$@"
for so in ucnerf_b200/csrc/build/var_*/lib.so; do
  n=$(basename $(dirname $so)); export UCNERF_B200_LIB=$PWD/$so; run $n "$@"
done
