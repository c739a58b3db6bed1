#!/bin/bash
# One GPU call that re-validates everything and refreshes the evidence under gpurun_out/ (copy what should be judged into
# profiles/):   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'      (~4 GPU-minutes)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest.log
timeout 200 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v_bench_reference_arm.json 2>&1
timeout 120 python bench_train.py > gpurun_out/v_train.json 2>&1
timeout 120 python bench_train.py --sky --steps 5 > gpurun_out/v_train_sky.json 2>&1
timeout 60 python bench_pooled.py --reps 5 > gpurun_out/v_pooled.json 2>&1
# per-launch times of one frame (shares of the step; serialised, cold caches)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv \
    python bench.py --extras none --steps 2 --warmup 1 > gpurun_out/v_ncu_bench.log 2>&1
tail -3 gpurun_out/v_pytest.log
tail -c 400 gpurun_out/v_bench.json
