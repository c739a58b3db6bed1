#!/bin/bash
# First GPU call of the next round: validate the training ops that were written without a
# GPU, take their numbers, and capture the pooled backward for the atomic-throughput question.
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/first_gpu_call.sh'
# Everything lands in gpurun_out/n1_*.  Budget: ~3 GPU-minutes.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_train_pooled_encode.py tests/test_train_render_composite.py tests/test_train_resample.py \
    tests/test_train_sample_cast_rays.py tests/test_train_zz_level_chain.py tests/test_train_zzz_forward.py -m gpu -q \
    > gpurun_out/n1_train_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/n1_train_pytest.log
timeout 60 python bench_pooled.py --reps 5 > gpurun_out/n1_bench_pooled.log 2>&1
timeout 90 python bench_train.py --steps 10 --warmup 3 > gpurun_out/n1_bench_train.log 2>&1
timeout 60 python bench_train.py --steps 10 --warmup 3 --merge-runs ray > gpurun_out/n1_bench_train_rayruns.log 2>&1
# per-launch times of one training step (shares of the step; serialised, cold caches)
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n1_train_launches.csv \
    python bench_train.py --steps 1 --warmup 1 > gpurun_out/n1_ncu_train.log 2>&1
# the pooled backward: where do the reductions go (L2 red sectors, LSU red accesses, DRAM)
timeout 180 ncu --set full --clock-control none --import-source on -k regex:pooled_backward_kernel -c 2 \
    -o gpurun_out/n1_pooled_backward python bench_pooled.py --reps 1 > gpurun_out/n1_ncu_pooled.log 2>&1
tail -3 gpurun_out/n1_train_pytest.log
tail -c 600 gpurun_out/n1_bench_train.log
