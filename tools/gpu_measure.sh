#!/bin/bash
# One GPU-box visit: parity tests, the bench lines and the ncu launch list
# (full ncu captures: tools/gpu_profile.sh, tools/gpu_profile_xxz.sh).
# Usage (from the repo root on the GPU box): bash tools/gpu_measure.sh [tag]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json
timeout 300 python bench.py --circuit XXZ --steps 2 --warmup 3 --samples 2048 --no-cpu-baseline > $out/bench_xxz.json 2>> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline > $out/launch_run.log 2>&1
ls -la $out
