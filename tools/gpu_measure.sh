#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the ncu launch list and full captures.
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
B="python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_layer_pass<3, 1>' -s 48 -c 3 -f -o $out/layer_full $B > $out/layer_full.log 2>&1
[ -f $out/layer_full.ncu-rep ] || timeout 600 ncu --set full --clock-control none -k regex:k_layer_pass -s 132 -c 44 -f -o $out/layer_full $B > $out/layer_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gram_real -s 3 -c 2 -f -o $out/gram_full $B > $out/gram_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_xsum_gather -s 50 -c 2 -f -o $out/xsum_full $B > $out/xsum_full.log 2>&1
ls -la $out
