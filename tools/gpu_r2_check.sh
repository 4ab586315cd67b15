#!/bin/bash
# Parity suite + bench lines on one B200.  Usage: bash tools/gpu_r2_check.sh TAG [pytest -k expr]
tag=${1:-r4b}
out=gpurun_out/$tag; mkdir -p $out
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$2" > $out/pytest_gpu.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
fi
echo "pytest exit $?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json
timeout 300 python bench.py --circuit XXZ --steps 2 --warmup 3 --samples 2048 --no-cpu-baseline > $out/bench_xxz.json 2>> $out/bench.err
tail -c 600 $out/bench_xxz.json; tail -5 $out/bench.err
