#!/bin/bash
# Round 2, visit A: the tile-pipe kernel against the per-tile kernels -- parity, then A/B timings.
tag=${1:-r2a}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "tile_pipe or layer_pass or layer_sequence or multi_pass" > $out/pytest.log 2>&1; tail -15 $out/pytest.log
for pipe in 1 0; do
  PQC_PIPE=$pipe timeout 300 python tools/bench_configs.py c3:TFIM:16:16:4096 c3:XXZ:16:16:2048 c3:TFIM:24:8:16 c3:NPQC:16:16:4096 > $out/apply_pipe$pipe.jsonl 2> $out/apply_pipe$pipe.err
  PQC_PIPE=$pipe timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_pipe$pipe.json 2> $out/bench_pipe$pipe.err
done
cat $out/apply_pipe1.jsonl $out/apply_pipe0.jsonl
tail -c 1500 $out/bench_pipe1.json; echo; tail -c 1500 $out/bench_pipe0.json
tail -5 $out/bench_pipe1.err
