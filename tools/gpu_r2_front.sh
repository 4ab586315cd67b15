#!/bin/bash
# Round 2: the front planner on k_tile_pipe -- parity, then gate-apply throughput by template layers.
tag=${1:-r2f}
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x -k "front_plan or full_depth or tile_pipe" > $out/pytest.log 2>&1; tail -15 $out/pytest.log
cfgs="c3:XXZ:16:16:2048 c3:generic_HE:16:16:2048 c3:NPQC:16:16:4096 c3:TFIM:16:16:4096 c3:NPQC:24:16:16 c3:NPQC:28:8:2 c3:NPQC:28:20:2 c3:XXZ:20:8:256"
for fr in 1 0; do
  PQC_FRONT=$fr timeout 600 python tools/bench_configs.py $cfgs > $out/apply_front$fr.jsonl 2> $out/apply_front$fr.err
done
cat $out/apply_front1.jsonl; echo; cat $out/apply_front0.jsonl; tail -3 $out/apply_front1.err
