#!/bin/bash
# front-plan parity + gate-apply timings per plan.  Usage: bash tools/gpu_r2_front.sh TAG
tag=${1:-r4e}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "front_plan or full_depth or tile_pipe" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
cfgs="c3:XXZ:16:16:2048 c3:XXZ:20:8:256 c3:generic_HE:16:16:2048 c3:NPQC:16:16:4096 c3:TFIM:16:16:4096"
PQC_FRONT=1 timeout 300 python tools/bench_configs.py $cfgs > $out/apply_front.jsonl 2> $out/apply.err
PQC_FRONT=1 PQC_FRONT_NORZZR=1 timeout 300 python tools/bench_configs.py c3:XXZ:16:16:2048 c3:XXZ:20:8:256 > $out/apply_front_norzzr.jsonl 2>> $out/apply.err
PQC_FRONT=0 timeout 300 python tools/bench_configs.py c3:XXZ:16:16:2048 c3:XXZ:20:8:256 > $out/apply_block.jsonl 2>> $out/apply.err
for f in apply_front apply_front_norzzr apply_block; do echo $f; python - $out/$f.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    print(" ", j["config"], "ms", round(j["ms"], 2), "passes", j["passes"], "by-layers GB/s", round(j["algorithmic_GBps_layers"]))
PY
done
tail -3 $out/apply.err
