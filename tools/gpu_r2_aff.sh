out=gpurun_out/r6a; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "front_plan or full_depth or tile_pipe or random_states" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
cfgs="c3:generic_HE:16:16:2048 c3:XXZ:16:16:2048 c3:NPQC:16:16:4096 c3:NPQC:28:20:2 c3:qg_circuit:16:8:1024 c3:generic_HE:20:8:128"
timeout 300 python tools/bench_configs.py $cfgs > $out/apply.jsonl 2> $out/apply.err
python - $out/apply.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    print(" ", j["config"], "ms", round(j["ms"], 2), "passes", j["passes"], "by-layers GB/s", round(j["algorithmic_GBps_layers"]))
PY
tail -3 $out/apply.err
