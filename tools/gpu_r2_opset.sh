#!/bin/bash
# op-set instances of k_tile_pipe: parity, then each family on its instance vs the catch-all (same binary)
out=gpurun_out/r6d; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "front_plan or full_depth or tile_pipe" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
cfgs="c3:XXZ:16:16:2048 c3:XXZ:20:8:256 c3:XXZ:12:16:8192 c3:generic_HE:16:16:2048 c3:generic_HE:20:8:128 c3:NPQC:16:16:4096 c3:NPQC:28:20:2 c3:qg_circuit:16:8:1024"
for v in sets all sets; do
  unset PQC_PIPE_OPSET
  if [ $v = all ]; then export PQC_PIPE_OPSET=all; fi
  timeout 300 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo $v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
done
tail -3 $out/apply.err
