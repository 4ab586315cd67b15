#!/bin/bash
# shear-form XY / R_zz ops: parity, then timings against the previous build on the same box
out=gpurun_out/r6c; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "front_plan or full_depth or tile_pipe" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
cfgs="c3:XXZ:16:16:2048 c3:XXZ:20:8:256 c3:XXZ:12:16:8192 c3:generic_HE:16:16:2048 c3:NPQC:16:16:4096"
for v in cur noshear prev cur; do
  unset PQC_LIB_PATH PQC_FRONT_NOSHEAR
  if [ $v = prev ]; then export PQC_LIB_PATH=$PWD/pyramaterised_b200/variants/libprev.so; fi
  if [ $v = noshear ]; then export PQC_FRONT_NOSHEAR=1; fi
  timeout 200 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo $v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f ms" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
done
tail -3 $out/apply.err
