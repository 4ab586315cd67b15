#!/bin/bash
out=gpurun_out/r6e; mkdir -p $out
cfgs="c3:XXZ:16:16:2048 c3:XXZ:20:8:256 c3:generic_HE:16:16:2048 c3:generic_HE:20:8:128 c3:NPQC:16:16:4096 c3:qg_circuit:16:8:1024"
for v in cur inl shear cur; do
  unset PQC_LIB_PATH PQC_PIPE_OPSET
  if [ $v != cur ]; then export PQC_LIB_PATH=$PWD/pyramaterised_b200/variants/lib$v.so; fi
  timeout 200 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo $v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
  if [ $v != shear ]; then
    PQC_PIPE_OPSET=all timeout 200 python tools/bench_configs.py c3:generic_HE:16:16:2048 c3:qg_circuit:16:8:1024 > $out/apply_${v}_all.jsonl 2>> $out/apply.err
    echo "$v, catch-all instance"; python - $out/apply_${v}_all.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
  fi
done
tail -3 $out/apply.err
