out=gpurun_out/r9; mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_depth or front_plan" > $out/pytest.log 2>&1; tail -2 $out/pytest.log
cfgs="c3:TFIM:18:16:512 c3:TFIM:20:8:256 c3:TFIM:28:2:2"
for v in default 1 0; do
  unset PQC_FRONT; if [ $v != default ]; then export PQC_FRONT=$v; fi
  timeout 200 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo front=$v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f (%d passes)" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"], json.loads(l)["passes"]) for l in open(sys.argv[1])))
PY
done
tail -2 $out/apply.err
