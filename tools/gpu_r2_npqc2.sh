out=gpurun_out/r6i; mkdir -p $out
cfgs="c3:NPQC:16:16:4096 c3:NPQC:20:20:128 c3:NPQC:24:16:16 c3:NPQC:28:20:2"
for v in npqc2 nonpqc2 npqc2; do
  unset PQC_PIPE_OPSET; if [ $v = nonpqc2 ]; then export PQC_PIPE_OPSET=nonpqc2; fi
  timeout 200 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo $v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
done
tail -2 $out/apply.err
