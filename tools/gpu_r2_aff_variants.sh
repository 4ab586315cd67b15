out=gpurun_out/r6b; mkdir -p $out
cfgs="c3:XXZ:16:16:2048 c3:generic_HE:16:16:2048 c3:NPQC:16:16:4096"
for v in cur affA affB affOld cur; do
  if [ $v = cur ]; then unset PQC_LIB_PATH; else export PQC_LIB_PATH=$PWD/pyramaterised_b200/variants/lib$v.so; fi
  timeout 200 python tools/bench_configs.py $cfgs > $out/apply_$v.jsonl 2>> $out/apply.err
  echo $v; python - $out/apply_$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f ms" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"]) for l in open(sys.argv[1])))
PY
done
tail -3 $out/apply.err
