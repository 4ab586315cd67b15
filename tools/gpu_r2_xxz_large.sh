out=gpurun_out/r6f; mkdir -p $out
cfgs="c3:XXZ:24:4:16 c3:XXZ:26:4:4 c3:XXZ:28:2:2 c3:XXZ:22:4:64"
for v in 1 0; do
  PQC_FRONT=$v timeout 300 python tools/bench_configs.py $cfgs > $out/apply_front$v.jsonl 2>> $out/apply.err
  echo front=$v; python - $out/apply_front$v.jsonl <<'PY'
import json, sys
print("  " + " | ".join("%s %.2f (%d passes)" % (json.loads(l)["config"].split("only ")[1].split(" layers")[0], json.loads(l)["ms"], json.loads(l)["passes"]) for l in open(sys.argv[1])))
PY
done
tail -3 $out/apply.err
