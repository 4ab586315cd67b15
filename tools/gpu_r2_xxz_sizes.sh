out=gpurun_out/r4f; mkdir -p $out
cfgs="c3:XXZ:12:16:8192 c3:XXZ:14:16:4096 c3:XXZ:16:16:2048 c3:XXZ:18:8:512 c3:XXZ:24:4:16"
timeout 300 python tools/bench_configs.py $cfgs > $out/apply_default.jsonl 2> $out/apply.err
PQC_FRONT=1 timeout 300 python tools/bench_configs.py $cfgs > $out/apply_front.jsonl 2>> $out/apply.err
PQC_FRONT=0 timeout 300 python tools/bench_configs.py $cfgs > $out/apply_block.jsonl 2>> $out/apply.err
for f in apply_default apply_front apply_block; do echo $f; python - $out/$f.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    print(" ", j["config"], "ms", round(j["ms"], 2), "passes", j["passes"], "by-layers GB/s", round(j["algorithmic_GBps_layers"]))
PY
done
tail -3 $out/apply.err
