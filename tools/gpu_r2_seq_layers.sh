out=gpurun_out/r6g; mkdir -p $out
for v in spec full spec; do
  unset PQC_SEQ_LAYERS; if [ $v = full ]; then export PQC_SEQ_LAYERS=1; fi
  timeout 200 python bench.py --circuit XXZ --steps 3 --warmup 3 --samples 2048 --no-cpu-baseline > $out/xxz_$v.json 2>> $out/err.txt
  python - $out/xxz_$v.json $v <<'PY'
import json, sys
j = json.load(open(sys.argv[1]))
print(sys.argv[2], "XXZ QFIM sets/s", round(j["value"], 1), "k_layer_seq GB/s", round(j["roofline"]["achieved"]), "eqd", j["config"]["eqd_histogram"][-4:], "clk", j["clocks"]["sm_mhz"])
PY
done
tail -2 $out/err.txt
