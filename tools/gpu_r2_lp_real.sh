out=gpurun_out/r6h; mkdir -p $out
for v in spec full spec full; do
  unset PQC_LP_REAL; if [ $v = full ]; then export PQC_LP_REAL=1; fi
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/tfim_$v.json 2>> $out/err.txt
  python - $out/tfim_$v.json $v <<'PY'
import json, sys
j = json.load(open(sys.argv[1]))
print(sys.argv[2], "TFIM sets/s", round(j["value"], 1), "k_layer_pass GB/s", round(j["roofline"]["achieved"]), "apply states/s", round(j["roofline_apply_only"]["states_per_s"]), "eqd", j["config"]["eqd_histogram"][16:], "clk", j["clocks"]["sm_mhz"])
PY
done
tail -2 $out/err.txt
