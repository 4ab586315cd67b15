#!/bin/bash
# X-sum gather as thread-block clusters (DSMEM for the partner chunks): parity + headline bench per
# cluster size.  Usage: bash tools/gpu_r2_xsum.sh TAG
tag=${1:-r4a}
out=gpurun_out/$tag; mkdir -p $out
for cl in 4 16; do
  PQC_XSUM_CLUSTER=$cl timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "TFIM or qfim or QFIM" \
      > $out/pytest_cl$cl.log 2>&1
  tail -2 $out/pytest_cl$cl.log
done
for cl in 1 2 4 8 16 1; do
  PQC_XSUM_CLUSTER=$cl timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      >> $out/bench_cl$cl.json 2>> $out/bench.err
  python - $out/bench_cl$cl.json $cl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    j = json.loads(line)
    print("cluster", sys.argv[2], "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "eqd",
          j["config"]["eqd_histogram"][16:18], "clk", j["clocks"]["sm_mhz"])
PY
done
