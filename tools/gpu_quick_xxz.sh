#!/bin/bash
# Quick A/B visit on the XXZ-16 workload: selected GPU tests + the bench line with / without a knob.
tag=$1; kexpr=$2; knob=$3
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -k "$kexpr" > $out/pytest.log 2>&1; tail -15 $out/pytest.log
B="python bench.py --circuit XXZ --steps 2 --warmup 3 --samples 2048 --no-cpu-baseline"
timeout 600 $B > $out/bench_a.json 2> $out/bench_a.err
[ -n "$knob" ] && env $knob timeout 600 $B > $out/bench_b.json 2> $out/bench_b.err
python - <<PY
import json
for f in ("a","b"):
    try:
        b=json.load(open("$out/bench_%s.json"%f))
        print(f, round(b["value"]), "sets/s; pass kernel", round(b["roofline"]["achieved"]), "GB/s", round(b["roofline"]["frac"],3), "share", round(b["roofline"]["kernel_share_of_step"],3), "apply-only", round(b["roofline_apply_only"]["pass_kernel_GBps"]), round(b["roofline_apply_only"]["states_per_s"]), b["config"]["eqd_histogram"][60:], b["clocks"]["sm_mhz"], b["clocks"]["reasons"])
    except Exception as e:
        print(f, "no line", e); print(open("$out/bench_%s.err"%f).read()[-1500:])
PY
