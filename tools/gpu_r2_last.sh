#!/bin/bash
# Last visit of round 2 (one B200): smoke, parity suite, bench lines, configs, launch list, two captures.
tag=${1:-r5}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 400 $out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2>> $out/bench.err
timeout 300 python bench.py --circuit XXZ --steps 2 --warmup 3 --samples 2048 --no-cpu-baseline > $out/bench_xxz.json 2>> $out/bench.err
cfgs="c3:TFIM:16:16:4096 c3:XXZ:16:16:2048 c3:generic_HE:16:16:2048 c3:NPQC:16:16:4096 c3:NPQC:20:20:128 c3:NPQC:24:16:16 c3:NPQC:28:20:2 c3:XXZ:20:8:256 c3:TFIM:24:8:16"
timeout 600 python tools/bench_configs.py $cfgs > $out/apply_default.jsonl 2> $out/apply.err
timeout 900 python tools/bench_configs.py c1 c2 c4 c5:28 > $out/configs.jsonl 2> $out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline > $out/launch_run.log 2>&1
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f \
      -o $out/$name "$@" > $out/$name.log 2>&1
  if [ -f $out/$name.ncu-rep ]; then
    ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
    ncu -i $out/$name.ncu-rep --page source --csv > $out/${name}_source.csv 2>/dev/null
    rm -f $out/$name.ncu-rep
  fi
}
cap pipe_xxz16 k_tile_pipe 3 2 python tools/bench_configs.py c3:XXZ:16:16:1024
cap mw_small k_mw_small 0 1 python tools/bench_configs.py c2:20000
du -sh $out; ls $out
