#!/bin/bash
# ncu captures of the XXZ-16 QFIM step (generic sweep kernel) + launch list, reduced to CSV.
tag=${1:-xxz}
out=gpurun_out/$tag
mkdir -p $out
B="python bench.py --circuit XXZ --steps 1 --warmup 3 --samples 256 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv $B > $out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_layer_seq|k_sweep_pass' -s 150 -c 4 -f -o $out/sweep $B > $out/sweep.log 2>&1
ncu -i $out/sweep.ncu-rep --page raw --csv > $out/sweep_raw.csv 2>/dev/null
ncu -i $out/sweep.ncu-rep --page source --csv > $out/sweep_source.csv 2>/dev/null
rm -f $out/sweep.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_magic12 -s 1 -c 1 -f -o $out/magic python tools/bench_configs.py c4 > $out/magic.log 2>&1
ncu -i $out/magic.ncu-rep --page raw --csv > $out/magic_raw.csv 2>/dev/null
ncu -i $out/magic.ncu-rep --page source --csv > $out/magic_source.csv 2>/dev/null
rm -f $out/magic.ncu-rep
du -sh $out
