#!/bin/bash
# ncu --set full of k_tile_pipe on front-planner passes (XXZ 16q and NPQC 16q)
tag=${1:-r2g}
out=gpurun_out/$tag; mkdir -p $out
cap() {  # name skip count cmd...
  local name=$1 skip=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pipe -s $skip -c $cnt -f -o $out/$name "$@" > $out/$name.log 2>&1
  ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
  ncu -i $out/$name.ncu-rep --page source --csv > $out/${name}_source.csv 2>/dev/null
  ncu -i $out/$name.ncu-rep --page details > $out/${name}_details.txt 2>/dev/null
  rm -f $out/$name.ncu-rep
}
cap xxz 12 2 python tools/bench_configs.py c3:XXZ:16:16:1024
cap npqc 3 2 python tools/bench_configs.py c3:NPQC:16:16:1024
cap he 10 2 python tools/bench_configs.py c3:generic_HE:16:16:1024
ls -la $out
