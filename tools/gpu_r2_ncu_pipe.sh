#!/bin/bash
# ncu --set full of k_tile_pipe and of k_layer_pass on the TFIM-16 apply-only workload
tag=${1:-r2b}
out=gpurun_out/$tag; mkdir -p $out
cmd="python tools/bench_configs.py c3:TFIM:16:16:1024"
cap() {  # name regex env
  env $3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 20 -c 2 -f -o $out/$1 $cmd > $out/$1.log 2>&1
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
  ncu -i $out/$1.ncu-rep --page source --csv > $out/$1_source.csv 2>/dev/null
  ncu -i $out/$1.ncu-rep --page details > $out/$1_details.txt 2>/dev/null
  rm -f $out/$1.ncu-rep
}
cap pipe k_tile_pipe PQC_PIPE=1
cap layer k_layer_pass PQC_PIPE=0
ls -la $out
