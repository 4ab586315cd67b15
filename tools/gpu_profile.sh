#!/bin/bash
# ncu --set full captures of the QFIM step's kernels, reduced to CSV pages on the box
# (the .ncu-rep files are only kept when small: gpurun copies back at most 64 MiB).
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f \
      -o $out/$name "$@" > $out/$name.log 2>&1
  if [ -f $out/$name.ncu-rep ]; then
    ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
    ncu -i $out/$name.ncu-rep --page source --csv > $out/${name}_source.csv 2>/dev/null
    [ $(stat -c %s $out/$name.ncu-rep) -gt 6000000 ] && rm -f $out/$name.ncu-rep
  fi
}
cap layerF k_layer_pass 34 1 $B        # largest forward-pipeline pass <3,1>, 65536 CTAs
cap layerB k_layer_pass 50 2 $B        # largest backward-pipeline pass <3,1> and the closing <2,1>
cap gram k_gram_real 1 1 $B
cap jacobi k_jacobi_eigvals 1 1 $B
timeout 300 python tools/bench_configs.py c5:26 c5:28 > $out/c5.jsonl 2> $out/c5.err
cap bigpass 'k_sweep_pass|k_layer_pass' 12 3 python tools/bench_configs.py c5:26
du -sh $out; ls -la $out
