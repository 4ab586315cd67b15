#!/bin/bash
# Re-capture k_tile_pipe (ncu --set full) on the final build + the bench line.
tag=${1:-r3d}
out=gpurun_out/$tag; mkdir -p $out
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f \
      -o $out/$name "$@" > $out/$name.log 2>&1
  if [ -f $out/$name.ncu-rep ]; then
    ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
    ncu -i $out/$name.ncu-rep --page source --csv > $out/${name}_source.csv 2>/dev/null
    rm -f $out/$name.ncu-rep
  fi
}
cap pipe_npqc16 k_tile_pipe 2 2 python tools/bench_configs.py c3:NPQC:16:16:1024
cap pipe_npqc28 k_tile_pipe 5 5 python tools/bench_configs.py c3:NPQC:28:20:2
cap pipe_he16 k_tile_pipe 8 2 python tools/bench_configs.py c3:generic_HE:16:16:1024
cat > $out/fidblock.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from pyramaterised_b200 import engine
A = torch.randn(20, 1 << 26, dtype=torch.complex128, device="cuda")
B = torch.randn(4, 1 << 26, dtype=torch.complex128, device="cuda")
for _ in range(3):
    engine.fidelity_hist(A, B, bins=7)
torch.cuda.synchronize()
PY
cap fidblock 'k_fid_block$' 1 1 python $out/fidblock.py
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err
ls $out
