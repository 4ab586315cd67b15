#!/bin/bash
# Round 2, N GPUs of one box: NCCL parity check, bench line with the exchange leg, config-5 path.
# Usage: bash tools/gpu_r2_multi.sh TAG N "n p S block [resident]" ...
tag=$1; N=$2; shift 2
out=gpurun_out/$tag; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_nccl_check.py > $out/nccl_check.log 2>&1; grep "^{" $out/nccl_check.log | tail -1
NCCL_DEBUG=INFO timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_${N}gpu.json 2> $out/bench_${N}gpu.err
cat $out/bench_${N}gpu.json | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print(round(b['value']), 'sets/s', b['n_gpus'], 'gpus; exchange:', json.dumps(b.get('exchange')))"
grep -m3 -i "NVLS\|via P2P\|Using network" $out/bench_${N}gpu.err | cut -c1-200
for a in "$@"; do
  timeout 900 $TR --master-port 29513 tools/c5_streamed.py $a 2>> $out/c5.err | grep "^{" | tee -a $out/c5_${N}gpu.jsonl
done
tail -2 $out/c5.err 2>/dev/null | cut -c1-300
