#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
echo "== sharded parity n$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/sharded_check.py 2>&1 | grep -E "rank|Error|error" | sort | tail -24
for n in 4 8; do
  if [ $n -le $N ]; then
  echo "== bench n$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -3 gpurun_out/bench_n$n.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print('n$n', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
  echo "== bench n$n reference arm"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 3 --warmup 3 2>/dev/null | cut -c1-200
  fi
done
