#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (post-change)"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== aux"; timeout 600 python tools/bench_aux.py 2>&1 | grep -E "upsample|sos|up4" 
echo "== bench n1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e'], d['cpu_baseline'])"
echo "== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-700
echo "== sharded parity n2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_check.py 2>&1 | tail -8
