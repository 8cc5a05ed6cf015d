#!/bin/bash
mkdir -p gpurun_out
echo "== sharded parity (FIR nccl/peer halo, IIR state carry) + timing"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_check.py 2>&1 | grep -E "rank|Error|error" | sort | tail -12
echo "== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('n2', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
echo "== aux (gpu 0)"; CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/bench_aux.py 2>&1 | grep -E "sos|cfg" | cut -c1-200
