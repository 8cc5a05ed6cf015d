#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== sharded parity + timing n2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_check.py 2>&1 | grep -E "rank|Error|error" | tail -8
echo "== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('n2', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
echo "== ncu tc2"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 1 -f -o gpurun_out/prof_fir_tc2_final python tools/dbg_tc2.py time > gpurun_out/ncu_tc2.log 2>&1; tail -2 gpurun_out/ncu_tc2.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; grep -c fir_tc2 gpurun_out/launches.csv
