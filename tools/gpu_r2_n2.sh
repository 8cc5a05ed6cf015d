#!/bin/bash
# round-2 two-GPU bundle: sharded parity (kept as evidence), bench weak (peer / nccl halo), bench strong 2^31
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== sharded parity (FIR nccl/peer halo, up/dn, IIR state carry)"
timeout 600 $TR --master-port 29512 tools/sharded_check.py 2>&1 | grep -E "^rank|Error|error" | sort > gpurun_out/sharded_check_n2.txt; cat gpurun_out/sharded_check_n2.txt | tail -20
for mode in peer nccl; do
echo "== bench n2 weak, halo $mode"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --halo $mode > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; tail -2 gpurun_out/bench_n2_$mode.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$mode.json')); print('n2 $mode', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'] if d['e2e'] else None, d['gpu_launches'], d['parity']['boundary_max_err'], d['config']['parallelism'][:60])"
done
echo "== bench n2 strong 2^31"
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --total 2147483648 --no-e2e > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err; tail -2 gpurun_out/bench_n2_strong.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_strong.json')); print('n2 strong', d['value'], d['ms_per_step'], d['roofline']['frac'], d['scaling'], d['config']['samples_per_gpu'], d['parity']['boundary_max_err'])"
