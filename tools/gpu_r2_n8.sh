#!/bin/bash
# round-2 eight-GPU check: the driver's own launch line for bench.py (weak scaling, peer-memory halo), kept as evidence
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 3 > gpurun_out/bench_n8_peer.json 2> gpurun_out/bench_n8_peer.err; tail -2 gpurun_out/bench_n8_peer.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n8_peer.json')); print('n8 peer', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'] and d['e2e']['value'], d['gpu_launches'], d['parity']['boundary_max_err'], d['config']['parallelism'][:80])"
