#!/bin/bash
# round-2 four-GPU check: the driver's own launch line for bench.py (weak scaling, peer-memory halo), kept as evidence
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 --e2e-steps 4 > gpurun_out/bench_n4_peer.json 2> gpurun_out/bench_n4_peer.err; tail -2 gpurun_out/bench_n4_peer.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n4_peer.json')); print('n4 peer', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['gpu_launches'], d['parity']['boundary_max_err'], d['config']['parallelism'][:80])"
