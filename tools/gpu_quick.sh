mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_min'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"
B200DSP_VARIANT=12 timeout 100 python tools/dbg_tc2.py time | tail -1
