#!/bin/bash
# round-2 final evidence: whole GPU test suite, smoke, bench N=1 (un-profiled), then the ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > /dev/null 2>&1
tail -3 gpurun_out/r02_pytest_gpu.txt; tail -2 gpurun_out/r02_smoke.txt; tail -3 gpurun_out/r02_bench_n1_final.err | cut -c1-400
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_n1_final.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'], d['parity']['boundary_max_err'])
for s in d['secondary']:
    print('  %-62s %.4f ms  %.3f' % (s['config'][:62], s['ms'], s['roofline']['frac']))
PY
grep -c . gpurun_out/r02_bench_launch_list.csv
