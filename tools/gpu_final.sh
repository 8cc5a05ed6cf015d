mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-250
