#!/bin/bash
# First GPU pass: smoke, parity tests, bench, launch list.  Run under gpurun from the repo root.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pcie"; timeout 120 python tools/pcie_bw.py 2>&1 | tail -6
echo "== dbg sos"; timeout 250 python tools/dbg_sos.py 2>&1 | grep -v None | tail
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu" ; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
for v in 1 2 3 4 5; do
  echo "== bench variant $v"; timeout 300 python bench.py --steps 5 --warmup 3 --variant $v --no-e2e --no-cpu-baseline > gpurun_out/bench_v$v.json 2>&1; cat gpurun_out/bench_v$v.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -8 gpurun_out/launches.csv
