#!/bin/bash
mkdir -p gpurun_out
echo "== aux"; timeout 600 python tools/bench_aux.py 2>&1 | tail -12
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
echo "== ncu full fir"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_poly -s 3 -c 1 -f -o gpurun_out/prof_fir python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fir.log 2>&1; tail -2 gpurun_out/ncu_fir.log
echo "== ncu full sos"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sos_pass -s 2 -c 2 -f -o gpurun_out/prof_sos python tools/bench_aux_sos.py > gpurun_out/ncu_sos.log 2>&1; tail -2 gpurun_out/ncu_sos.log
ls -la gpurun_out/*.ncu-rep
