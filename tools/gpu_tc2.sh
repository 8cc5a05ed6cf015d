for o in 43210 23410 21430 32410 01234; do B200DSP_TC_ORDER=$o timeout 100 python tools/dbg_tc2.py err 2>&1 | tail -2; done
for d in 0 1 2 4 3 5 6 7; do B200DSP_TC_DBG=$d timeout 100 python tools/dbg_tc2.py time 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv -lms 100 > gpurun_out/tc_clocks.csv &
SMI=$!
timeout 100 python tools/dbg_tc2.py time | tail -1
kill $SMI
sort gpurun_out/tc_clocks.csv | uniq -c | sort -rn | head -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc -s 3 -c 1 -f -o gpurun_out/prof_fir_tc python tools/dbg_tc2.py time > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
