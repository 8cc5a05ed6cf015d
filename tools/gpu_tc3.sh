export B200DSP_VARIANT=12
for d in 0 1 2 4 3 5 6 7; do B200DSP_TC_DBG=$d timeout 100 python tools/dbg_tc2.py time 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 1 -f -o gpurun_out/prof_fir_tc2 python tools/dbg_tc2.py time > gpurun_out/ncu_tc2.log 2>&1; tail -2 gpurun_out/ncu_tc2.log
