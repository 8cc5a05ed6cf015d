mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 1 -f -o gpurun_out/prof_fir_tc2_final python tools/dbg_tc2.py time > gpurun_out/ncu_tc2.log 2>&1; tail -1 gpurun_out/ncu_tc2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; grep -c fir_tc2 gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none -k regex:fir_tc_real -s 6 -c 2 -f -o gpurun_out/prof_fir_tc_real python tools/prof_cfg3.py > gpurun_out/ncu_tcr.log 2>&1; tail -1 gpurun_out/ncu_tcr.log
timeout 600 ncu --set full --clock-control none -k regex:sos_pass -s 2 -c 2 -f -o gpurun_out/prof_sos python tools/bench_aux_sos.py > gpurun_out/ncu_sos.log 2>&1; tail -1 gpurun_out/ncu_sos.log
