mkdir -p gpurun_out
bash tools/gpu_final.sh
echo "== ncu fir_tc2"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 1 -f -o gpurun_out/prof_fir_tc2_final2 python tools/dbg_tc2.py time > gpurun_out/ncu_tc2.log 2>&1; tail -1 gpurun_out/ncu_tc2.log
