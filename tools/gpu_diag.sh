timeout 200 python tools/dbg_hostpipe.py 2>&1 | tail -12
for i in 1 2 3; do B200DSP_VARIANT=12 timeout 100 python tools/dbg_tc2.py time | tail -1; done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv -lms 50 > gpurun_out/clk.csv &
SMI=$!; B200DSP_VARIANT=12 timeout 100 python tools/dbg_tc2.py time | tail -1; kill $SMI; sort gpurun_out/clk.csv | uniq -c | sort -rn | head -4
