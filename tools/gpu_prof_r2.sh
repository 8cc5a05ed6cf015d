#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one --set full capture per kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
cap() { # name regex which
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/r02_$1 python tools/prof_r2.py $3 > /dev/null 2>&1; ls -la gpurun_out/r02_$1.ncu-rep 2>/dev/null | awk '{print $5, $9}'
}
cap fir_tc2 fir_tc2 fir_c64
cap sos_tc sos_tc sos_f32
cap sos_pass1_f64 sos_pass1 sos_f64
cap sos_tile_scan_f64 sos_tile_scan sos_f64
cap sos_pass2_f64 sos_pass2 sos_f64
cap fir_poly_f64 fir_poly fir_f64
cap fir_up_short fir_up_short pulse_c64
cap upsample upsample upsample
cap downsample downsample downsample
