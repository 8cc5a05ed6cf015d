#!/bin/bash
# round-2 final evidence, part 1: whole GPU test suite, smoke, fir_fft ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fir_fft -s 2 -c 1 -o gpurun_out/r02_fir_fft python tools/prof_r2.py fir_fft_c64 > /dev/null 2>&1
tail -3 gpurun_out/r02_pytest_gpu.txt; tail -2 gpurun_out/r02_smoke.txt; ls -la gpurun_out/r02_fir_fft.ncu-rep
