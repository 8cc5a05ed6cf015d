for v in 12 14 13; do B200DSP_VARIANT=$v timeout 100 python tools/dbg_tc2.py time 2>&1 | tail -1; done
for v in 12 14; do B200DSP_VARIANT=$v timeout 100 python tools/dbg_tc2.py err 2>&1 | head -1; done
B200DSP_VARIANT=14 B200DSP_TC_DBG=8 timeout 100 python tools/dbg_tc2.py time 2>&1 | grep tc2 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or cfg2 or halo" 2>&1 | tail -3
