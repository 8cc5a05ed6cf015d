for v in 0 12 0 15; do B200DSP_VARIANT=$v timeout 100 python tools/dbg_tc2.py time | tail -1; done
B200DSP_VARIANT=0 timeout 100 python tools/dbg_tc2.py err 2>&1 | head -1
B200DSP_VARIANT=0 B200DSP_TC_DBG=8 timeout 100 python tools/dbg_tc2.py time 2>&1 | grep -E "tc2" | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or cfg2 or halo or host_pipeline" 2>&1 | tail -2
