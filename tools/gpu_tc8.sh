for i in 1 2; do B200DSP_VARIANT=12 timeout 100 python tools/dbg_tc2.py time | tail -1; done
B200DSP_VARIANT=12 B200DSP_TC_DBG=8 timeout 100 python tools/dbg_tc2.py time 2>&1 | grep tc2 | tail -4
timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -3
