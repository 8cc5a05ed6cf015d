for v in 15 12 15 12; do B200DSP_VARIANT=$v timeout 100 python tools/dbg_tc2.py time | tail -1; done
B200DSP_VARIANT=15 timeout 100 python tools/dbg_tc2.py err | head -1
B200DSP_VARIANT=15 B200DSP_TC_DBG=8 timeout 100 python tools/dbg_tc2.py time 2>&1 | grep tc2 | tail -4
