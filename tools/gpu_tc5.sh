timeout 200 python tools/dbg_tc.py 12 2>&1 | tail -12
export B200DSP_VARIANT=12
for d in 8 12 15; do echo "--- dbg $d"; B200DSP_TC_DBG=$d timeout 100 python tools/dbg_tc2.py time 2>&1 | grep -E "tc2|dbg" | tail -5; done
B200DSP_VARIANT=12 timeout 100 python tools/dbg_tc2.py err | tail -2
