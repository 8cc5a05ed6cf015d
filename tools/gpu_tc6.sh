timeout 200 python tools/dbg_tc.py 13 2>&1 | tail -12
for v in 10 12 13; do
export B200DSP_VARIANT=$v
for d in 8 12; do echo "--- variant $v dbg $d"; B200DSP_TC_DBG=$d timeout 100 python tools/dbg_tc2.py time 2>&1 | grep -E "tc2|dbg" | tail -5; done
done
B200DSP_VARIANT=10 timeout 100 python tools/dbg_tc2.py err | tail -2
