export B200DSP_VARIANT=12
for d in 8 15 9 10 12; do echo "--- dbg $d"; B200DSP_TC_DBG=$d timeout 100 python tools/dbg_tc2.py time 2>&1 | grep -E "tc2|dbg" | tail -5; done
