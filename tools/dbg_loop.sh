for i in 1 2 3 4 5 6; do timeout 80 python tools/dbg_stc3.py sos_tenband 0 8 4194304 zf 2>&1 | grep -v "Search\|Compile\|debugging\|CUDA kernel errors" | sort | uniq -c | head -30; echo ---; done
