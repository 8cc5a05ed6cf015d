mkdir -p gpurun_out
timeout 300 python tools/dbg_tcr.py 2>&1 | grep -E "dn4|dn3|variant|hist" | tail -14
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/bench_aux.py 2>&1 | tail -10
