mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 600 python tools/bench_aux.py 2>&1 | grep -E "sos|cfg4"
