timeout 250 python tools/bench_rate12.py 2>&1 | tail -9
timeout 600 python -m pytest tests -x -q -m gpu -k "dn or sharded or golden or round2" 2>&1 | tail -4
