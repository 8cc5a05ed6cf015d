mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== bench pulse"; timeout 200 python tools/bench_pulse.py 2>&1 | grep -E "stateless|zi/zf|wall" | cut -c1-200
