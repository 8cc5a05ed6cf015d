mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
echo "== bench pulse"; timeout 200 python tools/bench_pulse.py 2>&1 | grep -v "wall" | cut -c1-200 | tail -24
