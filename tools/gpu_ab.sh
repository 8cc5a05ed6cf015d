# A/B helper: FIR headline + SOS cfg4 timing for the wait-loop variants (same box, interleaved)
for rep in 1 2; do
for v in 0 1 2 3; do
export B200DSP_LIB=/root/repo/build_ab/lib_v$v.so
python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v fir', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
python tools/dbg_stc_time.py sos6 0 2>&1 | tail -1
done; done
