"""Wait sites of a warp-specialised kernel from an ncu source page (csv): samples parked on each mbarrier try_wait
loop, i.e. how much of its time each role spends waiting.  python tools/ncu_waits.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot)
for i, r in enumerate(data):
    src = r[ix["Source"]]
    if "SYNCS.PHASECHK" in src:
        s = sum(int(x[ix["# Samples"]]) for x in data[i:i + 16])
        if s > tot * 0.0005:
            print("%5d %7d %5.1f%%  %s" % (i, s, 100.0 * s / tot, src.strip()[:80]))
