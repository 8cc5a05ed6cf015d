"""Hot spots of an ncu source page: ncu -i rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]
Prints the N SASS instructions with the most stall samples, with their dominant stall reasons, and a
per-region total (regions split at every 200 instructions) so a role's share of samples is visible."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        smp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    data.append((smp, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "instructions", len(data))
print("--- regions (every 100 instr): start idx, samples, %")
for i in range(0, len(data), 100):
    s = sum(x for x, _ in data[i:i + 100])
    if s > tot * 0.01:
        print("%5d %8d %5.1f%%  %s" % (i, s, 100.0 * s / tot, data[i][1][ix["Source"]][:60]))
print("--- top instructions")
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:N]
for i in sorted(order):
    smp, r = data[i]
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print("%5d %7d %4.1f%%  %-58s %s" % (i, smp, 100.0 * smp / tot, r[ix["Source"]][:58], " ".join("%s:%d" % (h, v) for v, h in st if v)))
