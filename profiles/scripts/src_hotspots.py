"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export per source line:
executed warp instructions and stall samples, top N lines and totals per file region."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0]); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or cur is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    if r[2] == '-':
        try:
            agg[(cur, ln)][0] += int(r[iI]); agg[(cur, ln)][1] += int(r[iS]); src[(cur, ln)] = r[1]
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp instructions", tot, "stall samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9d %5.1f%%  samples %6d %5.1f%%  %s:%d  %s" % (v[0], 100 * v[0] / max(1, tot), v[1], 100 * v[1] / max(1, tots), k[0], k[1], src[k][:100]))
