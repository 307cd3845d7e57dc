"""top source lines by warp-stall samples of one kernel in an ncu report: python scripts/ncu_top_stalls.py rep.ncu-rep [n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Samples" in c for c in r):
        hdr = i; break
if hdr is None:
    print(out[:2000]); sys.exit(0)
h = rows[hdr]
ci = {c: k for k, c in enumerate(h)}
samp = next(c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
src = "Source"
tot = 0; items = []
for r in rows[hdr + 1:]:
    if len(r) <= ci[samp]: continue
    try: v = float(r[ci[samp]])
    except ValueError: continue
    tot += v; items.append((v, r[ci[src]][:110], r))
items.sort(key=lambda t: -t[0])
print("columns:", [c for c in h][:12])
print("total samples", tot)
for v, s_, r in items[:n]:
    print("%7.0f %5.1f%%  %s" % (v, 100 * v / max(tot, 1), s_))
