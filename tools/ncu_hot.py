"""Top stall sites of an ncu --page source --csv dump (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci = h.index("Warp Stall Sampling (All Samples)"); src = h.index("Source"); ie = h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ci] or 0), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
texec = sum(float(d[2][ie] or 0) for d in data)
print(f"total samples {tot:.0f}, total warp-instructions {texec:.0f}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for v, k, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{v / tot * 100:5.1f}%  line {k:4d} exec={float(r[ie]):>9.0f}  {r[src].strip()[:70]:70s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
