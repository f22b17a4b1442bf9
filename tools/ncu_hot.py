"""Top SASS instructions by stall samples from an ncu-rep source page, with dominant stall reason.
    python tools/ncu_hot.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rd[0]; rows = rd[1:]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
agg = {}
for i, h in stalls:
    agg[h] = sum(int(r[i] or 0) for r in rows)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(rows)), key=lambda k: -int(rows[k][si] or 0))[:top]
for k in sorted(idx):
    r = rows[k]
    best = max(stalls, key=lambda ih: int(r[ih[0]] or 0))
    print(f"{k:5d} {int(r[si]):6d} ({100*int(r[si])/tot:4.1f}%) exec {r[ie]:>8} {best[1]:<18} {r[src].strip()[:90]}")
