"""Top stall sites of an .ncu-rep source page (SASS level): python tools/ncu_source.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be concatenated; take the first block
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2: break
    body.append(r)
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for i in stall_cols:
    agg[hdr[i]] = sum(int(r[i] or 0) for r in body)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(body)), key=lambda j: -int(body[j][iS] or 0))[:top]
for j in sorted(idx):
    r = body[j]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{j:5d} {int(r[iS]):7d} {100*int(r[iS])/tot:5.1f}% ex={r[iEx]:>10s} {r[iSrc].strip()[:70]:70s} {st}")
