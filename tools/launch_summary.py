"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X ...).
usage: python tools/launch_summary.py launches.csv [--last N] [--title "..."]   (--last N: only the final N launches = one step)"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else ""
op = gzip.open if path.endswith(".gz") else open
rows = []
with op(path, "rt", errors="replace") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r["Kernel Name"])).split("::")[-1].strip()
    rows.append((name, ms))
if last:
    rows = rows[-last:]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for n, ms in rows:
    tot[n] += ms
    cnt[n] += 1
total = sum(tot.values())
if title:
    print(title)
print(f"{len(rows)} launches, {total:.2f} ms (cold-cache, serialised per launch: compare SHARES)")
for n, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{n:40s} n={cnt[n]:4d} total={ms:9.3f} ms share={ms / total:.3f}")
