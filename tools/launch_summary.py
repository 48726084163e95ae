"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X ...).
usage: python tools/launch_summary.py launches.csv [--last N | --step-marker KERNEL] [--title "..."]\n(--last N: only the final N launches; --step-marker K: the last complete span between two launches of kernel K = one step)"""
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
    name = r["Kernel Name"].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", name)).split("::")[-1].strip()
    rows.append((name, ms))
if last:
    rows = rows[-last:]
if "--step-marker" in sys.argv:  # the last complete span between two launches of this kernel = one step
    mk = sys.argv[sys.argv.index("--step-marker") + 1]
    idx = [i for i, (n, _) in enumerate(rows) if n == mk]
    if len(idx) >= 2:
        rows = rows[idx[-2]:idx[-1]]
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
