"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA = UTCHMMA / UTCQMMA, tensor-memory
loads / stores = LDTM / STTM, TMA = UTMALDG / UTMASTG, tcgen05.commit = UTCBAR, elect.sync = ELECT, 64-bit atomicMin =
REDG/ATOMG ... MIN.64) in the built library.  usage: python tools/sass_extract.py [lib.so] > profiles/rNN_sass_extract.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "evoworld_b200/_lib/libevoworld_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "ELECT", "SYNCS", "REDG", "ATOMG", "HMMA", "MUFU.EX2", "REDUX"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
        if op.startswith(("REDG", "ATOMG")) and "MIN.64" in op:
            counts[cur]["RED/ATOM.MIN.64"] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass), kernels with none of the listed mnemonics omitted")
for (name, c), dn in zip(counts.items(), demangle if len(demangle) == len(counts) else list(counts)):
    keys = [k for k in c if k != "_total"]
    if not keys:
        continue
    short = re.sub(r"\((int|bool|unsigned int)\)", "", dn)
    short = re.sub(r"\(.*", "", short)
    print(f"{short[:110]:110s} instr {c['_total']:6d}  " + "  ".join(f"{k} {c[k]}" for k in sorted(keys)))
