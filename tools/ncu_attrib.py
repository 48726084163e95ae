"""Attribute the executed instructions of one kernel in an .ncu-rep (captured with --import-source on) to CUDA source lines
and SASS opcodes.  The report's source page only lists SASS, so addresses are mapped to lines through `nvdisasm -g` of the
cubin extracted from the object file that was profiled (build with -lineinfo).
    python tools/ncu_attrib.py <file.ncu-rep> <object.o> <mangled-name-substring> [top_n]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, needle = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("_Z") and needle in l and l.rstrip().endswith(":"))
amap, sass, cur = {}, {}, None
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = cur
        sass[int(m.group(1), 16)] = m.group(2)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print("kernel:", rows[0][1][:120])
hdr, data = rows[1], rows[2:]
ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = min(int(r[ia], 16) for r in data if len(r) > ie and r[ia])
byline, samples, mix = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= ie or not r[ie]:
        continue
    a = int(r[ia], 16) - base
    byline[amap.get(a)] += int(r[ie])
    samples[amap.get(a)] += int(r[iss] or 0)
    op = sass.get(a, "?").split()
    op = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")).split(".")[0]
    mix[op] += int(r[ie])
tot, ts = sum(byline.values()), max(1, sum(samples.values()))
print(f"warp instructions executed: {tot}")
print("opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in mix.most_common(14)))
root = os.environ.get("EVW_SRC_DIR") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "evoworld_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().split("\n") for f in os.listdir(root) if f.endswith((".cu", ".cuh", ".h"))}
for k, v in byline.most_common(top):
    if k is None:
        continue
    f, ln = k
    line = src[f][ln - 1].strip()[:105] if f in src and ln <= len(src[f]) else ""
    print(f"{v:10d} {100 * v / tot:5.1f}%  samples {100 * samples[k] / ts:5.1f}%  {f}:{ln}: {line}")
