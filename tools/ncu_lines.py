"""Per-source-line profile of one kernel: joins the SASS page of an .ncu-rep with `nvdisasm -g`
line info of the same kernel in libcmpy_b200.so (same build!).

usage: ncu_lines.py file.ncu-rep <mangled-kernel-substring> [top_n]
"""
import csv, collections, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cmpy_b200", "libcmpy_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = None
for i, l in enumerate(dis):
    if l.startswith("_Z") and kern in l and l.rstrip().endswith(":"):
        start = i
        break
assert start is not None, "kernel not found in cubin"
lines = []  # (file, line) per instruction in order
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith("_Z") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h)]
print(f"ncu SASS rows {len(data)}, nvdisasm instructions {len(lines)}")
n = min(len(data), len(lines))
iE, iS = h.index("Instructions Executed"), h.index("# Samples")
iW = h.index("L1 Wavefronts Shared")
agg = collections.defaultdict(lambda: [0, 0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += int(data[k][iE] or 0); a[1] += int(data[k][iS] or 0); a[2] += int(data[k][iW] or 0)
totE = sum(a[0] for a in agg.values()); totS = sum(a[1] for a in agg.values()); totW = sum(a[2] for a in agg.values())
srcs = {}
def text(f, ln):
    if f not in srcs:
        p = os.path.join(ROOT, "cmpy_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = srcs[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
print(f"total executed {totE}, samples {totS}, smem wavefronts {totW}")
print(f"{'file:line':28s} {'exec%':>6s} {'samp%':>6s} {'smemwf%':>7s}  source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f + ':' + str(ln):28s} {a[0] / totE * 100:6.1f} {a[1] / totS * 100:6.1f} {a[2] / max(totW, 1) * 100:7.1f}  {text(f, ln)}")
