#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel of an .ncu-rep (test infrastructure).

    python tools/ncu_lines.py report.ncu-rep <kernel regex> <object.cubin or lib.so> <mangled-name substring> [top]

Joins ncu's SASS page (instructions executed / samples per SASS address) with nvdisasm's line table
(-g: file + line per instruction, outermost "inlined at" frame of our own files last)."""
import csv, io, re, subprocess, sys, collections, os, tempfile, glob

rep, kern, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
if obj.endswith(".so"):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubins = glob.glob(d + "/*.cubin")
else:
    cubins = [obj]
dis = None
for c in cubins:
    out = subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
    if re.search(r"\.text\.[^\n]*" + re.escape(mangled), out):
        dis = out
        break
assert dis is not None, "kernel not found in " + obj
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith("//---") and ".text." in l and mangled in l)
off2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(data[0][0], 16)
agg = collections.OrderedDict()
tot_i = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    loc, _ = off2line.get(off, (None, ""))
    key = (loc[0], loc[1]) if loc else ("?", 0)
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += n; a[1] += s; a[2] += 1
    tot_i += n; tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}, SASS lines {len(data)}")
src_cache = {}
def src(fn, ln):
    for root in ("splat_one_b200/csrc", "."):
        p = os.path.join(root, fn)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[p]) else ""
    return ""
for (fn, ln), (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  {k:4d} sass  {fn}:{ln:<4d} {src(fn, ln)}")
