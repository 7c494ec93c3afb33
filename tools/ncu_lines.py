#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: per CUDA source line, share of executed
warp instructions and of stall samples.   python tools/ncu_lines.py rep.ncu-rep obj.o kernel_substr [topN]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, obj, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# collect (line) per instruction of the wanted function, in order
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith(".text."):
        infn = sub in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels may be in the report: take the table whose kernel name matches
tables, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            j += 1
        tables.append((name, hdr, rows[i + 2:j])); i = j
    else:
        i += 1
m_ = re.search(r"([a-z0-9_]+_kernel)", sub)
short = m_.group(1) if m_ else sub
name, hdr, body = ([t for t in tables if short in t[0]] or tables)[0]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0])
sagg = collections.defaultdict(lambda: collections.Counter())
n = min(len(body), len(lines))
for k in range(n):
    try:
        agg[lines[k]][0] += int(body[k][ie]); agg[lines[k]][1] += int(body[k][ss] or 0)
        for ci, nm in stall_cols:
            if body[k][ci]:
                sagg[lines[k]][nm] += int(body[k][ci])
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{name[:80]}  sass instrs {len(body)} / disasm {len(lines)}; executed {ti}, samples {ts}")
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:100] if ln - 1 < len(src[f]) else ""
    why = ", ".join(f"{n}={c}" for n, c in sagg[(f, ln)].most_common(4))
    print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / max(ts, 1):5.1f}% smp  {f}:{ln:<4d} {text[:70]:70s} | {why}")
tot = collections.Counter()
for c in sagg.values():
    tot.update(c)
print("all samples by reason:", ", ".join(f"{n}={c}" for n, c in tot.most_common(10)))
