#!/usr/bin/env python
"""Per source line: warp-instructions per frame per warp (ncu source page joined with nvdisasm line
info), in source order. usage: tools/ncu_perframe.py <rep> <kernel-substr> <frames*warps> [file]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, denom = sys.argv[1], sys.argv[2], float(sys.argv[3])
only = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "text_b200", "lib", "libflt_decoder.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = kern in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
i = 0
blocks = []
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        blocks.append((name, hdr, body)); i = j
    else:
        i += 1
name, hdr, body = next(b for b in blocks if kern in b[0])
ci = {h: k for k, h in enumerate(hdr)}
base = int(body[0][0], 16)
agg = {}
for r in body:
    key = line_of.get(int(r[0], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += float(r[ci["Instructions Executed"]] or 0)
    a[1] += float(r[ci["# Samples"]] or 0)
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total {tot/denom:.0f} warp-instr per frame per warp; {ts:.0f} samples")
byfile = {}
for (f, n), (inst, samp) in agg.items():
    byfile.setdefault(f, [0, 0]); byfile[f][0] += inst; byfile[f][1] += samp
for f, (inst, samp) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f}: {inst/denom:.0f} instr, {samp/ts*100:.1f}% samples")
for (f, n), (inst, samp) in sorted(agg.items()):
    if only and f != only: continue
    if inst / denom < 3 and samp / ts < 0.004: continue
    src = ""
    p = os.path.join(root, "text_b200", "csrc", f)
    if os.path.exists(p):
        L = open(p).read().splitlines()
        src = L[n - 1].strip()[:80] if 0 < n <= len(L) else ""
    print(f"{f}:{n:<5} {inst/denom:7.1f} instr {samp/ts*100:5.1f}% | {src}")
