#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info of
the shipped cubin, and print the hottest source lines (instructions executed, stall samples).
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "text_b200", "lib", "libflt_decoder.so")
tmp = tempfile.mkdtemp()
# one cubin per translation unit; the seven objects of kern_step.cu carry cubins of the same name, so each
# object file (text_b200/build/*.o, else the library) is unpacked into a directory of its own
build = os.path.join(root, "text_b200", "build")
units = sorted(os.path.join(build, f) for f in os.listdir(build) if f.endswith(".o")) if os.path.isdir(build) else [so]
dis = ""
for k, unit in enumerate(units):
    d = os.path.join(tmp, str(k))
    os.makedirs(d)
    subprocess.run(["cuobjdump", "-xelf", "all", unit], cwd=d, capture_output=True)
    for cb in sorted(os.listdir(d)):
        if cb.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cb)], capture_output=True, text=True).stdout
# offset -> (file,line) for the kernel
line_of, cur, inside = {}, None, False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = re.search(re.escape(kern) + r"(N3flt|[^A-Za-z0-9_])", l) is not None  # not flt_k_x_wide for flt_k_x
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
# multiple kernels may be present: take the first block whose name matches
blocks, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j])
            j += 1
        blocks.append((name, hdr, body))
        i = j
    else:
        i += 1
name, hdr, body = next(b for b in blocks if kern in b[0])
ci = {h: k for k, h in enumerate(hdr)}
base = int(body[0][0], 16)
agg = {}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = tot_s = 0
for r in body:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    inst = float(r[ci["Instructions Executed"]] or 0)
    samp = float(r[ci["# Samples"]] or 0)
    a = agg.setdefault(key, [0.0, 0.0, {}])
    a[0] += inst
    a[1] += samp
    for s in stall_cols:
        v = float(r[ci[s]] or 0)
        if v:
            a[2][s] = a[2].get(s, 0) + v
    tot_i += inst
    tot_s += samp
src_cache = {}


def src(f, n):
    p = os.path.join(root, "text_b200", "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[p]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""


print(f"kernel {name}: {tot_i:.0f} warp-instructions, {tot_s:.0f} samples")
for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -(kv[1][0] if os.environ.get("BY_INST") else kv[1][1]))[:top]:
    tops = ",".join(f"{k[6:]}:{v / max(samp, 1) * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{samp / tot_s * 100:5.1f}% samp {inst / tot_i * 100:5.1f}% inst  {key[0]}:{key[1]:<4} [{tops}] {src(*key)}")
