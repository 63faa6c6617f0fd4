#!/usr/bin/env python
"""Per-kernel SASS opcode census of the shipped library (cuobjdump -sass text_b200/lib/libflt_decoder.so):
instruction count, registers / shared memory from the ELF resource usage, and the opcode families that tell
what a kernel is made of (TMA bulk copies + mbarriers, shared / global / local memory traffic, fp64 adds,
FMA contraction, atomics, barriers). usage: tools/sass_census.py > profiles/rNN_sass_census.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "text_b200", "lib", "libflt_decoder.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for l in res.splitlines():
    m = re.search(r"Function (\S+):", l)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*SHARED:(\d+).*LOCAL:(\d+)", l)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
fams = [("UBLKCP (TMA bulk copy)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("BAR (CTA barrier)", r"^BAR"),
        ("LDS", r"^LDS"), ("STS", r"^STS"), ("ATOMS (smem atomics)", r"^ATOMS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
        ("ATOMG/RED", r"^(ATOMG|RED)"), ("LDL/STL (spills)", r"^(LDL|STL)"), ("LDC", r"^LDC"), ("DADD/DSETP", r"^(DADD|DSETP)"),
        ("DFMA/DMUL", r"^(DFMA|DMUL)"), ("FFMA", r"^FFMA"), ("IMAD", r"^IMAD"), ("SHFL", r"^SHFL"), ("REDUX/VOTE/MATCH", r"^(REDUX|VOTE|MATCH)"),
        ("HMMA/UTC* (tensor cores)", r"^(HMMA|IMMA|DMMA|UTC)"), ("BRA", r"^BRA")]
kern, name = collections.OrderedDict(), None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and name:
        op = m.group(1)
        kern[name]["_total"] += 1
        for label, rx in fams:
            if re.match(rx, op):
                kern[name][label] += 1
arch = re.findall(r"arch = (sm_\w+)", sass)
print(f"# SASS census of {os.path.relpath(so, ROOT)}\n")
print(f"cubins: {sorted(set(arch))} (one target, no PTX fallback needed on B200); built with `-fmad=false`: FFMA / DFMA only where the source")
print("calls fma / exp / log1p explicitly (the `*_wide` log-add kernels).\n")
names = list(kern)
short = [subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.split("(")[0].replace("void ", "").strip() for n in names]
print("| kernel | SASS instr | regs | static smem | local (spill) B | " + " | ".join(f for f, _ in fams) + " |")
print("|---|---|---|---|---|" + "---|" * len(fams))
for n, s in zip(names, short):
    u = usage.get(n, ("", "", ""))
    print(f"| `{s}` | {kern[n]['_total']} | {u[0]} | {u[1]} | {u[2]} | " + " | ".join(str(kern[n][f]) for f, _ in fams) + " |")
