#!/usr/bin/env python
"""Copy a tools/r2_snap.sh result (gpurun_out/<tag>) into profiles/r02_<tag>_* and refresh profiles/traffic.json
(DRAM bytes per launch of the dominant kernel from the ncu --set full capture, tied to the kernel sources' hash).
usage: tools/r2_profiles.py <tag> [name used in profiles/, default = tag]"""
import json, os, shutil, subprocess, sys
tag = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else tag
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
for f in ("bench.json", "bench_reference.json", "topm.jsonl", "launches.csv", "pytest_gpu.txt", "smoke.txt", "smi.txt"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f"r02_{name}_{f}"))
rep = os.path.join(src, "prof.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                          "flt_k_fused, cfg 2 full size (B=256, T=1000, N=10000, beam=50, bst=N), ncu --set full --clock-control none"],
                         capture_output=True, text=True).stdout
    body, _, tr = out.partition("TRAFFIC ")
    open(os.path.join(dst, f"r02_{name}_ncu_fused.md"), "w").write(body)
    tr = json.loads(tr)
    import bench
    b = json.loads(open(os.path.join(src, "bench.json")).read().strip().splitlines()[-1])
    json.dump({"flt_k_fused": {"workload": b["config"]["workload"], "dram_bytes_per_launch": tr["flt_k_fused"]["dram_bytes"],
                               "algorithmic_bytes_per_launch": b["roofline"]["algorithmic_bytes_per_launch"],
                               "src_sha": bench.src_hash(), "file": f"profiles/r02_{name}_ncu_fused.md",
                               "source": f"ncu --set full, one launch at the benchmark's full size, snapshot {tag}"}},
              open(os.path.join(dst, "traffic.json"), "w"), indent=1)
# source-level hot lines (needs the local library to be the build that ran)
for rep_name, kern, out in (("prof.ncu-rep", "flt_k_fused", "ncu_lines_fused"), ("prof_lexicon.ncu-rep", "flt_k_decode512", "ncu_lines_lexicon_step")):
    rp = os.path.join(src, rep_name)
    if os.path.exists(rp):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rp, kern, "60"], capture_output=True, text=True).stdout
        open(os.path.join(dst, f"r02_{name}_{out}.txt"), "w").write(txt)
print(open(os.path.join(dst, f"r02_{name}_ncu_fused.md")).read())
