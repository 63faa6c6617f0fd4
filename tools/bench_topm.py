"""Roofline of the token-beam select alone (flt_topm_rows): GB/s of 4*rows*N + 8*rows*M algorithmic bytes,
CUDA events on the launch stream, inputs (10 GB) far larger than L2. usage: tools/bench_topm.py [M ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from text_b200 import capi  # noqa: E402

api = capi.Api()
rows, N = 256000, 10000
Ms = [int(x) for x in sys.argv[1:]] or [53, 105, 205]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
em = torch.empty((rows, N), dtype=torch.float32, device=dev)
for r0 in range(0, rows, 32000):
    z = torch.randn((32000, N), generator=gen, device=dev, dtype=torch.float32)
    em[r0:r0 + 32000] = torch.log_softmax(z, dim=-1)
del z
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
for M in Ms:
    tok = torch.empty((rows, M), dtype=torch.int32, device=dev)
    val = torch.empty((rows, M), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream()
    for _ in range(3):
        api.topm_rows(em.data_ptr(), rows, N, M, tok.data_ptr(), val.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        api.topm_rows(em.data_ptr(), rows, N, M, tok.data_ptr(), val.data_ptr(), s.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    b = rows * (4 * N + 8 * M)
    print(json.dumps({"kernel": "token_select", "N": N, "M": M, "rows": rows, "ms": ms, "algorithmic_bytes": b,
                      "achieved_gbs": b / ms / 1e6, "peak_gbs": peak, "frac": b / ms / 1e6 / peak,
                      "stream_kernel": os.environ.get("FLT_NO_STREAM") is None}))
