"""Debug helper (GPU): decode the bench's cfg 2 / cfg 3 batch once with the in-kernel counters on and print
the step's work / redo statistics, including the diagnostics of the first give-up (capacity retry)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from cases import Built, spec_lexfree, spec_lexicon  # noqa: E402
from flt_backend import FltBackend  # noqa: E402
from text_b200 import synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "lexfree"
B, T, N = int(os.environ.get("DBG_B", 256)), int(os.environ.get("DBG_T", 1000)), 10000
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
em = torch.empty((B, T, N), dtype=torch.float32, device=dev)
for b0 in range(0, B, 32):
    z = torch.randn((min(32, B - b0), T, N), generator=gen, device=dev, dtype=torch.float32)
    em[b0:b0 + z.shape[0]] = torch.log_softmax(z, dim=-1)
G = FltBackend("cuda")
if kind == "lexfree":
    spec = spec_lexfree(N, 50, N, 1e9, sil=0, blank=N - 1)
else:
    sp = synth.lexicon(200000, N, 2, 5, seed=7, exclude=(0, N - 1))
    spec = spec_lexicon(N, 100, N, sp, 1e9, sil=0, blank=N - 1, unk=200000)
b = Built(G, spec)
api, dec = G.api, b.dec
api.set_timing(dec, 2)
api.decode_batch_async(dec, em.data_ptr(), B, T, N)
api.synchronize(dec)
print(json.dumps({"kind": kind, "kernel_ms": api.last_kernel_ms(dec), "work": api.last_stats(dec)}))
