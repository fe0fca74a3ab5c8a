#!/usr/bin/env python
"""Per-CTA timeline of two consecutive GEMV launches of a chain (BNB_B200_GEMV_PROBE=2): where a launch's period goes.
    BNB_B200_GEMV_PROBE=2 python tools/gemv_trace.py 4096 4096 [n_chain]
Prints, for the last two kernels A -> B of a graph-replayed chain: the spread of CTA entry / exit times, the hand-off
(last exit of A -> first `previous kernel complete` in B), and the phases of the CTA that finishes last."""
import ctypes as ct
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bnb_b200 import functional as F  # noqa: E402

N, K = int(sys.argv[1]), int(sys.argv[2])
chain = int(sys.argv[3]) if len(sys.argv) > 3 else 24
torch.manual_seed(0)
W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
packs = [q.clone() for _ in range(chain)]
x = torch.randn(1, K, device="cuda").bfloat16()
outs = [torch.empty(1, N, dtype=torch.bfloat16, device="cuda") for _ in range(chain)]


def step():
    for i in range(chain):
        F.gemv_4bit(x, packs[i].t(), out=outs[i], state=st)


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
buf = (ct.c_ulonglong * (2 * 320 * 8))()
F.lib.cbnb_debug_gemv_trace(buf)
t = np.array(buf, dtype=np.uint64).reshape(2, 320, 8).astype(np.int64)
live = [t[s_][t[s_][:, 0] > 0] for s_ in range(2)]
if not len(live[0]) or not len(live[1]):
    sys.exit("no trace (BNB_B200_GEMV_PROBE=2?)")
a, b = (live[0], live[1]) if live[0][:, 0].min() < live[1][:, 0].min() else (live[1], live[0])
t0 = a[:, 0].min()
rel = lambda v: (v - t0) / 1e3   # noqa: E731
names = ["entry", "loads_issued", "prev_done", "x_ready", "all_done", "exit"]
out = {"shape": [N, K], "ctas": [int(len(a)), int(len(b))]}
for tag, k in (("A", a), ("B", b)):
    out[tag] = {n: {"min": round(float(rel(k[:, i]).min()), 2), "median": round(float(np.median(rel(k[:, i]))), 2),
                    "max": round(float(rel(k[:, i]).max()), 2)} for i, n in enumerate(names)}
    last = k[np.argmax(k[:, 5])]
    out[tag]["last_cta"] = {n: round(float(rel(last[i])), 2) for i, n in enumerate(names)}
    out[tag]["last_cta"].update(sm=int(last[6]), tiles=int(last[7]))
    out[tag]["tiles_hist"] = {int(v): int(c) for v, c in zip(*np.unique(k[:, 7], return_counts=True))}
out["period_us"] = round(float(rel(b[:, 5].max()) - rel(a[:, 5].max())), 2)
out["handoff_lastexitA_to_first_prevdoneB_us"] = round(float(rel(b[:, 2].min()) - rel(a[:, 5].max())), 2)
# CTAs of B that entered after A's last exit (nothing of their prologue overlapped A)
late = b[b[:, 0] >= a[:, 5].max() - 200]
out["B_ctas_entering_within_0.2us_of_A_end_or_later"] = int(len(late))
print(json.dumps(out))
