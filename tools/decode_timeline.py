"""Timeline of one decode step under CUDA-graph replay: %globaltimer stamps of the first and the last CTA of every kernel of
the layers (engine option "trace").  python tools/decode_timeline.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth

sd = synth.make_partial_state_dict(0, ("detector", "heads", "lm"))
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
feats = torch.randn(928, 1024, generator=torch.Generator().manual_seed(1)).cuda()
T = int(os.environ.get("T", "48"))
eng.set_option("trace", 1)
eng.lm_generate(feats, T)
eng.lm_generate(feats, T)
tr = eng.debug_read("decode_trace", (256, 2, 8), np.int64).astype(np.float64)
# slot order per layer: view_attn takes the fused kernel's slot before it launches LayerNorm 1
names = ["ATTN", "LN1", "CPROJ", "LN2", "CFC", "MPROJ"]
ev = ["entry", "setup", "pred-done", "mma0", "mma-issued", "acc-ready", "body-done", "exit"]
# the fused attention kernel uses slots 3 / 4 for "tiles written" (epilogue done) and "first K/V chunk landed"

for l in (11, 12):
    t0 = tr[6 * l, 0, 0]
    print("layer %d (us since the fused attention kernel's first CTA entered; first CTA | last CTA)" % l)
    for k, nm in enumerate(names):
        r = tr[6 * l + k]
        def fmt(c):
            return " ".join("%s=%6.2f" % (ev[i], (r[c, i] - t0) / 1e3) for i in range(8) if r[c, i] > 0)
        print("  %-6s first: %s" % (nm, fmt(0)))
        print("  %-6s last : %s" % ("", fmt(1)))
    nxt = tr[6 * (l + 1), 0, 0]
    print("  next layer's attention entry: %.2f us" % ((nxt - t0) / 1e3))
# averages over layers 2..22: duration from a kernel's first entry to its last exit, and the gap to the next kernel's pred-done
dur = {nm: [] for nm in names}
for l in range(2, 22):
    for k, nm in enumerate(names):
        r = tr[6 * l + k]
        dur[nm].append((max(r[0, 7], r[1, 7]) - min(r[0, 0], r[1, 0])) / 1e3)
print("mean entry->exit (us):", {nm: round(float(np.mean(v)), 2) for nm, v in dur.items()})
per_layer = (tr[6 * 22, 0, 0] - tr[6 * 2, 0, 0]) / 20 / 1e3
print("layer period: %.2f us" % per_layer)
# phases of the GEMM kernels, first CTA (a full 128-row tile): predecessor done -> first MMA, main loop, epilogue, exit
for nm in ("CPROJ", "CFC", "MPROJ"):
    k = names.index(nm)
    rows = np.stack([tr[6 * l + k, 0] for l in range(2, 22)])
    print("%-6s wait->mma0 %.2f  main loop %.2f  epilogue %.2f  exit %.2f us" % (
        nm, np.mean(rows[:, 3] - rows[:, 2]) / 1e3, np.mean(rows[:, 5] - rows[:, 3]) / 1e3,
        np.mean(rows[:, 6] - rows[:, 5]) / 1e3, np.mean(rows[:, 7] - rows[:, 6]) / 1e3))
