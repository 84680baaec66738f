"""Attribute decode-step time to kernel groups by ablation under CUDA-graph replay (results are meaningless when a
kernel is skipped; only the timing is used).  python tools/ablate.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth

sd = synth.make_state_dict(0)
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
feats = torch.randn(928, 1024, generator=torch.Generator().manual_seed(1)).cuda()
T = 64

def run(mask, pdl=1, graph=1):
    eng.set_option("ablate", mask); eng.set_option("pdl", pdl); eng.set_option("cuda_graph", graph)
    for _ in range(2):
        eng.lm_generate(feats, T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        eng.lm_generate(feats, T)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n / (T - 1) * 1e3  # ms per decode step

if "variants" in sys.argv:
    for occ in (5, 7, 6):
        eng.set_option("attn_occ", occ)
        print("full step, attention %d CTAs/SM: %.3f ms" % (occ, run(0)), flush=True)
    for bn in (256, 128, 0):
        eng.set_option("cattn_bn", bn)
        print("full step, c_attn N tile %s: %.3f ms" % (bn or "auto (192)", run(0)), flush=True)
    eng.set_option("dual", 1)
    print("full step, dual halves   : %.3f ms" % run(0), flush=True)
    eng.set_option("dual", 0)
base = run(0)
print("full step                : %.3f ms" % base, flush=True)
import numpy as np
print("grid barrier: %d cycles" % int(eng.debug_read("grid_sync_cycles", (1,), np.int64)[0]), flush=True)
eng.set_option("megakernel", 1 if "mega" in sys.argv else 0)
print("full step, megakernel    : %.3f ms" % run(0), flush=True)
import numpy as np
tr = eng.debug_read("mega_trace", (256,), np.int64).astype(np.float64)
d = np.diff(tr[:172]) / 1.965e3
names = ["LN1+prefetch", "c_attn", "attention", "proj", "LN2", "c_fc", "mproj"]
per = d[:168].reshape(24, 7)
print("megakernel phase durations (us, CTA 0, last step, mean over 24 layers):", {n: round(float(per[:, i].mean()), 2) for i, n in enumerate(names)}, flush=True)
print("  layer 1 phases:", [round(float(v), 2) for v in per[1]], " tail (LNf, lm_head, greedy):", [round(float(v), 2) for v in d[168:171]], flush=True)
eng.set_option("megakernel", 0)
print("full step, PDL off       : %.3f ms" % run(0, pdl=0), flush=True)
print("full step, eager no graph: %.3f ms" % run(0, graph=0), flush=True)
configs = [("attention", 1), ("layernorm", 2), ("c_attn", 4), ("attn_c_proj", 8), ("mlp_c_fc", 16), ("mlp_c_proj", 32),
           ("all GEMMs of the layers", 4 | 8 | 16 | 32), ("everything in the layers", 63)]
if len(sys.argv) > 1:
    configs = [c for c in configs if c[0] in sys.argv[1:]]
for name, mask in configs:
    t = run(mask)
    print("without %-24s: %.3f ms  (saves %.3f ms = %.1f us per layer)" % (name, t, base - t, (base - t) / 24 * 1e3), flush=True)
