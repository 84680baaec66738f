"""Attribute decode-step time to kernel groups by ablation under CUDA-graph replay (results are meaningless when a
kernel is skipped; only the timing is used), and compare the decode-step variants.  python tools/ablate.py [variants]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth

sd = synth.make_partial_state_dict(0, ("detector", "heads", "lm"))  # uncalibrated detector: only the decoder is timed
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
ROWS = int(os.environ.get("ROWS", "928"))
T = int(os.environ.get("T", "64"))
feats = torch.randn(ROWS, 1024, generator=torch.Generator().manual_seed(1)).cuda()


def run(mask=0, **opts):
    eng.set_option("ablate", mask)
    for k, v in opts.items():
        eng.set_option(k, v)
    for _ in range(2):
        eng.lm_generate(feats, T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        eng.lm_generate(feats, T)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n / (T - 1) * 1e3  # ms per decode step


print("rows %d, T %d" % (ROWS, T), flush=True)
import numpy as np
print("max active 16-CTA clusters:", int(eng.debug_read("cluster16_max_active", (1,), np.int32)[0]), flush=True)
eng.set_option("ln_head", 0)
base = run(0)
print("full step (fused attention, CTA-pair projections): %.3f ms" % base, flush=True)
print("  gemm_2cta=0                       : %.3f ms" % run(0, gemm_2cta=0), flush=True)
eng.set_option("gemm_2cta", 1)
print("  attn_balance=0 (128-row tiles)    : %.3f ms" % run(0, attn_balance=0), flush=True)
eng.set_option("attn_balance", 1)
print("  attn_early=1                      : %.3f ms" % run(0, attn_early=1), flush=True)
eng.set_option("attn_early", 0)
print("  attn_mc=1 (A multicast, head pairs): %.3f ms" % run(0, attn_mc=1), flush=True)
print("  attn_mc=0                         : %.3f ms" % run(0, attn_mc=0), flush=True)
print("  epi_tma=0 (register epilogue)     : %.3f ms" % run(0, epi_tma=0), flush=True)
eng.set_option("epi_tma", 1)
if "variants" in sys.argv:
    print("  dual=1 (two halves out of phase)  : %.3f ms" % run(0, dual=1), flush=True)
    eng.set_option("dual", 0)
if "variants" in sys.argv:
    for aw, sl in ((8, 4), (24, 1), (16, 2)):
        print("  attn_warps=%d attn_slots=%d        : %.3f ms" % (aw, sl, run(0, attn_warps=aw, attn_slots=sl)), flush=True)
    print("  l2_ahead=2 (16 warps x 2 slots)   : %.3f ms" % run(0, l2_ahead=2), flush=True)
    eng.set_option("l2_ahead", 0)
    print("  fused_attn=0 (round 1 step)       : %.3f ms" % run(0, fused_attn=0), flush=True)
    eng.set_option("fused_attn", 1)
    print("  ln_head=1 (group barrier)         : %.3f ms" % run(0, ln_head=1), flush=True)
    print("  ln_head=1 fused_attn=0            : %.3f ms" % run(0, fused_attn=0), flush=True)
    eng.set_option("fused_attn", 1)
    eng.set_option("ln_head", 0)
    print("  PDL off                           : %.3f ms" % run(0, pdl=0), flush=True)
    print("  eager, no graph                   : %.3f ms" % run(0, pdl=1, cuda_graph=0), flush=True)
    eng.set_option("cuda_graph", 1)
configs = [("attn_fused", 64), ("layernorm", 2), ("attn_c_proj", 8), ("mlp_c_fc", 16), ("mlp_c_proj", 32),
           ("everything in the layers", 64 | 2 | 8 | 16 | 32)]
for name, mask in configs:
    t = run(mask)
    print("without %-24s: %.3f ms  (saves %.3f ms = %.1f us per layer)" % (name, t, base - t, (base - t) / 24 * 1e3), flush=True)
eng.set_option("ablate", 0)
