"""Larger-configuration sanity on the GPU box: BASELINE.json configs 3-5 in miniature (1024x1024 inputs, beam search at
hundreds of rows, max_length 128).  Prints timings; asserts only structural invariants."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth

sd = synth.make_state_dict(0)
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()

def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t

# config 3 flavour: max_length 128
imgs = synth.synthetic_images(8, 512, seed=1003).cuda()
out, dt = timed(lambda: eng.generate(imgs, 128))
print("B=8 S=512 T=128 greedy: R=%d ids %s %.3f s" % (out["R"], out["ids"].shape, dt), flush=True)
assert out["ids"].shape == (out["R"], 128)

# config 4 flavour: beam search, 4 beams, 16 images, T=32 (rows = R*4)
imgs = synth.synthetic_images(16, 512, seed=1004).cuda()
out, dt = timed(lambda: eng.generate(imgs, 32, num_beams=4, early_stopping=True))
print("B=16 S=512 T=32 beams=4: R=%d (rows %d) ids %s %.3f s" % (out["R"], out["R"] * 4, out["ids"].shape, dt), flush=True)
assert out["ids"].shape[0] == out["R"] and out["ids"].shape[1] <= 32 and (out["ids"][:, 0] == 50256).all()

# config 5 flavour: 1024x1024 inputs
imgs = synth.synthetic_images(4, 1024, seed=1005).cuda()
det, dt = timed(lambda: eng.detect(imgs))
print("B=4 S=1024 detect: proposals %s detected/img %s selected/img %s %.3f s" % (det["num_proposals"].tolist(),
      det["detected"].sum(1).tolist(), det["selected"].sum(1).tolist(), dt), flush=True)
out, dt = timed(lambda: eng.generate(imgs, 16))
print("B=4 S=1024 T=16 greedy: R=%d ids %s %.3f s" % (out["R"], out["ids"].shape, dt), flush=True)
print("launches", eng.kernel_launches)
