"""One generate() at BASELINE configs[1] geometry (batch 32, 512x512, 29 regions) with a short decode, eager launches —
the target of `ncu --kernel-id :::1` (first invocation of every kernel).  env: T (default 4), B (default 32)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth
sd = synth.make_state_dict(0)
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
eng.set_option("cuda_graph", 0)
imgs = synth.synthetic_images(int(os.environ.get("B", "32")), 512, seed=1000).cuda()
out = eng.generate(imgs, int(os.environ.get("T", "4")))
torch.cuda.synchronize()
print("done R=%d" % out["R"])
