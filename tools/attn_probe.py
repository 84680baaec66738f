"""One eager lm_generate (928 rows by default) — the target of `ncu -k regex:...` captures of decode kernels.
env: ROWS, T, OPTS="key=value,key=value" (engine options)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth
sd = synth.make_partial_state_dict(0, ("detector", "heads", "lm"))
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
eng.set_option("cuda_graph", 0)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("=")
    eng.set_option(k, int(v))
feats = torch.randn(int(os.environ.get("ROWS", "928")), 1024, generator=torch.Generator().manual_seed(1)).cuda()
eng.lm_generate(feats, int(os.environ.get("T", "64")))
torch.cuda.synchronize()
print("done")
