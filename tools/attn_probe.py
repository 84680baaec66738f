import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, synth
sd = synth.make_state_dict(0)
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(torch.device("cuda", 0)); m.eval()
eng = m._engine()
eng.set_option("cuda_graph", 0)
feats = torch.randn(928, 1024, generator=torch.Generator().manual_seed(1)).cuda()
eng.lm_generate(feats, 64)
torch.cuda.synchronize()
print("done")
