"""torchrun --nproc-per-node 2 tools/comm_check.py — the engine-side result gather (rgrg_allgather_results: device-side pack
+ one ncclAllGather) against the torch.distributed gather of host-packed blobs: identical merged results, greedy and beam."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import ReportGenerationModel, parallel, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sd = synth.make_state_dict(0) if rank == 0 else None
dist.barrier()
if rank != 0:
    sd = synth.make_state_dict(0)
m = ReportGenerationModel(True); m.load_state_dict(sd); m.to(dev); m.eval()
eng = m._engine()
parallel.init_engine_comm(eng, device=dev)
B = 3
imgs = synth.synthetic_images(B, 512, seed=500 + rank).to(dev)
for T, nb in ((9, 1), (8, 4)):
    out = eng.generate(imgs, T, nb, nb > 1)
    a = parallel.all_gather_results_native(eng, B, T)
    b = parallel.all_gather_results(out, B, T, device=dev)
    for k in ("ids", "selected", "detected", "boxes", "scores"):
        assert np.array_equal(a[k], b[k]), (T, nb, k)
    assert a["R"] == b["R"] and a["ids"].shape[0] == a["R"]
    if rank == 0:
        print("T=%d beams=%d: native gather == torch gather, R=%d rows from %d ranks" % (T, nb, a["R"], world), flush=True)
dist.barrier()
dist.destroy_process_group()
