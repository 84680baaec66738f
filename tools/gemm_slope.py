"""Per-k-block cost and fixed cost of the projection kernels at decode shapes: time(K) for back-to-back launches
(rgrg_gemm_bench, no PDL, warm L2), 1-CTA kernel (N tile 256) vs CTA-pair kernel (bn = 512)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import Engine
e = Engine(0)
for M, N in ((928, 4096), (928, 1024)):
    for bn in (256, 512):
        ts = []
        for K in (256, 1024, 2048, 4096):
            ms, _ = e.gemm_bench(M, N, K, bn, iters=200)
            ts.append((K // 64, ms * 1e3))
        slope = (ts[-1][1] - ts[1][1]) / (ts[-1][0] - ts[1][0])
        print("M=%d N=%d %s: " % (M, N, "CTA-pair" if bn == 512 else "1-CTA   ") + "  ".join("%2d kb: %6.2f us" % t for t in ts) +
              "   -> %.3f us per k-block, fixed ~%.1f us" % (slope, ts[1][1] - 16 * slope), flush=True)
