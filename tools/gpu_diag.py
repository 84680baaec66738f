"""Bring-up diagnostics for the tcgen05 / TMA kernels (run on the GPU box, one step per process so that a trap in one
step cannot poison the next):   python tools/gpu_diag.py <step>"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgrg_b200 import Engine  # noqa: E402


def rnd(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to(torch.bfloat16).cuda()


def report(name, out, ref):
    err = (out - ref).abs()
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    ok = err.max().item() <= tol
    print("%-40s max_err %.4g (tol %.3g) ref_absmax %.3g  %s" % (name, err.max().item(), tol, ref.abs().max().item(),
                                                                "OK" if ok else "MISMATCH"), flush=True)
    if not ok:
        M, N = out.shape[-2], out.shape[-1]
        e2 = err.reshape(-1, N)
        bad_rows = (e2.max(1).values > tol).nonzero().flatten()
        bad_cols = (e2.max(0).values > tol).nonzero().flatten()
        print("   bad rows: %d of %d, first %s" % (len(bad_rows), e2.shape[0], bad_rows[:12].tolist()))
        print("   bad cols: %d of %d, first %s" % (len(bad_cols), N, bad_cols[:12].tolist()))
        print("   out[0,:8] ", out.reshape(-1, N)[0, :8].tolist())
        print("   ref[0,:8] ", ref.reshape(-1, N)[0, :8].tolist())
        print("   nan count", torch.isnan(out).sum().item(), " zeros", (out == 0).sum().item())
    return ok


def gemm_case(e, M, N, K, impl, seed=0):
    A, W = rnd((M, K), seed + 1), rnd((N, K), seed + 2, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    t = time.time()
    out = e.gemm(A, W, bias, 0, impl)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T + bias
    return report("gemm impl=%d M=%d N=%d K=%d (%.1f ms)" % (impl, M, N, K, (time.time() - t) * 1e3), out, ref)


def main():
    step = sys.argv[1]
    e = Engine(0)
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if step == "simt":
        gemm_case(e, 128, 128, 64, 2)
        gemm_case(e, 300, 200, 192, 2)
    elif step == "tc_small":
        # K = 64: one k-block, 4 UMMAs; then identity-structured operands to localise layout errors
        gemm_case(e, 128, 128, 64, 0)
        gemm_case(e, 128, 64, 64, 1)
        A = torch.zeros(128, 64, dtype=torch.bfloat16).cuda()
        A[torch.arange(128), torch.arange(128) % 64] = 1.0
        W = (torch.arange(128 * 64).reshape(128, 64) % 251).to(torch.bfloat16).cuda()
        out = e.gemm(A, W, None, 0, 0)
        report("gemm structured (A=one-hot)", out, A.float() @ W.float().T)
    elif step == "tc_k":
        gemm_case(e, 128, 128, 128, 0)
        gemm_case(e, 128, 128, 512, 0)   # more k-blocks than stages (6): exercises the ring + phase bits
        gemm_case(e, 128, 128, 1024, 1)
    elif step == "tc_shapes":
        for (M, N, K) in [(300, 200, 192), (928, 3072, 1024), (37, 800, 2048), (257, 50257, 1024), (1000, 64, 576),
                          (4096, 1024, 4096)]:
            for impl in (0, 1, 3, 4):
                gemm_case(e, M, N, K, impl)
    elif step == "conv":
        for implicit in (False, True):
            for (B, H, Cin, Cout) in [(2, 16, 64, 64), (1, 32, 128, 128), (2, 16, 2048, 256), (1, 128, 64, 64)]:
                x = rnd((B, H, H, Cin), 6)
                w = rnd((Cout, 3, 3, Cin), 7, 0.05)
                bias = torch.randn(Cout, generator=torch.Generator().manual_seed(8)).cuda()
                out = e.conv3x3(x, w.reshape(Cout, -1), bias, relu=False, implicit=implicit)
                ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias,
                                                 padding=1).permute(0, 2, 3, 1)
                report("conv3x3 implicit=%d B=%d H=%d Cin=%d Cout=%d" % (implicit, B, H, Cin, Cout), out, ref)
    elif step == "gemm_perf":
        cases = [(928, 3072, 1024, i) for i in (0, 1, 3, 4)] + [(928, 1024, 1024, 1), (928, 4096, 1024, 4), (928, 1024, 4096, 1),
                 (928, 50257, 1024, 4)] + [(8192, 8192, 8192, i) for i in (0, 3, 4)] + [(27200, 1024, 8192, 4)]
        for (M, N, K, impl) in cases:
            A, W = rnd((M, K), 1), rnd((N, K), 2, 0.05)
            for _ in range(3):
                out = e.gemm(A, W, None, 0, impl)
            e.set_option("profile", 1)
            for _ in range(10):
                out = e.gemm(A, W, None, 0, impl)
            ms, n = e.profile_read()["test_gemm"]
            e.set_option("profile", 0)
            ms /= n
            print("perf impl=%d M=%d N=%d K=%d: %.4f ms  %.1f TFLOP/s (CUDA events around the kernel, fp32 store epilogue)" %
                  (impl, M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
    elif step == "timeline":
        import numpy as np
        for (M, N, K, bn) in [(928, 3072, 1024, 192), (928, 4096, 1024, 256), (928, 4096, 1024, 128), (928, 1024, 4096, 64),
                              (928, 1024, 1024, 64), (928, 50257, 1024, 256), (8192, 8192, 1024, 256)]:
            for inter in (False, True):
                tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
                ctas = min(tiles, 148)
                ms, tr = e.gemm_bench(M, N, K, bn, 50, inter, ctas)
                tr = tr.astype(np.float64)
                d = lambda i, j: (tr[:, i] - tr[:, j]).mean() / 1.965e3  # us at 1965 MHz
                print("M=%d N=%d K=%d bn=%d ln_interleaved=%d: %.2f us/iter | CTA timeline us (mean): setup %.2f, first-data %.2f, "
                      "mma-issue-done %.2f, epi-start %.2f, epi-done %.2f, exit %.2f (tiles/cta %.1f)" %
                      (M, N, K, bn, inter, ms * 1e3, d(1, 0), d(2, 0), d(3, 0), d(4, 0), d(5, 0), d(6, 0), tiles / ctas), flush=True)
    print("step %s done" % step, flush=True)


if __name__ == "__main__":
    main()
