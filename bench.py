#!/usr/bin/env python
"""bench.py — reports/s of the region-guided report-generation path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--max-length T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (ReportGenerationModel.generate, greedy, 29 regions) over one batch of
synthetic 512x512 images: BASELINE.json configs[1] (batch 32 / GPU, max_length 64).  Images shard across ranks with no
data-path collective (weak scaling); the one NCCL all-gather that collects the finished token buffers is inside the step.
Prints ONE JSON line on rank 0.

  value       device-resident inputs, CUDA-event timing, max over ranks
  e2e         the same metric through the reference-facing call (`ReportGenerationModel.generate(images)`) with
              pinned HOST images: H2D of the batch and D2H of ids / masks / boxes inside the timed region
  roofline    the dominant kernel category, timed live with CUDA events on the launch stream (engine "profile"
              option) against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
              the reference's own CPU implementation of the path on the box's host cores: the UNMODIFIED reference
              files (oracle/_ref/, kind "reference") when the build container has installed them, else the oracle
              port (kind "port"); a bounded sample of the workload (1-4 images per step, same T, full generate())
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}  # B200_PROFILING.md


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            p["_source"] = "measured"
            return p
        except Exception:
            pass
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback"
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench %s] %s" % (time.strftime("%H:%M:%S"), msg), file=sys.stderr, flush=True)


def load_cpu_reference(sd):
    """The CPU implementation of the path that the CPU legs time: the UNMODIFIED reference (oracle/_ref/, copied by
    oracle/install_ref.py in the build container; `kind: "reference"`) or, when that is absent, the oracle port
    (`kind: "port"`).  Returns (kind, generate(images, max_length) -> output, threads)."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: undo it for the CPU arm
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_harness

        if not ref_harness.reference_available():
            raise RuntimeError("no reference checkout")
        model = ref_harness.build_reference_model(None)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected[:3]
        model.eval()

        def gen(images, T):
            with torch.no_grad():
                return model.generate(images, max_length=T)

        return "reference", gen, torch.get_num_threads()
    except Exception as exc:  # noqa: BLE001 - any import / build problem -> the port, stated in the output
        log("reference not usable (%s): timing the oracle port" % (str(exc).splitlines()[0] if str(exc) else type(exc).__name__))
        import rgrg_oracle as O

        def gen(images, T):
            return O.generate(sd, images, max_length=T)

        return "port", gen, torch.get_num_threads()


def timed_cpu_generate(gen, images, T):
    t0 = time.perf_counter()
    out = gen(images, T)
    dt = time.perf_counter() - t0
    R = 0 if isinstance(out, int) else int(out[0].shape[0])
    return dt, R


def flops_per_image(S, P, R, T):
    """SURVEY.md §8(d) algorithmic work (MACs x 2)."""
    s2 = (S / 512.0) ** 2
    macs = 20.942e9 * s2 + 10.083e9 * s2 + 135.42e6 * P + 77.9e6 + R * (52.4e6 + (T - 1) * 353.45e6)
    return 2.0 * macs


# algorithmic FLOPs of one launch of each GEMM category, as a function of the run's geometry
def category_flops(cat, rows, P_total, B, S):
    f = S // 32
    d = 1024
    table = {
        "lm_head": 2.0 * rows * 50257 * d,
        "c_attn": 2.0 * rows * 3 * d * d,
        "attn_c_proj": 2.0 * rows * d * d,
        "mlp_c_fc": 2.0 * rows * 4 * d * d,
        "mlp_c_proj": 2.0 * rows * 4 * d * d,
        "fc6": 2.0 * P_total * 1024 * 131072,
        "fc7": 2.0 * P_total * 1024 * 1024,
        "rpn_conv": 2.0 * B * f * f * 2048 * 9 * 2048,
        "lm_image_kv": 2.0 * rows * 49152 * d,
    }
    return table.get(cat)


def category_bytes(cat, rows, mean_L):
    """algorithmic HBM bytes of one launch of the HBM-bound categories."""
    if cat == "attention":
        return rows * mean_L * 2 * 1024 * 2.0  # K and V rows of one layer, bf16
    if cat == "attn_fused":
        # fused c_attn + KV append + attention: cached K / V rows read (mean_L - 1 keys) + the appended k, v + q-side
        # operands (x read, attention output written) + the layer's c_attn weight
        return rows * (mean_L - 1) * 4096.0 + rows * 4096.0 + rows * 1024 * 2 * 2.0 + 3072 * 1024 * 2.0
    if cat == "layernorm":
        return rows * 1024 * (4 + 2.0)
    return None


def ncu_traffic(cat):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the category's kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json records the figure, the launch it was taken on and the report)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(cat)
    except Exception:
        return None


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores; rank 0 only.  Every step
    is one full `generate()` (no time-boxing, no extrapolation) on a bounded sample of the workload: `b` images per
    step, with b chosen after the first warm-up step so that the whole run stays within a few minutes."""
    if rank != 0:
        return
    from rgrg_b200 import synth

    sd = synth.make_state_dict(0)
    kind, gen, cores = load_cpu_reference(sd)
    T = args.max_length
    n_steps = args.warmup + args.steps
    budget = 360.0 / max(1, n_steps)  # seconds per step
    b = 1
    times = []
    R_seen = 0
    for step in range(n_steps):
        img = synth.synthetic_images(b, args.image_size, seed=2000 + step)
        dt, R_seen = timed_cpu_generate(gen, img, T)
        log("reference step %d: %.1f s (%d image(s), R=%d rows)" % (step, dt, b, R_seen))
        if step >= args.warmup:
            times.append(dt)
        elif step == 0 and args.warmup >= 1:
            b = int(max(1, min(4, budget // max(dt, 1e-3))))  # fixed from here on: every timed step uses the same b
    total = sum(times)
    value = len(times) * b / total
    what = "the UNMODIFIED reference (oracle/_ref)" if kind == "reference" else "fp32 oracle port of the reference algorithm"
    sample = "%d image(s) per step (of the %d-image workload batch), %dx%d, greedy max_length=%d, %s, fp32 CPU, full generate() per step; last step R=%d rows" % (
        b, args.batch, args.image_size, args.image_size, T, what, R_seen)
    cfg = workload_config(args, args.batch, world)
    cfg["timed_batch_per_step"] = b  # what this arm actually times per step; `workload` names the GPU arm's configuration
    line = {
        "impl": "reference", "metric": "reports_per_sec", "value": value, "unit": "reports/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "reports/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reports/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch_per_gpu, world):
    mode = "greedy" if args.num_beams == 1 else "beam search x%d%s" % (args.num_beams, ", early stopping" if args.early_stopping else "")
    which = ("configs[1]" if (args.num_beams == 1 and args.max_length == 64 and args.image_size == 512 and batch_per_gpu == 32) else
             "configs[2]" if (args.num_beams == 1 and args.max_length == 128 and args.image_size == 512) else
             "configs[3]" if (args.num_beams == 4 and args.max_length == 128) else
             "configs[4]" if args.image_size == 1024 else "variant")
    return {"workload": "ReportGenerationModel.generate: %dx%d synthetic grayscale, batch %d per GPU, %s, 29 regions, "
                        "max_length %d (BASELINE.json %s)" % (args.image_size, args.image_size, batch_per_gpu, mode, args.max_length, which),
            "global_batch": batch_per_gpu * world, "image_size": args.image_size, "max_length": args.max_length,
            "num_beams": args.num_beams,
            "parallelism": "images sharded over %d GPU(s), no data-path collective; 1 NCCL all-gather of result blobs per step" % world,
            "weights": "rgrg_b200.synth seed 0 (conditioned random init, SURVEY.md §8(d))",
            "l2": "per-step working set (KV cache + weights + RoI features > 10 GB) >> 126 MB L2; input batch alternates between two seeds"}


def cpu_baseline(args, sd):
    """Bounded CPU sample on rank 0: ONE image of the same workload through the reference's own generate()."""
    from rgrg_b200 import synth

    kind, gen, cores = load_cpu_reference(sd)
    img = synth.synthetic_images(1, args.image_size, seed=2000)
    dt, R = timed_cpu_generate(gen, img, args.max_length)
    what = "the UNMODIFIED reference (oracle/_ref)" if kind == "reference" else "fp32 oracle port of the reference algorithm"
    return {"value": 1.0 / dt, "unit": "reports/s", "cores": cores, "kind": kind,
            "sample": "1 image, %dx%d, greedy max_length=%d, %s, fp32 CPU, one full generate(): %.1f s, R=%d rows"
                      % (args.image_size, args.image_size, args.max_length, what, dt, R)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--max-length", type=int, default=64)
    ap.add_argument("--image-size", type=int, default=512)
    ap.add_argument("--num-beams", type=int, default=1, help="beam search width (BASELINE.json configs[3] uses 4)")
    ap.add_argument("--early-stopping", action="store_true")
    ap.add_argument("--gather", default="native", choices=["native", "torch"],
                    help="multi-GPU result gather: the engine's own ncclAllGather (device-side pack) or torch.distributed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    from rgrg_b200 import ReportGenerationModel, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    log("rank %d/%d: building synthetic checkpoint" % (rank, world))
    sd = synth.make_state_dict(0) if rank == 0 else None
    if world > 1:
        dist.barrier()
        if rank != 0:
            sd = synth.make_state_dict(0)  # rank 0 has filled the cache
    model = ReportGenerationModel(pretrain_without_lm_model=True)
    model.load_state_dict(sd)
    model.to(dev)
    model.eval()
    log("loading weights into the engine")
    eng = model._engine()

    B, T, S = args.batch, args.max_length, args.image_size
    host_batches = [synth.synthetic_images(B, S, seed=1000 + 17 * rank + i).pin_memory() for i in range(2)]
    dev_batches = [b.to(dev) for b in host_batches]

    from rgrg_b200 import parallel

    for kv in filter(None, os.environ.get("RGRG_OPTS", "").split(",")):  # tuning only: engine options, e.g. RGRG_OPTS=epi_tma=0
        k, v = kv.split("=")
        eng.set_option(k, int(v))
        log("engine option %s=%s" % (k, v))

    native_gather = False
    if world > 1 and args.gather == "native":
        try:
            parallel.init_engine_comm(eng, device=dev)  # the engine's own NCCL communicator (rgrg_comm_init)
            native_gather = True
        except Exception as exc:  # noqa: BLE001
            log("engine-side NCCL gather unavailable (%s): using torch.distributed" % exc)
        flag = torch.tensor([1 if native_gather else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank must take the same path
        native_gather = bool(flag.item())

    def collect(out):
        """C1 of SURVEY.md: one all-gather of the fixed-size per-rank result blob (ids + masks + boxes + scores): packed and
        gathered by the engine on the device (rgrg_allgather_results), or through torch.distributed on a host-packed blob."""
        if world == 1:
            return out
        if native_gather:
            return parallel.all_gather_results_native(eng, B, T)
        return parallel.all_gather_results(out, B, T, device=dev)

    def step_device(i):
        out = eng.generate(dev_batches[i % 2], T, args.num_beams, args.early_stopping)
        collect(out)
        return out

    def step_e2e(i):
        out = model.generate(host_batches[i % 2], max_length=T, num_beams=args.num_beams,
                             early_stopping=args.early_stopping)  # reference-facing call, host images
        if out != -1:
            ids = out[0]
            _ = ids.cpu()  # D2H read of the result, as the reference's caller does before tokenizer.batch_decode
        return out

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    log("warm-up x%d" % args.warmup)
    # ---- warm-up
    last = None
    for i in range(args.warmup):
        last = step_device(i)
    sync_all()

    log("timed region (device-resident inputs) x%d" % args.steps)
    # ---- timed: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for i in range(args.steps):
        last = step_device(i)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * args.steps / (ms_max / 1e3)

    log("%.1f reports/s; timed region (e2e, host buffers)" % value)
    # ---- timed: end to end through the reference-facing API, host buffers
    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    sync_all()
    ev0.record()
    for i in range(args.steps):
        step_e2e(i)
    ev1.record()
    sync_all()
    t2 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t2.item()) / 1e3)
    R = int(last["R"])
    d2h = R * T * 4 + 2 * B * 29 + B * 29 * 16 + B * 29 * 4 + R * T * 8  # engine outputs + the caller's ids.cpu()
    e2e = {"value": e2e_value, "unit": "reports/s", "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": d2h}

    log("e2e %.1f reports/s; profiled step" % e2e_value)
    # ---- roofline: one profiled step (CUDA events around every kernel category on the launch stream)
    roofline, breakdown = None, None
    if rank == 0:
        peaks = load_peaks()
        eng.set_option("profile", 1)
        eng.generate(dev_batches[0], T, args.num_beams, args.early_stopping)
        prof = eng.profile_read()
        eng.set_option("profile", 0)
        P_total = int(sum(eng.detect(dev_batches[0])["num_proposals"]))
        tot = sum(v[0] for v in prof.values())
        breakdown = {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / tot, 4)}
                     for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        top = next(iter(breakdown))
        ms_total, n = prof[top]
        avg_ms = ms_total / n
        timing = "CUDA events around every launch of the category, eager launches (includes launch gaps)"
        # The eager per-launch events include launch gaps that the timed region (CUDA-graph replay) does not pay.  For the
        # decode-step kernels, re-measure the dominant category under graph replay by ablation: decode-only time with and
        # without that category, CUDA events on the launch stream, difference / launches.
        ABLATE = {"attention": 1, "layernorm": 2, "c_attn": 4, "attn_c_proj": 8, "mlp_c_fc": 16, "mlp_c_proj": 32, "attn_fused": 64}
        if top in ABLATE and R > 0 and args.num_beams == 1:
            det = eng.detect(dev_batches[0])
            feats = torch.from_numpy(det["region_features"][det["selected"]]).to(dev)

            def decode_ms(mask):
                eng.set_option("ablate", mask)
                eng.lm_generate(feats, T)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    eng.lm_generate(feats, T)
                b.record()
                torch.cuda.synchronize(dev)
                return a.elapsed_time(b) / 3

            full, without = decode_ms(0), decode_ms(ABLATE[top])
            eng.set_option("ablate", 0)
            avg_ms = (full - without) / ((T - 1) * 24)
            timing = ("ablation under CUDA-graph replay: (decode with - decode without the category) / launches, CUDA events; "
                      "eager per-launch events gave %.4f ms" % (ms_total / n))
        fl = category_flops(top, R, P_total, B, S)
        if fl is not None:
            achieved = fl / (avg_ms / 1e3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
            roofline = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": None,
                        "peak_source": peaks["_source"] + " (sustained bf16 GEMM)",
                        "launches": n, "avg_ms": avg_ms, "timing": timing}
        else:
            by = category_bytes(top, R, (T + 2) / 2.0)
            if by is not None:
                achieved = by / (avg_ms / 1e3) / 1e9
                roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / peaks["hbm_gbs"],
                            "traffic": (by * ncu_traffic(top)["ratio"]) if (ncu_traffic(top) or {}).get("ratio") else None,
                            "traffic_note": (ncu_traffic(top) or {}).get("note"), "peak_source": peaks["_source"],
                            "launches": n, "avg_ms": avg_ms, "timing": timing, "algorithmic_bytes_per_launch": by,
                            "note": "algorithmic bytes per launch = rows x (mean L - 1) x 4096 B cached K/V read + rows x 4096 B "
                                    "appended + x / attention-output rows + the c_attn weight (DESIGN.md §4), mean L over the T - 1 steps"}

    if rank == 0:
        P_est = P_total / B if breakdown is not None else 850
        whole = flops_per_image(S, P_est, R * args.num_beams / B, T) * B * world * args.steps / (ms_max / 1e3) / 1e12
        peaks = load_peaks()
        tensor_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * world
        line = {
            "metric": "reports_per_sec", "value": value, "unit": "reports/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "whole_path": {"algorithmic_tflops": whole, "frac": whole / tensor_peak, "peak_tflops": tensor_peak,
                           "peak_source": peaks["_source"] + " (sustained bf16 GEMM x n_gpus)", "rows_per_gpu": R * args.num_beams,
                           "proposals_per_image": P_est,
                           "note": "fraction of the conv + GEMM roofline: SURVEY.md §8(d) algorithmic FLOPs of the whole path / step time"},
            "kernel_breakdown": breakdown,
        }
        if not args.no_cpu_baseline and world == 1:
            log("cpu baseline (bounded sample)")
            line["cpu_baseline"] = cpu_baseline(args, sd)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
