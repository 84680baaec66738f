"""Thin Python owner of one C-ABI engine handle (one per GPU; calls are serialised by the caller)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi

NUM_REGIONS = 29
VOCAB = 50257
EOS = 50256


def _ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(None)
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("rgrg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = int(device)
        h = C.c_void_p()
        if self._lib.rgrg_create(self.device, C.byref(h)) != 0:
            raise RuntimeError(self._lib.rgrg_last_error(None).decode())
        self._h = h
        self._keepalive = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rgrg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- errors: reference callers string-match "out of memory" (evaluate_language_model.py:1208)
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(self._lib.rgrg_last_error(self._h).decode())

    def set_option(self, key: str, value: int):
        self._check(self._lib.rgrg_set_option(self._h, key.encode(), int(value)))

    def profile_read(self):
        """-> {category: (total_ms, launches)} since set_option("profile", 1)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.rgrg_profile_read(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.rgrg_kernel_launches(self._h))

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], detector_precise: bool = False):
        """detector_precise: also keep fp32 twins of the detector weights and run the detector in fp32 on CUDA cores
        (parity mode: region indices comparable with the fp32 reference; ~100x slower than the bf16 tensor-core path)."""
        if detector_precise:
            self.set_option("detector_precise", 1)
        keep = []
        for name, t in state_dict.items():
            if not torch.is_tensor(t) or not t.dtype.is_floating_point:
                continue  # num_batches_tracked, causal_mask ...
            t = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            shape = (C.c_int64 * max(t.dim(), 1))(*(list(t.shape) or [1]))
            self._check(self._lib.rgrg_load_weight(self._h, name.encode(), _ptr(t), shape, max(t.dim(), 1)))
        self._check(self._lib.rgrg_finalize_weights(self._h))
        del keep

    # ---- the path
    def generate(self, images: torch.Tensor, max_length: int, num_beams: int = 1, early_stopping: bool = False):
        B, S = int(images.shape[0]), int(images.shape[-1])
        on_host = images.device.type == "cpu"
        images = images.to(torch.float32).contiguous()
        ids = np.full((B * NUM_REGIONS, max_length), EOS, dtype=np.int32)
        sel = np.zeros((B, NUM_REGIONS), dtype=np.uint8)
        det = np.zeros((B, NUM_REGIONS), dtype=np.uint8)
        boxes = np.zeros((B, NUM_REGIONS, 4), dtype=np.float32)
        scores = np.zeros((B, NUM_REGIONS), dtype=np.float32)
        R, width = C.c_int(0), C.c_int(0)
        self._check(self._lib.rgrg_generate(self._h, _ptr(images), int(on_host), B, S, int(max_length), int(num_beams),
                                            int(bool(early_stopping)), _ptr(ids), C.byref(width), _ptr(sel), _ptr(det),
                                            _ptr(boxes), _ptr(scores), C.byref(R), _stream(self.device)))
        return {"ids": ids[: R.value, : width.value], "R": R.value, "selected": sel.astype(bool),
                "detected": det.astype(bool), "boxes": boxes, "scores": scores}

    # ---- multi-GPU result gather (one ncclAllGather issued by the engine, device to device)
    def comm_init(self, rank: int, world: int, broadcast_bytes):
        """broadcast_bytes(buf: bytearray-like uint8 tensor of 128 bytes, src=0): fills `buf` on every rank with rank 0's
        content (e.g. a torch.distributed broadcast)."""
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            if self._lib.rgrg_comm_unique_id(_ptr(buf)) != 0:
                raise RuntimeError(self._lib.rgrg_last_error(None).decode())
            ident = torch.from_numpy(buf)
        ident = broadcast_bytes(ident)
        ident = np.ascontiguousarray(ident.cpu().numpy())
        self._check(self._lib.rgrg_comm_init(self._h, _ptr(ident), int(rank), int(world)))
        self._world = int(world)

    def allgather_results(self, batch: int, max_length: int) -> np.ndarray:
        """-> uint8 [world, blob_bytes]: the packed results (rgrg_b200.parallel layout) of the last generate() of every rank."""
        from . import parallel

        n = parallel.blob_bytes(batch, max_length)
        out = np.zeros((self._world, n), dtype=np.uint8)
        self._check(self._lib.rgrg_allgather_results(self._h, int(batch), int(max_length), _ptr(out), n, _stream(self.device)))
        return out

    def preprocess(self, images) -> torch.Tensor:
        """list of uint8 grayscale arrays [H, W] (numpy, host) or CUDA uint8 tensors -> fp32 CUDA tensor [B, 1, 512, 512]
        (generate_reports_for_images.py:129-147 `get_image_tensor`, batched)."""
        dev = torch.device("cuda", self.device)
        out = torch.empty((len(images), 1, 512, 512), dtype=torch.float32, device=dev)
        for i, im in enumerate(images):
            if torch.is_tensor(im) and im.is_cuda:
                im = im.contiguous()
                assert im.dtype == torch.uint8 and im.dim() == 2
                on_host, H, W = 0, int(im.shape[0]), int(im.shape[1])
            else:
                im = np.ascontiguousarray(im)
                if im.dtype != np.uint8 or im.ndim != 2:
                    raise ValueError("pre-processing expects 8-bit single-channel images [H, W]")
                on_host, H, W = 1, im.shape[0], im.shape[1]
            self._check(self._lib.rgrg_preprocess(self._h, _ptr(im), on_host, H, W, C.c_void_p(out[i].data_ptr()), 0,
                                                  _stream(self.device)))
        return out

    def lm_generate(self, feats: torch.Tensor, max_length: int, num_beams: int = 1, early_stopping: bool = False):
        R = int(feats.shape[0])
        on_host = feats.device.type == "cpu"
        feats = feats.to(torch.float32).contiguous()
        ids = np.full((R, max_length), EOS, dtype=np.int32)
        width = C.c_int(0)
        self._check(self._lib.rgrg_lm_generate(self._h, _ptr(feats), int(on_host), R, int(max_length), int(num_beams),
                                               int(bool(early_stopping)), _ptr(ids), C.byref(width), _stream(self.device)))
        return ids[:, : width.value]

    def detect(self, images: torch.Tensor, abnormal: bool = False):
        B, S = int(images.shape[0]), int(images.shape[-1])
        on_host = images.device.type == "cpu"
        images = images.to(torch.float32).contiguous()
        sel = np.zeros((B, NUM_REGIONS), dtype=np.uint8)
        det = np.zeros((B, NUM_REGIONS), dtype=np.uint8)
        boxes = np.zeros((B, NUM_REGIONS, 4), dtype=np.float32)
        scores = np.zeros((B, NUM_REGIONS), dtype=np.float32)
        trf = np.zeros((B, NUM_REGIONS, 1024), dtype=np.float32)
        top_idx = np.zeros((B, NUM_REGIONS), dtype=np.int32)
        nprop = np.zeros((B,), dtype=np.int32)
        abn = np.zeros((B, NUM_REGIONS), dtype=np.uint8) if abnormal else None
        R = C.c_int(0)
        self._check(self._lib.rgrg_detect(self._h, _ptr(images), int(on_host), B, S, _ptr(sel), _ptr(det), _ptr(boxes),
                                          _ptr(scores), _ptr(trf), _ptr(top_idx), _ptr(nprop), _ptr(abn), C.byref(R),
                                          _stream(self.device)))
        out = {"selected": sel.astype(bool), "detected": det.astype(bool), "boxes": boxes, "scores": scores,
               "region_features": trf, "top_idx": top_idx, "num_proposals": nprop, "R": R.value}
        if abnormal:
            out["predicted_abnormal_regions"] = abn.astype(bool)  # report_generation_model.py:103-106 (eval-mode forward)
        return out

    def bbox_features(self, images: torch.Tensor, boxes) -> np.ndarray:
        """boxes: [B,29,4] (tensor / array / list of 29x4 tensors) -> fp32 [B*29, 1024]"""
        B, S = int(images.shape[0]), int(images.shape[-1])
        on_host = images.device.type == "cpu"
        images = images.to(torch.float32).contiguous()
        if isinstance(boxes, (list, tuple)):
            boxes = torch.stack([b.detach().to("cpu", torch.float32) for b in boxes])
        boxes = np.ascontiguousarray(torch.as_tensor(boxes).detach().to("cpu", torch.float32).numpy().reshape(B, NUM_REGIONS, 4))
        out = np.zeros((B * NUM_REGIONS, 1024), dtype=np.float32)
        self._check(self._lib.rgrg_bbox_features(self._h, _ptr(images), int(on_host), B, S, _ptr(boxes), _ptr(out),
                                                 _stream(self.device)))
        return out

    # ---- stage-level (tests / roofline harness); all tensors are CUDA tensors on self.device
    def lm_forced_logits(self, feats: torch.Tensor, forced_ids: torch.Tensor) -> torch.Tensor:
        R, n = int(forced_ids.shape[0]), int(forced_ids.shape[1])
        feats = feats.to(torch.float32).contiguous()
        forced_ids = forced_ids.to(torch.int32).contiguous()
        out = torch.empty((n, R, VOCAB), dtype=torch.float32, device=feats.device)
        self._check(self._lib.rgrg_lm_forced_logits(self._h, _ptr(feats), R, _ptr(forced_ids), n, _ptr(out),
                                                    _stream(self.device)))
        return out

    def greedy_bookkeeping(self, logits_steps: torch.Tensor, max_length: int):
        n_steps, R = int(logits_steps.shape[0]), int(logits_steps.shape[1])
        ids = np.full((R, max_length), EOS, dtype=np.int32)
        width = C.c_int(0)
        self._check(self._lib.rgrg_greedy_bookkeeping(self._h, _ptr(logits_steps.contiguous()), n_steps, R, max_length,
                                                      _ptr(ids), C.byref(width), _stream(self.device)))
        return ids[:, : width.value]

    def beam_bookkeeping(self, logits_steps: torch.Tensor, sentences: int, num_beams: int, max_length: int,
                         early_stopping: bool):
        n_steps = int(logits_steps.shape[0])
        ids = np.full((sentences, max_length), EOS, dtype=np.int32)
        width = C.c_int(0)
        self._check(self._lib.rgrg_beam_bookkeeping(self._h, _ptr(logits_steps.contiguous()), n_steps, sentences, num_beams,
                                                    max_length, int(bool(early_stopping)), _ptr(ids), C.byref(width),
                                                    _stream(self.device)))
        return ids[:, : width.value]

    def rpn_filter(self, objectness, deltas=None, decoded=None, feat=16, image_size=512):
        B = int(objectness.shape[0])
        dev = objectness.device
        boxes = torch.zeros((B, 1000, 4), dtype=torch.float32, device=dev)
        scores = torch.zeros((B, 1000), dtype=torch.float32, device=dev)
        count = torch.zeros((B,), dtype=torch.int32, device=dev)
        topk = torch.full((B, 1000), -1, dtype=torch.int32, device=dev)
        keep = torch.full((B, 1000), -1, dtype=torch.int32, device=dev)
        objectness = objectness.contiguous()
        deltas = deltas.contiguous() if deltas is not None else None
        decoded = decoded.contiguous() if decoded is not None else None
        self._check(self._lib.rgrg_rpn_filter(self._h, _ptr(objectness), _ptr(deltas), _ptr(decoded), B, feat, image_size,
                                              _ptr(boxes), _ptr(scores), _ptr(count), _ptr(topk), _ptr(keep),
                                              _stream(self.device)))
        return boxes, scores, count, topk, keep

    def roi_align(self, feats_nhwc_bf16, boxes, count, image_size=512):
        B, f, _, Cc = feats_nhwc_bf16.shape
        total = int(count.sum().item())
        out = torch.empty((max(total, 1), 64, Cc), dtype=torch.bfloat16, device=feats_nhwc_bf16.device)
        self._check(self._lib.rgrg_roi_align(self._h, _ptr(feats_nhwc_bf16.contiguous()), _ptr(boxes.contiguous()),
                                             _ptr(count.contiguous()), int(B), int(f), int(Cc), image_size, _ptr(out),
                                             _stream(self.device)))
        return out[:total]

    def roi_tail(self, class_logits, box_regression, boxes, count, image_size=512):
        B = int(count.shape[0])
        dev = class_logits.device
        det = torch.zeros((B, 29), dtype=torch.uint8, device=dev)
        idx = torch.zeros((B, 29), dtype=torch.int32, device=dev)
        scores = torch.zeros((B, 29), dtype=torch.float32, device=dev)
        tb = torch.zeros((B, 29, 4), dtype=torch.float32, device=dev)
        self._check(self._lib.rgrg_roi_tail(self._h, _ptr(class_logits.contiguous()), _ptr(box_regression.contiguous()),
                                            _ptr(boxes.contiguous()), _ptr(count.contiguous()), B, image_size, _ptr(det),
                                            _ptr(idx), _ptr(scores), _ptr(tb), _stream(self.device)))
        return det.bool(), idx, scores, tb

    def gemm(self, A_bf16, W_bf16, bias=None, act=0, impl=0):
        M, K = A_bf16.shape
        N = W_bf16.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=A_bf16.device)
        self._check(self._lib.rgrg_gemm_bf16(self._h, _ptr(A_bf16.contiguous()), _ptr(W_bf16.contiguous()), _ptr(bias),
                                             int(M), int(N), int(K), int(act), int(impl), _ptr(out), _stream(self.device)))
        return out

    def gemm_bench(self, M, N, K, bn, iters=50, interleave=False, trace_ctas=0):
        ms = C.c_float(0)
        trace = np.zeros((max(trace_ctas, 1), 8), dtype=np.int64)
        self._check(self._lib.rgrg_gemm_bench(self._h, M, N, K, bn, iters, int(interleave), C.byref(ms),
                                              _ptr(trace) if trace_ctas else None, trace_ctas))
        return ms.value, trace

    def conv3x3(self, x_nhwc_bf16, w_bf16, bias=None, relu=False, implicit=True):
        B, H, W, Cin = x_nhwc_bf16.shape
        Cout = w_bf16.shape[0]
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x_nhwc_bf16.device)
        self._check(self._lib.rgrg_conv3x3_bf16(self._h, _ptr(x_nhwc_bf16.contiguous()), _ptr(w_bf16.contiguous()),
                                                _ptr(bias), int(B), int(H), int(W), int(Cin), int(Cout), int(relu),
                                                int(implicit), _ptr(out), _stream(self.device)))
        return out

    def backbone(self, images: torch.Tensor) -> torch.Tensor:
        B, S = int(images.shape[0]), int(images.shape[-1])
        f = S // 32
        out = torch.empty((B, f, f, 2048), dtype=torch.bfloat16, device=images.device)
        self._check(self._lib.rgrg_backbone(self._h, _ptr(images.to(torch.float32).contiguous()), B, S, _ptr(out),
                                            _stream(self.device)))
        return out

    def debug_read(self, name: str, shape, dtype=np.float32) -> np.ndarray:
        out = np.zeros(shape, dtype=dtype)
        self._check(self._lib.rgrg_debug_read(self._h, name.encode(), _ptr(out), out.nbytes))
        return out
