"""Multi-GPU plumbing: images shard embarrassingly across ranks (one process per GPU); the only exchange on the path is
ONE all-gather of fixed-size per-rank result blobs at the end of generate().  The reference has no
distributed code at all (single process, single GPU: generate_reports_for_images.py:23)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

NUM_REGIONS = 29
EOS = 50256


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def blob_bytes(batch: int, max_length: int) -> int:
    rows = batch * NUM_REGIONS
    return 8 + rows * max_length * 4 + 2 * rows + rows * 16 + rows * 4


def pack_result(out: Dict, batch: int, max_length: int) -> np.ndarray:
    """Engine.generate() result of one rank -> fixed-size uint8 blob (equal counts are what all_gather needs)."""
    rows = batch * NUM_REGIONS
    ids = np.full((rows, max_length), EOS, dtype=np.int32)
    r, w = out["ids"].shape if out["R"] > 0 else (0, 0)
    if r:
        ids[:r, :w] = out["ids"]
    head = np.array([out["R"], w], dtype=np.int32)
    parts = [head.view(np.uint8), ids.reshape(-1).view(np.uint8),
             np.ascontiguousarray(out["selected"], dtype=np.uint8).reshape(-1),
             np.ascontiguousarray(out["detected"], dtype=np.uint8).reshape(-1),
             np.ascontiguousarray(out["boxes"], dtype=np.float32).reshape(-1).view(np.uint8),
             np.ascontiguousarray(out["scores"], dtype=np.float32).reshape(-1).view(np.uint8)]
    blob = np.concatenate(parts)
    assert blob.size == blob_bytes(batch, max_length)
    return blob


def unpack_result(blob: np.ndarray, batch: int, max_length: int) -> Dict:
    rows = batch * NUM_REGIONS
    o = 0
    head = blob[o:o + 8].view(np.int32); o += 8
    ids = blob[o:o + rows * max_length * 4].view(np.int32).reshape(rows, max_length); o += rows * max_length * 4
    sel = blob[o:o + rows].reshape(batch, NUM_REGIONS).astype(bool); o += rows
    det = blob[o:o + rows].reshape(batch, NUM_REGIONS).astype(bool); o += rows
    boxes = blob[o:o + rows * 16].view(np.float32).reshape(batch, NUM_REGIONS, 4); o += rows * 16
    scores = blob[o:o + rows * 4].view(np.float32).reshape(batch, NUM_REGIONS)
    R, w = int(head[0]), int(head[1])
    return {"R": R, "ids": ids[:R, :w].copy(), "selected": sel, "detected": det, "boxes": boxes.copy(), "scores": scores.copy()}


def merge_results(parts: List[Dict]) -> Dict:
    """Concatenate per-rank results in rank (= image) order, as if one rank had processed the whole batch.  Greedy ids of
    different ranks may have different widths (each rank stops when ITS rows are done): pad to the widest with EOS,
    which is what the single-process reference would have emitted for finished rows (language_model.py:636)."""
    width = max([p["ids"].shape[1] for p in parts if p["R"] > 0], default=0)
    ids = [np.pad(p["ids"], ((0, 0), (0, width - p["ids"].shape[1])), constant_values=EOS) for p in parts if p["R"] > 0]
    return {"R": sum(p["R"] for p in parts),
            "ids": np.concatenate(ids, 0) if ids else np.zeros((0, 0), np.int32),
            "selected": np.concatenate([p["selected"] for p in parts], 0),
            "detected": np.concatenate([p["detected"] for p in parts], 0),
            "boxes": np.concatenate([p["boxes"] for p in parts], 0),
            "scores": np.concatenate([p["scores"] for p in parts], 0)}


def all_gather_results(out: Dict, batch: int, max_length: int, device=None) -> Dict:
    """One all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) of the packed per-rank results."""
    import torch.distributed as dist

    world = dist.get_world_size()
    blob = torch.from_numpy(pack_result(out, batch, max_length))
    if device is not None:
        blob = blob.to(device)
    buf = torch.empty(world * blob.numel(), dtype=torch.uint8, device=blob.device)
    dist.all_gather_into_tensor(buf, blob)
    flat = buf.cpu().numpy()
    n = blob.numel()
    return merge_results([unpack_result(flat[r * n:(r + 1) * n], batch, max_length) for r in range(world)])


def init_engine_comm(engine, device=None):
    """Create the engine's own NCCL communicator (rgrg_comm_init): rank 0's ncclUniqueId travels through the already
    initialised torch.distributed group (any backend)."""
    import torch.distributed as dist

    def bcast(buf):
        t = buf.to(device) if (device is not None and dist.get_backend() == "nccl") else buf
        dist.broadcast(t, src=0)
        return t

    engine.comm_init(dist.get_rank(), dist.get_world_size(), bcast)


def all_gather_results_native(engine, batch: int, max_length: int) -> Dict:
    """The same merge as all_gather_results, with pack + all-gather done by the engine on the device
    (rgrg_allgather_results: one ncclAllGather, no host-side packing, no torch tensors on the data path)."""
    blobs = engine.allgather_results(batch, max_length)
    return merge_results([unpack_result(blobs[r], batch, max_length) for r in range(blobs.shape[0])])

