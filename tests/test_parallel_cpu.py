"""world_size-2 gloo test of the N>1 plumbing (image sharding + the single all-gather of result blobs)."""
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgrg_b200 import parallel


def _fake_result(rank, batch, T):
    rng = np.random.default_rng(100 + rank)
    sel = rng.random((batch, 29)) > 0.3
    R = int(sel.sum())
    w = T - rank  # ranks may stop at different widths
    ids = rng.integers(0, 50000, size=(R, w)).astype(np.int32)
    ids[:, 0] = 50256
    return {"R": R, "ids": ids, "selected": sel, "detected": sel | (rng.random((batch, 29)) > 0.5),
            "boxes": rng.random((batch, 29, 4)).astype(np.float32), "scores": rng.random((batch, 29)).astype(np.float32)}


def _worker(rank, world, port, batch, T, q):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    merged = parallel.all_gather_results(_fake_result(rank, batch, T), batch, T)
    if rank == 0:
        q.put({k: (v if isinstance(v, int) else v.copy()) for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    for n in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_pack_roundtrip():
    r = _fake_result(0, 3, 9)
    u = parallel.unpack_result(parallel.pack_result(r, 3, 9), 3, 9)
    for k in ("ids", "selected", "detected", "boxes", "scores"):
        assert np.array_equal(u[k], r[k])
    assert u["R"] == r["R"]


def test_all_gather_world2_gloo():
    world, batch, T = 2, 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (torch.randint(0, 2000, (1,)).item())
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [_fake_result(r, batch, T) for r in range(world)]
    assert merged["R"] == sum(p["R"] for p in parts)
    assert merged["selected"].shape == (world * batch, 29)
    assert np.array_equal(merged["selected"], np.concatenate([p["selected"] for p in parts]))
    assert np.array_equal(merged["boxes"], np.concatenate([p["boxes"] for p in parts]))
    assert merged["ids"].shape == (merged["R"], T)
    r0 = parts[0]["R"]
    assert np.array_equal(merged["ids"][:r0, :T], parts[0]["ids"])
    assert np.array_equal(merged["ids"][r0:, :T - 1], parts[1]["ids"])
    assert (merged["ids"][r0:, T - 1] == 50256).all()
