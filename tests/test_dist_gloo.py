"""world_size-2 gloo tests (CPU) of the host-side sharding logic: endpoint ownership,
exchange of the shard blobs, merge of the per-shard journals, and the bench's
max-over-ranks / sum-over-ranks aggregation.  The CUDA path itself needs GPUs
(tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from maelstrom_b200 import _lib
        from maelstrom_b200.sharded import exchange_blobs, merge_journals, shard_owner
        L = _lib.lib()
        # 1. ownership: Python mirror == C ABI helper; every endpoint has exactly one owner
        n_servers, n_ep = 4096, 4096 + 64
        owners = [shard_owner(e, n_servers, world) for e in range(n_ep)]
        assert owners == [int(L.ms_shard_owner(e, n_servers, world)) for e in range(n_ep)]
        mine = [e for e in range(n_ep) if owners[e] == rank]
        counts = [None] * world
        dist.all_gather_object(counts, len(mine))
        assert sum(counts) == n_ep and max(counts) - min(counts) <= 1
        assert owners[:n_servers] == sorted(owners[:n_servers])          # contiguous row ranges
        # 2. blob exchange (the bytes every rank hands to ms_shard_connect)
        blob = bytes([rank]) * _lib.SHARD_BLOB_BYTES
        blobs = exchange_blobs(blob)
        assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 512 for b in blobs)
        # 3. journal merge: build the oracle journal, give each rank the events of its endpoints
        s = O.Sim(25, topology="grid", n_values=8)
        c = s.add_endpoint("c0")
        for v in range(4):
            s.send(c, (7 * v) % 25, O.body("broadcast", msg_id=v + 1, p0=v))
        s.run(2_000_000)
        ev, _ = s.journal()
        recv = (ev["event_id"] >> np.uint64(63)) != 0
        actor = np.where(recv, ev["dest"], ev["src"])                      # who journals the event
        # sends injected by the host belong to shard 0 (injector), like in the engine
        own = np.array([0 if (not r and a >= 25) else shard_owner(int(a), 25, world) for a, r in zip(actor, recv)])
        part = ev.copy()
        part[own != rank] = np.frombuffer(b"\xff" * ev.dtype.itemsize, dtype=ev.dtype)[0]
        parts = [None] * world
        dist.all_gather_object(parts, part)
        merged = merge_journals(parts)
        assert np.array_equal(merged, ev)
        with pytest.raises(ValueError):
            merge_journals([part, part])
        # 4. bench aggregation: time = max over ranks, work = sum over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([100.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        assert float(t) == 10.0 + world - 1 and float(w) == 100.0 * world * (world + 1) / 2
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharding_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
