"""The N>1 path on CPU: world_size 2 over gloo.  Each rank owns a contiguous shard of independent streams, runs
them (here through the oracle, standing in for the per-rank GPU work), and the only exchange is the counter sum."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import audio_modem_b200 as amb
    from helpers import make_stream
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_streams = 5
    lo, hi = amb.shard.shard_range(n_streams, rank, world)
    c = {k: 0 for k in amb.shard.COUNTER_KEYS}
    for s in range(lo, hi):
        x, pay = make_stream(O, "16-QAM", 2, 5000 + s)
        rx = O.StreamingReceiver("standard", "QAM16", 1)
        ev = rx.push(x)
        cc = rx.counters()
        c["framesDecoded"] += cc["framesDecoded"]
        c["frameErrors"] += cc["frameErrors"]
        c["crcValidFrames"] += sum(1 for e in ev if e.get("crcValid"))
        c["payloadBytes"] += sum(e.get("dataLen", 0) for e in ev if e.get("crcValid") and e.get("frameType") == 0xFF)
        c["samples"] += len(x)
    tot = amb.shard.gather_counters(c)
    slow = amb.shard.max_over_ranks(1.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, (lo, hi), tot, slow))


def test_two_rank_shard_and_counter_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, t0, m0), (r1, s1, t1, m1) = out
    assert s0 == (0, 3) and s1 == (3, 5)
    assert t0 == t1                      # every rank holds the same totals after the all_reduce
    assert t0["framesDecoded"] == 15 and t0["crcValidFrames"] == 15 and t0["payloadBytes"] == 10 * 4096 and t0["frameErrors"] == 0
    assert m0 == m1 == 2.0
