"""The N > 1 host logic on CPU: world_size-2 gloo processes split the work the way bench.py / a deployment
would (ARFCN a -> rank a mod G; contiguous 117-frame blocks of one stream per rank), run their shard through the
oracle (standing in for the GPU kernels, which need no inter-rank state), gather the SoftVectors with
shard.gather_soft, and the union equals the single-process result."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from openbts_ttsou_b200 import shard
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "normal_sps1.npz"))
    o = Oracle("port")
    n_arfcn = g["bursts"].shape[0] // 8
    rows = shard.burst_rows_of_arfcns(shard.arfcn_shard(n_arfcn, rank, world))
    r = o.rx_normal_batch(g["bursts"][rows], g["lens"][rows], g["tsc"][rows])
    parts = shard.gather_soft(torch.from_numpy(r["soft"]))
    # the three ways to call it agree (counts exchanged / counts known / equal shards in one collective)
    n_mine = len(rows)
    counts = [len(shard.burst_rows_of_arfcns(shard.arfcn_shard(n_arfcn, k, world))) for k in range(world)]
    for alt in (shard.gather_soft(torch.from_numpy(r["soft"]), counts=counts),
                shard.gather_soft(torch.from_numpy(r["soft"]), counts="equal") if len(set(counts)) == 1 else parts):
        assert all(torch.equal(a, b) for a, b in zip(alt, parts)) and alt[rank].shape[0] == n_mine
    # ragged shards: rank r holds r + 3 rows
    rag = shard.gather_soft(torch.full((rank + 3, 5), float(rank)))
    assert [t.shape[0] for t in rag] == [k + 3 for k in range(world)] and all(bool((t == k).all()) for k, t in enumerate(rag))
    rows_all = [shard.burst_rows_of_arfcns(shard.arfcn_shard(n_arfcn, k, world)) for k in range(world)]
    full = np.zeros_like(g["soft"])
    for k in range(world):
        full[rows_all[k]] = parts[k].numpy()
    ok = bool(np.array_equal(full, g["soft"]))
    # stream sharding: blocks [lo, hi) per rank cover everything exactly once
    lo, hi = shard.stream_shard(7, rank, world)
    spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(spans, torch.tensor([lo, hi]))
    cover = sorted((int(a), int(b)) for a, b in spans)
    ok = ok and cover[0][0] == 0 and cover[-1][1] == 7 and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, len(rows)))


def test_two_rank_sharding_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == 96


def test_shard_helpers():
    from openbts_ttsou_b200 import shard
    assert list(shard.arfcn_shard(10, 1, 4)) == [1, 5, 9]
    assert list(shard.burst_rows_of_arfcns([2])) == list(range(16, 24))
    spans = [shard.stream_shard(855, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 855 and all(spans[i][1] == spans[i + 1][0] for i in range(7))
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert shard.BURSTS_PER_BLOCK * 625 // 4 == shard.CHUNKS_PER_BLOCK * 585
