"""world_size-2 gloo test (CPU) of the multi-GPU host logic bench.py uses: contiguous agent shards,
one all-gather of the 24-byte neighbour records per tick, identical replicated state on every rank."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, out):
    import importlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("bench").shard_range
    lo, hi = shard(n, rank, world)
    rng = np.random.default_rng(5)
    full = rng.random((n, 6)).astype(np.float32)            # the 24-byte records, uid order
    mine = torch.from_numpy(full[lo:hi].copy())
    counts = [shard(n, r, world)[1] - shard(n, r, world)[0] for r in range(world)]
    pad = max(counts)
    buf = torch.zeros((pad, 6)); buf[:hi - lo] = mine
    gathered = [torch.zeros((pad, 6)) for _ in range(world)]
    dist.all_gather(gathered, buf)
    rebuilt = torch.cat([g[:c] for g, c in zip(gathered, counts)]).numpy()
    ok = bool((rebuilt == full).all()) and sum(counts) == n and lo == sum(counts[:rank])
    t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t.item()))
    dist.destroy_process_group()


def test_shard_allgather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) == 1


def test_shard_range_covers_everything():
    import importlib
    shard = importlib.import_module("bench").shard_range
    for n in (0, 1, 7, 100_000, 1_000_003):
        for w in (1, 2, 4, 8):
            edges = [shard(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1


def test_bench_population_builds_for_8_ranks():
    """bench.py's weak-scaling populations at the driver's largest GPU count (config C4 at N = 8 = BASELINE configs[3]: 4 M
    agents, 64 flocks, one per cell of an 8 x 8 grid over the 2048^2 map -- ~94 % of the passable area): every flock must fit
    its cell (cells with more obstacles pack a few per cent tighter), and the N = 2 population must be a prefix of it."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    pf = importlib.import_module("permafrost-engine_b200")
    bench.set_workload("C4")
    try:
        w8 = bench.build_workload(pf, 8, 7)
        a = w8["agents"]
        assert len(a["pos"]) == 4_000_000 and w8["nflocks"] == 64 and w8["hi"] - w8["lo"] == 500_000
        assert (np.bincount(a["flock_of"], minlength=64) == 62500).all()
        for f in range(64):                                  # hex packing at (nearly) the nominal spacing, no overlaps
            p = a["pos"][a["flock_of"] == f][:400]
            d = np.linalg.norm(p[1:] - p[0], axis=1).min()
            assert 3.3 <= d <= 3.91, (f, d)
        w2 = bench.build_workload(pf, 2, 0)
        assert np.array_equal(w2["agents"]["pos"], a["pos"][:len(w2["agents"]["pos"])])
    finally:
        bench.set_workload("C2")
