"""N > 1 host logic on CPU: world_size-2 gloo run of the stream sharding and the max-over-ranks / whole-job
throughput reductions that bench.py uses under torchrun."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from diart_b200 import parallel

    info = parallel.init("gloo")
    streams = parallel.shard_streams(5, info["world"], info["rank"])
    parallel.barrier()
    seconds = 1.0 + rank                     # rank 1 is the slow one
    chunks = 256 * len(streams)
    value = parallel.throughput(chunks, seconds)
    slowest = parallel.max_over_ranks(seconds)
    maps = parallel.gather_maps([torch.full((2, 3), s, dtype=torch.int32) for s in streams])
    out.put((rank, streams, value, slowest, [[int(m[0, 0]) for m in r] for r in maps]))
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, v0, m0, g0), (r1, s1, v1, m1, g1) = results
    assert s0 == [0, 2, 4] and s1 == [1, 3]                    # stream s -> rank s mod world
    assert m0 == m1 == 2.0                                     # slowest rank
    assert abs(v0 - 5 * 256 * 0.5 / 2.0) < 1e-9 and v0 == v1   # all chunks / slowest time
    assert g0 == g1 == [[0, 2, 4], [1, 3]]


def test_single_process_defaults():
    from diart_b200 import parallel

    assert parallel.shard_streams(3, 1, 0) == [0, 1, 2]
    assert parallel.max_over_ranks(1.5) == 1.5
    assert parallel.throughput(256, 2.0) == 64.0
