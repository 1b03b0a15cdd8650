"""N > 1 host logic on CPU: contiguous image sharding + the single end-of-job reduction, run as
two real processes over gloo (what torchrun + NCCL does on the GPU box)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgic_b200 import dist as cdist


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [cdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = cdist.shard_range(n_images, rank, world)
        # per-image stream sizes are a deterministic function of the global image id
        sizes = torch.stack([torch.arange(5, dtype=torch.int32) + 10 * i for i in range(lo, hi)])
        nbytes = float(sizes.sum())
        b, p, s, bpp = cdist.reduce_rate_distortion(nbytes, (hi - lo) * 65536.0, 0.5 * (hi - lo))
        allsz = cdist.gather_sizes(sizes)
        out.put((rank, b, p, s, bpp, allsz.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    n_images, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_sizes = [[j + 10 * i for j in range(5)] for i in range(n_images)]
    want_bytes = float(sum(map(sum, want_sizes)))
    for rank, b, p, s, bpp, allsz in res:
        assert b == want_bytes and p == n_images * 65536.0 and s == 0.5 * n_images
        assert bpp == 8.0 * want_bytes / (n_images * 65536.0)
        assert allsz == want_sizes
    assert cdist.reduce_rate_distortion(3.0, 4.0, 5.0) == (3.0, 4.0, 5.0, 6.0)   # no process group: identity
