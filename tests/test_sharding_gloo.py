"""N > 1 host logic on CPU: world_size-2 gloo run of the shard/all-gather plumbing (no GPU involved: the
per-shard verdicts are synthetic), plus pure shard-bound properties."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rusty_kaspa_b200 import sharding


def test_shard_bounds_cover_and_align():
    for n in [0, 1, 7, 8, 9, 1000, 1 << 20, (1 << 20) + 3]:
        for w in [1, 2, 3, 4, 8]:
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and (hi - lo) % 8 == 0 or hi == n


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        status = rng.integers(0, 4, size=n, dtype=np.uint8)  # every rank derives the same global ground truth
        lo, hi = sharding.shard_range(n, rank, world)
        local = np.packbits((status[lo:hi] == 1).astype(np.uint8), bitorder="little")
        got = sharding.all_gather_bitmaps(torch.from_numpy(local.copy()), n).numpy()
        exp = np.packbits((status == 1).astype(np.uint8), bitorder="little")
        q.put((rank, bool((got == exp).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 4099])
def test_bitmap_all_gather_world2_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
