"""world_size-2 gloo tests of the multi-GPU host logic (sharding + clip-ordered all-gather +
PSNR reduction, model/pfnl.py:90,139-141) - runs on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pfnl_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    s, e = D.shard_range(n_total, rank, world)
    # the per-clip "MSE" of clip i is (i+1)/100 so ordering mistakes are visible
    local = torch.arange(s, e, dtype=torch.float32).add(1).div(100)
    full = D.all_gather_clips(local, n_total)
    frames = torch.arange(s, e, dtype=torch.float32)[:, None, None].expand(e - s, 2, 3).contiguous()
    full_frames = D.all_gather_clips(frames, n_total)
    tmax = D.max_over_ranks(1.0 + rank)
    q.put((rank, full.numpy(), full_frames[:, 0, 0].numpy(), tmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_allgather_orders_clips(n_total, built_lib):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (np.arange(n_total, dtype=np.float32) + 1) / 100
    for rank, full, frames, tmax in res:
        np.testing.assert_allclose(full, expect)
        np.testing.assert_array_equal(frames, np.arange(n_total, dtype=np.float32))
        assert tmax == 2.0
    from pfnl_b200.dist import psnr_np
    np.testing.assert_allclose(psnr_np(expect), 10 * np.log10(1 / expect.astype(np.float64)))
