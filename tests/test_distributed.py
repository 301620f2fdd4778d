"""CPU: frame sharding + all-gather over gloo, world_size 2 and 3 (ragged)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from confild_b200 import distributed as D
from oracle import cnf_oracle as O


def test_shard_bounds():
    assert D.shard_bounds(8, 2) == [(0, 4), (4, 8)]
    assert D.shard_bounds(7, 3) == [(0, 3), (3, 5), (5, 7)]
    assert D.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for T in (1, 5, 16, 1023):
        for w in (1, 2, 4, 8):
            b = D.shard_bounds(T, w)
            assert b[0][0] == 0 and b[-1][1] == T and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, chunks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sd = O.init_params(2, 16, 3, 2, 32, seed=0)
    coords, lat = O.synthetic_inputs(2, 16, T, 19)
    fn = lambda l: O.forward(sd, coords[None], l[:, None])  # noqa: E731 - checker stands in for the CUDA module
    full = D.decode_frame_sharded(fn, lat, gather=True, chunks=chunks)
    local = D.decode_frame_sharded(fn, lat, gather=False)
    want = fn(lat)
    s, e = D.shard_bounds(T, world)[rank]
    ok = torch.equal(full, want) and torch.equal(local, want[s:e])
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T,chunks", [(2, 8, 1), (2, 8, 2), (2, 7, 1), (3, 7, 1)])
def test_frame_sharded_decode_gloo(world, T, chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, chunks, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (T, 19, 3) for _, _, shape in res)
