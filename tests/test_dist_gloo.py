"""The N>1 path on CPU: world_size-2 gloo processes shard the frame with the library's tile
ownership rule and assemble it with the same sum-reduce the NCCL path uses."""
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


def _worker(rank, world, port, w, h, tile, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from mafrixraytracing_b200 import dist as mdist
    r, ws = mdist.init_process_group(backend="gloo")
    assert (r, ws) == (rank, world)
    pix = mdist.tile_pixels(w, h, tile, rank, world)
    frame = torch.zeros((h, w, 4), dtype=torch.float32)
    flat = frame.view(-1, 4)
    # stand-in for the per-rank render: a deterministic function of the absolute pixel id
    vals = torch.from_numpy(np.stack([pix * 0.5, pix % 7, pix // w, np.ones_like(pix)], 1).astype(np.float32))
    flat[torch.from_numpy(pix.astype(np.int64))] = vals
    mdist.reduce_frame(frame, dst=0)
    if rank == 0:
        q.put(frame.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,tile", [(96, 40, 16), (70, 33, 8)])
def test_two_rank_tile_shard_and_reduce(w, h, tile):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pix = np.arange(w * h)
    want = np.stack([pix * 0.5, pix % 7, pix // w, np.ones_like(pix)], 1).astype(np.float32).reshape(h, w, 4)
    assert np.array_equal(got, want)
