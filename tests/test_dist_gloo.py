"""The N>1 path on CPU: world_size-2 (and 3) gloo processes shard the frame with the library's column-stripe
ownership rule (mfx_stripe_map) and assemble it with the same stripe gather the NCCL path uses."""
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


def _worker(rank, world, port, w, h, stripe, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from mafrixraytracing_b200 import dist as mdist
    r, ws = mdist.init_process_group(backend="gloo")
    assert (r, ws) == (rank, world)
    pix = mdist.stripe_pixels(w, h, stripe, rank, world)        # the library's ownership rule, in the kernels' own order
    frame = torch.zeros((w, h, 4), dtype=torch.float64)         # Color[w,h], x-major like the reference's Texture2D
    # stand-in for the per-rank render: a deterministic function of the absolute pixel id
    x, y = pix % w, pix // w
    vals = torch.from_numpy(np.stack([pix * 0.5, pix % 7, pix // w, np.ones_like(pix)], 1).astype(np.float64))
    frame[torch.from_numpy(x.astype(np.int64)), torch.from_numpy(y.astype(np.int64))] = vals
    mdist.StripeGather(frame, stripe, rank, world)()
    if rank == 0:
        q.put(frame.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,stripe,world", [(96, 40, 16, 2), (72, 33, 8, 2), (80, 17, 16, 3)])
def test_stripe_shard_and_gather(w, h, stripe, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, stripe, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    pix = np.arange(w * h)
    want = np.stack([pix * 0.5, pix % 7, pix // w, np.ones_like(pix)], 1).astype(np.float64).reshape(h, w, 4).transpose(1, 0, 2)
    assert np.array_equal(got, want)
