"""One process per GPU (torchrun): the frame sharded over ranks, scene replicated.

Ownership is by COLUMN STRIPES (MFX_SAMPLE_STRIPES: columns [c*S, (c+1)*S) -> rank c % world, computed arithmetically in
the kernels).  Every rank renders its stripes into the reference's own frame layout on its GPU -- Color[w,h], x-major f64
(Texture.fs:21-28), where a stripe is one contiguous block -- and ONE gather over NCCL/NVLink brings the owned stripes
to rank 0: 1/world of the frame per rank instead of the full-frame sum-reduce of round 1.  The counter-based RNG is keyed
on the absolute pixel / sample, so the assembled frame is bit-identical for any world size.

The reference has no distributed code (its only parallelism is Array.Parallel.iter over pixels, Integrators.fs:164).
A single-process host (the reference's own shape: one render thread) uses mfx_multi_sample instead -- same stripes, the
devices write straight into the host texture and no collective is involved at all.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .scene import CudaPixelIntegrator, FAST_F32

# 16-column stripes: 120 stripes of a 1080p frame balance 8 ranks to 0.1 % (rays per stripe counted by the oracle) and a
# stripe row is half a warp of neighbouring pixels.  (Round 1's 16x16 tiles own exactly these columns when the number of
# tile columns is a multiple of the world size.)
STRIPE = int(os.environ.get("MFX_STRIPE", "16"))
TILE = STRIPE       # round-1 name


def tile_pixels(width, height, tile, rank, world):
    """This rank's linear pixel ids (y*width+x) under square-tile ownership (mfx_tile_map)."""
    n = C.c_int32()
    lib = _lib.load()
    _lib.check(lib.mfx_tile_map(width, height, tile, rank, world, None, C.byref(n)))
    out = np.zeros(n.value, np.int32)
    _lib.check(lib.mfx_tile_map(width, height, tile, rank, world, _lib.ptr(out), C.byref(n)))
    return out


def stripe_pixels(width, height, stripe, rank, world):
    """This rank's linear pixel ids under column-stripe ownership, in the kernels' own order (mfx_stripe_map)."""
    n = C.c_int32()
    lib = _lib.load()
    _lib.check(lib.mfx_stripe_map(width, height, stripe, rank, world, None, C.byref(n)))
    out = np.zeros(n.value, np.int32)
    _lib.check(lib.mfx_stripe_map(width, height, stripe, rank, world, _lib.ptr(out), C.byref(n)))
    return out


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend="nccl"):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


class StripeGather:
    """Gathers the owned stripes of an x-major frame tensor (width, height, C) onto rank `dst`.  Rank r owns stripes
    r, r+world, ...; it sends them packed (ceil(n_stripes / world) slots, the unused tail slot stays zero) and the
    destination writes each rank's slots back at their stripe positions.  Buffers are allocated once."""

    def __init__(self, frame, stripe, rank, world, dst=0):
        import torch
        self.torch, self.rank, self.world, self.dst = torch, rank, world, dst
        w = frame.shape[0]
        if w % stripe:
            raise ValueError(f"StripeGather needs the width ({w}) to be a multiple of the stripe width ({stripe})")
        self.n_stripes = w // stripe
        self.per = -(-self.n_stripes // world)
        self.view = frame.view(self.n_stripes, -1)                     # one row per stripe (contiguous in x-major)
        self.send = torch.zeros((self.per, self.view.shape[1]), dtype=frame.dtype, device=frame.device)
        # one receive buffer: rank r's packed stripes are row r, so when the stripes divide evenly the destination puts
        # them all back with ONE strided copy (stripe k*world + r <- recv[r, k]) instead of one copy per rank
        self.recv_all = torch.empty((world,) + tuple(self.send.shape), dtype=frame.dtype, device=frame.device) if rank == dst else None
        self.recv = [self.recv_all[r] for r in range(world)] if rank == dst else None
        self.even = (self.n_stripes % world == 0)

    def __call__(self):
        import torch.distributed as dist
        if self.world <= 1:
            return
        mine = self.view[self.rank::self.world]
        self.send[:mine.shape[0]].copy_(mine)
        dist.gather(self.send, self.recv, dst=self.dst)
        if self.rank == self.dst and self.even:
            self.view.view(self.per, self.world, -1).copy_(self.recv_all.transpose(0, 1))     # (its own row holds its own stripes)
        elif self.rank == self.dst:
            for r in range(self.world):
                if r == self.dst:
                    continue
                k = len(range(r, self.n_stripes, self.world))
                self.view[r::self.world].copy_(self.recv[r][:k])


class ShardedPixelIntegrator:
    """IPixelIntegrator over `world` processes: Sample(n) renders this rank's stripes and gathers them.
    Returns the (width, height, 4) float64 CUDA tensor = Color[w,h] on the device (complete on rank 0)."""

    def __init__(self, scene, rank, world, precision=FAST_F32, seed=1, stripe=STRIPE, device=None, tile=None):
        import torch
        self.torch = torch
        self.scene, self.rank, self.world = scene, rank, world
        stripe = stripe if tile is None else tile
        self.flags = (_lib.SAMPLE_STRIPES | _lib.SAMPLE_NO_CLEAR) if world > 1 else 0
        self.integ = CudaPixelIntegrator(scene, precision=precision, seed=seed,
                                         tile_size=stripe if world > 1 else 0, rank=rank, world=world)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.frame = torch.zeros((scene.width, scene.height, 4), dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)      # the library writes the frame on its own stream: the fill must have landed
        self.gather = StripeGather(self.frame, stripe, rank, world) if world > 1 else None

    def Sample(self, n, first_sample=0, gather=True, flags=0):
        self.integ.SampleDeviceColor(n, self.frame.data_ptr(), first_sample=first_sample, flags=self.flags | flags)
        if gather and self.gather is not None:
            self.gather()
        return self.frame

    @property
    def stats(self):
        return self.integ.stats
