"""One-process-per-GPU plumbing (torchrun): how a frame is split over ranks and how the ranks meet.

The data path has no collective: every rank holds the replicated scene + grid, renders its interleaved share
of the strips (``strip_owner``; cuda_trace_set_shard) and stores its pixels straight into rank 0's
device framebuffer, which the other ranks map through a CUDA-IPC handle (NVLink peer stores).
torch.distributed is only used for the rendezvous: broadcasting the 64-byte handle, barriers, and
the max-over-ranks of the timings.  The same helpers run on the gloo backend for CPU tests.

The strip partition below is the host-side statement of what csrc/trace_kernels.cu does on the
device (strips of about 128 rays, tile after tile, strips row-major inside a tile).
"""
import numpy as np

def strip_size(spp, frame_rays=1 << 30):
    """(w, h) of a strip: about 128 rays per strip (32 for frames below 16 M rays), one of 2x1 2x2
    4x2 4x4 8x4 (strip_size_for_spp in csrc/trace_kernels.cuh)."""
    target = 128 if frame_rays >= (16 << 20) else 32
    pixels = 32
    while pixels > 2 and pixels * spp > target:
        pixels //= 2
    return (8 if pixels >= 32 else 4 if pixels >= 8 else 2), (4 if pixels >= 16 else 2 if pixels >= 4 else 1)


def strip_prefix(rects, spp=1, frame_rays=1 << 30):
    """First global strip id of every tile (+ total), exactly as api.cu builds it."""
    sw, sh = strip_size(spp, frame_rays)
    prefix = [0]
    for x0, y0, x1, y1 in rects:
        nx = (x1 - x0 + sw - 1) // sw
        ny = (y1 - y0 + sh - 1) // sh
        prefix.append(prefix[-1] + nx * ny)
    return prefix


def strip_rect(rects, prefix, strip, spp=1, frame_rays=1 << 30):
    """Pixel rectangle (x0, y0, x1, y1) of global strip id ``strip``."""
    sw, sh = strip_size(spp, frame_rays)
    tile = int(np.searchsorted(prefix, strip, side="right") - 1)
    x0, y0, x1, y1 = rects[tile]
    local = strip - prefix[tile]
    nx = (x1 - x0 + sw - 1) // sw
    bx0 = x0 + (local % nx) * sw
    by0 = y0 + (local // nx) * sh
    return bx0, by0, min(bx0 + sw, x1), min(by0 + sh, y1)


SHARD_CHUNK = 32


def strip_owner(strip, world, chunk=SHARD_CHUNK):
    """Rank that renders global strip ``strip``: chunks of ``chunk`` consecutive strips are dealt
    round-robin over the ranks, the owner rotating by one from round to round (trace_kernels.cu)."""
    c = strip // chunk
    return (c % world + c // world) % world


def pixel_owner_map(width, height, rects, world, spp=1):
    """[H, W] int32: rank that renders each pixel, -1 where no tile covers it."""
    owner = np.full((height, width), -1, np.int32)
    rays = width * height * spp
    prefix = strip_prefix(rects, spp, rays)
    for s in range(prefix[-1]):
        x0, y0, x1, y1 = strip_rect(rects, prefix, s, spp, rays)
        owner[y0:y1, x0:x1] = strip_owner(s, world)
    return owner


class RankGroup:
    """Thin wrapper over torch.distributed for the three things the bench needs."""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device  # torch device for collectives (cuda:N under nccl, cpu under gloo)
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def broadcast_bytes(self, payload, src=0):
        """Broadcast a small bytes object (the CUDA IPC handle) from ``src`` to every rank."""
        if self.dist is None:
            return payload
        import torch
        n = torch.tensor([len(payload) if self.rank == src else 0], dtype=torch.int64, device=self.device)
        self.dist.broadcast(n, src)
        buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=self.device)
        if self.rank == src:
            buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        self.dist.broadcast(buf, src)
        return bytes(buf.cpu().numpy().tobytes())

    def allreduce_max(self, values):
        """Element-wise max over ranks of a float array (device timings)."""
        a = np.asarray(values, np.float64)
        if self.dist is None:
            return a
        import torch
        t = torch.tensor(a, dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def allreduce_sum(self, values):
        a = np.asarray(values, np.float64)
        if self.dist is None:
            return a
        import torch
        t = torch.tensor(a, dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


def share_framebuffer(ct, group, width, height):
    """Rank 0 allocates + exports its device framebuffer, everyone else imports it."""
    if group.world == 1:
        return
    handle = b""
    if group.rank == 0:
        ct.prepare_framebuffer(width, height)
        handle = ct.export_framebuffer()
    handle = group.broadcast_bytes(handle, 0)
    if group.rank != 0:
        ct.import_framebuffer(handle, width, height)
    group.barrier()
