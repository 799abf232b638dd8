"""CPU tests of the N > 1 path: the strip partition and the torch.distributed plumbing on the gloo
backend with world_size 2 (the data path itself -- IPC peer stores -- needs GPUs: tests/test_gpu_multi.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, pkg


@pytest.mark.parametrize("size,spp", [((512, 512), 1), ((1920, 1080), 4), ((640, 360), 16), ((131, 77), 16),
                                      ((131, 77), 1), ((8, 4), 64), ((7, 3), 3)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_covers_every_pixel_once(size, spp, world):
    mr, capi = pkg("multirank"), pkg("capi")
    w, h = size
    rects = capi.full_frame_tiles(w, h) if min(w, h) >= 16 else [(0, 0, w, h)]
    owner = mr.pixel_owner_map(w, h, rects, world, spp)
    sw, sh = mr.strip_size(spp, w * h * spp)
    assert owner.min() >= 0 and owner.max() <= world - 1
    counts = np.bincount(owner.ravel(), minlength=world)
    if w * h >= 10000:
        assert counts.min() > 0.7 * counts.mean()  # interleaving balances the pixel load
    prefix = mr.strip_prefix(rects, spp, w * h * spp)
    # strips never overlap and tile the rects exactly
    cover = np.zeros((h, w), np.int32)
    for s in range(prefix[-1]):
        x0, y0, x1, y1 = mr.strip_rect(rects, prefix, s, spp, w * h * spp)
        assert 0 < x1 - x0 <= sw and 0 < y1 - y0 <= sh
        cover[y0:y1, x0:x1] += 1
    assert (cover == 1).all()


WORKER = r'''
import os, sys, numpy as np, importlib
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
mr, capi, scenes, hostapi = pkg("multirank"), pkg("capi"), pkg("scenes"), pkg("hostapi")
from oracle import pyoracle as po
g = mr.RankGroup(dist, "cpu")
assert (g.rank, g.world) == (rank, world)
# 1. the 64-byte handle travels from rank 0 to everyone
payload = bytes(range(64)) if rank == 0 else b""
assert g.broadcast_bytes(payload, 0) == bytes(range(64))
# 2. max / sum over ranks
assert g.allreduce_max([1.0 + rank, 5.0 - rank]).tolist() == [float(world), 5.0]
assert g.allreduce_sum([1.0]).tolist() == [float(world)]
g.barrier()
# 3. every rank renders (with the CPU oracle standing in for the GPU) only the pixels of its own
#    strips; the sum over ranks must be the full frame: nothing missing, nothing rendered twice
host = hostapi.host_api()
m, fov, cam = scenes.build(host, "cornell")
vtx, tri = m.arrays()
w, h, spp = 96, 64, 2
full = po.Port.get().scene(vtx, tri, 64, n_threads=2).render(cam, fov, w, h, spp, n_threads=2)["bgra"]
owner = mr.pixel_owner_map(w, h, capi.full_frame_tiles(w, h), world, spp)
mine = np.where(owner == rank, full, 0).astype(np.float64)
total = g.allreduce_sum(mine)
assert np.array_equal(total.astype(np.uint32), full)
dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank)
'''


def free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RTM_QUIET="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


@pytest.mark.parametrize("size,spp", [((3840, 2160), 16), ((1920, 1080), 4), ((512, 512), 1), ((333, 217), 2), ((640, 360), 64)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_band_completion_targets_match_the_strip_partition(size, spp, world):
    """What cuda_trace_tiles programs into the kernel and waits for (per-GPU pieces of strips per row band, host
    arithmetic in csrc/api.cu) against the strip partition restated in multirank.py: every strip counts its pieces in
    the band of its first and, when it straddles a boundary, of its last row, for the rank that owns it."""
    capi, mr = pkg("capi"), pkg("multirank")
    w, h = size
    rects = capi.full_frame_tiles(w, h)
    got = capi.band_shares(w, h, spp, rects, world)
    rays = w * h * spp
    sw, sh = mr.strip_size(spp, rays)
    assert got["band_rows"] >= sh and got["n_bands"] == -(-h // got["band_rows"]) and 1 <= got["n_bands"] <= 32
    assert got["pieces_per_strip"] in (1, 2, 4)
    want = np.zeros((world, 32), np.int64)
    prefix = mr.strip_prefix(rects, spp, rays)
    for tile, (x0, y0, x1, y1) in enumerate(rects):
        nx = -(-(x1 - x0) // sw)
        for row, y in enumerate(range(y0, y1, sh)):
            b0, b1 = y // got["band_rows"], (min(y + sh, y1) - 1) // got["band_rows"]
            ids = prefix[tile] + row * nx + np.arange(nx)
            owners = np.array([mr.strip_owner(int(s), world) for s in ids[:: max(1, mr.SHARD_CHUNK // 4)]])  # spot ...
            c = ids // mr.SHARD_CHUNK
            own = (c % world + c // world) % world                                                       # ... and all
            assert np.array_equal(own[:: max(1, mr.SHARD_CHUNK // 4)], owners)
            np.add.at(want[:, b0], own, got["pieces_per_strip"])
            if b1 != b0:
                np.add.at(want[:, b1], own, got["pieces_per_strip"])
    assert np.array_equal(got["shares"].astype(np.int64), want)
    assert np.array_equal(got["gpus_in_band"], (want > 0).sum(axis=0))
    assert want[:, got["n_bands"]:].sum() == 0 and want.sum() >= prefix[-1] * got["pieces_per_strip"]

