"""Multi-GPU tests (need >= 2 devices; skipped otherwise): the tile-sharded frame must be the
same image, bit for bit, as the single-GPU frame."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, pkg

pytestmark = pytest.mark.gpu


def n_devices():
    return pkg("capi").load_library().cuda_trace_device_count()


@pytest.fixture(scope="module")
def two_gpus():
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.parametrize("name,size,spp", [("cornell", (200, 120), 4), ("killeroo", (320, 180), 16), ("room", (131, 77), 1)])
def test_single_process_two_devices(two_gpus, port, scene_data, name, size, spp):
    capi = pkg("capi")
    sd = scene_data(name)
    w, h = size
    fov_xs, aspect = port.camera_constants(sd.fov, w, h)
    imgs, hits = [], []
    for n in (1, 2):
        ct = capi.CudaTrace(n)
        ct.upload_scene(sd.vtx, sd.tri, 64)
        f = ct.make_frame(w, h, spp, sd.cam16, fov_xs, aspect, keep_hits=True)
        imgs.append(ct.trace_tiles(f))
        hits.append(ct.download_hits(w, h, spp))
        ct.close()
    assert np.array_equal(imgs[0], imgs[1])
    for a, b in zip(hits[0], hits[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    o = port.scene(sd.vtx, sd.tri, 64).render(sd.cam16, sd.fov, w, h, spp, want_hits=True)
    assert np.array_equal(hits[1][0], o["tri"]) and np.array_equal(imgs[1], o["bgra"])


def test_host_renderer_two_devices(two_gpus, port):
    hostapi, scenes = pkg("hostapi"), pkg("scenes")
    host = hostapi.host_api()
    m, fov, cam = scenes.build(host, "torusknot")
    vtx, tri = m.arrays()
    r = hostapi.HostRenderer(m, fov, cam, 64, n_gpus=2)
    _, img = r.render(256, 144, 4)
    r.close()
    o = port.scene(vtx, tri, 64).render(cam, fov, 256, 144, 4)
    assert np.array_equal(img, o["bgra"])


WORKER = r'''
import os, sys, numpy as np, importlib
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
mr, capi, scenes, hostapi = pkg("multirank"), pkg("capi"), pkg("scenes"), pkg("hostapi")
host = hostapi.host_api()
m, fov, cam = scenes.build(host, "killeroo")
vtx, tri = m.arrays()
w, h, spp = 384, 216, 4
ct = capi.CudaTrace(devices=[local])
ct.upload_scene(vtx, tri, 64)
ct.set_shard(rank, world)
group = mr.RankGroup(dist, "cuda")
mr.share_framebuffer(ct, group, w, h)
fov_xs, aspect = host.camera_constants(fov, w, h)
frame = ct.make_frame(w, h, spp, cam, fov_xs, aspect)
for _ in range(3):
    ct.trace_tiles_async(frame)
    ct.sync()
    group.barrier()
if rank == 0:
    img = np.zeros((h, w), np.uint32)
    ct.read_framebuffer(img)
    one = capi.CudaTrace(devices=[local])
    one.upload_scene(vtx, tri, 64)
    ref_img = one.trace_tiles(frame)
    one.close()
    assert np.array_equal(img, ref_img), "sharded frame differs from the single-GPU frame"
    print("sharded frame ok", int((img != 0).sum()))
group.barrier()
# overlapped gather: every rank signals finished strips per row band; rank 0's trace_tiles() copies
# each band to the host as soon as all ranks' strips of it are in (no barrier before the read-back)
ct.set_shard_signals(True)
pinned = capi.PinnedImage(w, h)
for it in range(4):
    pinned.array[:] = 0
    group.barrier()
    if rank == 0:
        ct.trace_tiles(frame, out=pinned.array)
        assert np.array_equal(pinned.array, ref_img), "overlapped gather differs (iteration %d)" % it
    else:
        ct.trace_tiles_async(frame)
        ct.sync()
    group.barrier()
if rank == 0:
    print("overlapped gather ok")
pinned.close()
ct.close()
dist.destroy_process_group()
'''


def free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_one_process_per_gpu_ipc_gather(two_gpus, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RTM_QUIET="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded frame ok" in r.stdout and "overlapped gather ok" in r.stdout
