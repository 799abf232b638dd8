"""Per-rank kernel times for a sharded frame, with and without the IPC framebuffer (diagnostics)."""
import importlib, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
mr, capi, scenes, hostapi = pkg("multirank"), pkg("capi"), pkg("scenes"), pkg("hostapi")
wl = sys.argv[1] if len(sys.argv) > 1 else "killeroo4k"
scene, w, h, spp, res = scenes.CONFIGS[wl]
host = hostapi.host_api()
m, fov, cam = scenes.build(host, scene)
vtx, tri = m.arrays()
group = mr.RankGroup(dist, "cuda")
for share, signals in ((False, False), (True, False), (True, True)):
    ct = capi.CudaTrace(devices=[local])
    ct.upload_scene(vtx, tri, res)
    ct.set_shard(rank, world)
    if share:
        mr.share_framebuffer(ct, group, w, h)
    ct.set_shard_signals(signals)
    fov_xs, aspect = host.camera_constants(fov, w, h)
    frame = ct.make_frame(w, h, spp, cam, fov_xs, aspect)
    ms = []
    for i in range(8):
        group.barrier()
        ct.trace_tiles_async(frame)
        ct.sync()
        ms.append(ct.last_kernel_ms())
    allms = group.allreduce_sum(np.eye(world)[rank] * np.median(ms[3:]))
    if rank == 0:
        print(wl, "shared_fb" if share else "local_fb", "signals" if signals else "no-signals", "per-rank kernel ms:", np.round(allms, 3).tolist(), flush=True)
    group.barrier()
    ct.close()
dist.destroy_process_group()
