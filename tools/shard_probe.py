"""Emulate an N-way sharded frame on ONE GPU: render each rank's share in turn and report the
per-rank kernel times (load balance of the strip partition).  python tools/shard_probe.py [workload]"""
import importlib, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
capi, scenes, hostapi = pkg("capi"), pkg("scenes"), pkg("hostapi")
wl = sys.argv[1] if len(sys.argv) > 1 else "killeroo4k"
scene, w, h, spp, res = scenes.CONFIGS[wl]
host = hostapi.host_api()
m, fov, cam = scenes.build(host, scene)
vtx, tri = m.arrays()
ct = capi.CudaTrace(1)
ct.upload_scene(vtx, tri, res)
fov_xs, aspect = host.camera_constants(fov, w, h)
frame = ct.make_frame(w, h, spp, cam, fov_xs, aspect)
for world in (1, 2, 4, 8):
    times = []
    for rank in range(world):
        ct.set_shard(rank, world)
        ms = []
        for i in range(5):
            ct.trace_tiles_async(frame); ct.sync(); ms.append(ct.last_kernel_ms())
        times.append(min(ms[1:]))
    t = np.array(times)
    print("%s chunk=%s world=%d: max %.3f ms  mean %.3f  sum %.3f  per-rank %s" % (wl, os.environ.get("RTM_SHARD_CHUNK", "32"), world, t.max(), t.mean(), t.sum(), np.round(t, 3).tolist()), flush=True)
