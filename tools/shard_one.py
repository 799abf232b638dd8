"""Render rank R's share of a WORLD-way sharded frame a few times on one GPU (for ncu).  args: workload world rank"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
capi, scenes, hostapi = pkg("capi"), pkg("scenes"), pkg("hostapi")
wl, world, rank = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene, w, h, spp, res = scenes.CONFIGS[wl]
host = hostapi.host_api()
m, fov, cam = scenes.build(host, scene)
vtx, tri = m.arrays()
ct = capi.CudaTrace(1)
ct.upload_scene(vtx, tri, res)
ct.set_shard(rank, world)
fov_xs, aspect = host.camera_constants(fov, w, h)
frame = ct.make_frame(w, h, spp, cam, fov_xs, aspect)
for i in range(5):
    ct.trace_tiles_async(frame); ct.sync(); print("ms", ct.last_kernel_ms())
