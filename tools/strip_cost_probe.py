"""Where does an N-way shard lose time?  Per-strip SM cycles of each emulated rank vs the unsharded frame.
python tools/strip_cost_probe.py [workload]   (RTM_COST_ORDER=1 so the unsharded frame records costs too)"""
import importlib, os, sys, numpy as np
os.environ.setdefault("RTM_COST_ORDER", "1")
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
WARPS = 148 * 32
for world in (1, 2, 4, 8):
    tot_cyc, tot_ms, rows = 0, 0.0, []
    for rank in range(world):
        ct.set_shard(rank, world)
        for i in range(4):
            ct.trace_tiles_async(frame); ct.sync()
        ms = ct.last_kernel_ms()
        cyc = ct.strip_cycles().astype(np.float64)
        busy_ms_at_1p9 = cyc.sum() / WARPS / 1.9e6  # if every warp slot were busy all the time at 1.9 GHz
        rows.append((ms, busy_ms_at_1p9, len(cyc), cyc.mean(), np.percentile(cyc, 99), cyc.max()))
        tot_cyc += cyc.sum(); tot_ms += ms
    r = np.array(rows)
    print("%s world=%d: kernel ms max %.3f sum %.3f | warp-busy ms (1.9 GHz) sum %.3f | strips/rank %d  cycles/strip mean %.0f p99 %.0f max %.0f"
          % (wl, world, r[:, 0].max(), tot_ms, r[:, 1].sum(), r[0, 2], r[:, 3].mean(), r[:, 4].mean(), r[:, 5].max()), flush=True)
