"""CPU model of K1's warp-level traversal (no GPU): where do the lanes of a warp wait?

For a sample of strips of a frame, the oracle port records for every primary ray the triangle-list lengths of the
cells it visits (tools/study/rt_study.c rts_ray_walk_profile, a study copy of the oracle's walk).  The model then replays K1's two-phase loop per warp round (32 rays,
lane = pixel-in-round * spp + sample): phase A costs the LONGEST empty-cell run among the lanes, phase B the LONGEST
list, and compares with (a) the useful work, i.e. perfect packing, and (b) re-pairing the strip's rays between phases
(the rays of one strip pooled, sorted by what they need next, and dealt to the strip's warps).

python tools/warp_walk_model.py [workload] [n_strips]        (test infrastructure: imports oracle/)"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

pkg = lambda s: importlib.import_module("cpp-11-ray-trace-march-framework_b200." + s)
scenes, hostapi, mr = pkg("scenes"), pkg("hostapi"), pkg("multirank")

C_DDA, C_TEST, C_ITER = 13, 61, 24  # SASS instructions: one DDA step, one triangle test, fixed cost of one A+B iteration

wl = sys.argv[1] if len(sys.argv) > 1 else "killeroo4k"
n_strips = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
scene, w, h, spp, res = scenes.CONFIGS[wl]
host = hostapi.host_api()
m, fov, cam = scenes.build(host, scene)
vtx, tri = m.arrays()
port = pyoracle.Port.get()
ps = port.scene(vtx, tri, res)
lib = port.lib
F, U16 = C.POINTER(C.c_float), C.POINTER(C.c_uint16)
import subprocess  # noqa: E402
STUDY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "study")
subprocess.check_call(["bash", os.path.join(STUDY, "build.sh")])
study = C.CDLL(os.path.join(STUDY, "librt_study.so"))
study.rts_ray_walk_profile.restype = C.c_uint32
study.rts_ray_walk_profile.argtypes = [C.c_void_p, F, F, C.c_uint32, U16, C.POINTER(C.c_int)]
smp = port.sample_table(spp)
fov_xs, aspect = port.camera_constants(fov, w, h)
cam32 = np.ascontiguousarray(cam, np.float32)
sw, sh = mr.strip_size(spp, w * h * spp)
rs = np.random.RandomState(7)
prof_buf = np.zeros(4096, np.uint16)


def ray_profile(px, py, s):
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.rto_generate_ray(cam32.ctypes.data_as(F), px, py, w, h, float(smp[s, 0]), float(smp[s, 1]), float(fov_xs),
                         float(aspect), o.ctypes.data_as(F), d.ctypes.data_as(F))
    hit = C.c_int(0)
    n = study.rts_ray_walk_profile(C.byref(ps.s), o.ctypes.data_as(F), d.ctypes.data_as(F), len(prof_buf),
                                 prof_buf.ctypes.data_as(U16), C.byref(hit))
    return prof_buf[:min(n, len(prof_buf))].astype(np.int64).copy()


def segments(profile):
    """-> list of (empty_steps_before, list_len) per occupied cell visited, + trailing empty steps to the exit.
    K1 steps once after every tested cell, so the step onto the next cell belongs to the following phase A."""
    segs, run = [], 0
    for k, ln in enumerate(profile):
        if ln == 0:
            run += 1
        else:
            segs.append((run + (1 if segs else 0), int(ln)))
            run = 0
    return segs, run + (1 if segs else 0)


def model_warp(rays):
    """rays: list of (segs, tail).  -> (cost of K1's loop, useful work), in thread-instruction units / 32"""
    idx = [0] * len(rays)
    alive = [True] * len(rays)
    cost = useful = 0
    while any(alive):
        a_steps, b_len = [], []
        for r, (segs, tail) in enumerate(rays):
            if not alive[r]:
                continue
            if idx[r] < len(segs):
                a_steps.append(segs[idx[r]][0]); b_len.append(segs[idx[r]][1]); idx[r] += 1
                if idx[r] == len(segs) and tail == 0:
                    alive[r] = False
            else:
                a_steps.append(tail); b_len.append(0); alive[r] = False
        # a ray whose last occupied cell holds its hit ends there (profile ends with that cell: tail == 0)
        cost += C_ITER + max(a_steps) * C_DDA + max(b_len) * C_TEST
        useful += (sum(a_steps) * C_DDA + sum(b_len) * C_TEST) / 32.0
    return cost, useful


def model_pooled(rays, n_warps):
    """All rays of the strip advance in lock step; before each phase the pool is sorted by what the rays need next and
    dealt to n_warps warps of 32: each phase costs the sum over warps of their longest lane."""
    idx = [0] * len(rays)
    alive = [True] * len(rays)
    cost = 0
    while any(alive):
        a_steps, b_len = [], []
        for r, (segs, tail) in enumerate(rays):
            if not alive[r]:
                continue
            if idx[r] < len(segs):
                a_steps.append(segs[idx[r]][0]); b_len.append(segs[idx[r]][1]); idx[r] += 1
                if idx[r] == len(segs) and tail == 0:
                    alive[r] = False
            else:
                a_steps.append(tail); b_len.append(0); alive[r] = False
        for vals, c in ((sorted(a_steps, reverse=True), C_DDA), (sorted(b_len, reverse=True), C_TEST)):
            for k in range(0, len(vals), 32):
                cost += vals[k] * c
        cost += C_ITER * ((len(a_steps) + 31) // 32) + 40  # + a shared-memory exchange per phase pair
    return cost


tot = dict(k1=0.0, useful=0.0, pooled=0.0)
n_x, n_y = w // sw, h // sh
for _ in range(n_strips):
    bx, by = rs.randint(n_x) * sw, rs.randint(n_y) * sh
    rays = []
    for slot in range(sw * sh):  # Morton slot order of the strip's pixels, lanes sample-fastest (trace_kernels.cu)
        ox = (slot & 1) | ((slot >> 1) & 2) | ((slot >> 2) & 4)
        oy = ((slot >> 1) & 1) | ((slot >> 2) & 2)
        for s in range(spp):
            rays.append(segments(ray_profile(bx + ox, by + oy, s)))
    warps = [rays[k:k + 32] for k in range(0, len(rays), 32)]
    for wr in warps:
        c, u = model_warp(wr)
        tot["k1"] += c
        tot["useful"] += u
    tot["pooled"] += model_pooled(rays, len(warps))
print("%s: %d strips of %dx%d px x %d spp (%d rays)" % (wl, n_strips, sw, sh, spp, n_strips * sw * sh * spp))
print("modelled warp-instructions per ray: K1 loop %.0f   useful (perfect packing) %.0f   strip-pooled re-pairing %.0f"
      % (tot["k1"] * 32 / (n_strips * sw * sh * spp) / 32 * 32, tot["useful"] * 32 / (n_strips * sw * sh * spp),
         tot["pooled"] * 32 / (n_strips * sw * sh * spp)))
print("lane utilisation of the K1 loop %.1f %%; re-pairing within a strip would cut the loop by %.1f %%"
      % (100 * tot["useful"] / tot["k1"], 100 * (1 - tot["pooled"] / tot["k1"])))
