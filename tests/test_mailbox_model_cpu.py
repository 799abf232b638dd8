"""CPU check of the mailboxing mode's claim (the reference author's TODO at grid.cpp:172; product: csrc/rt_device.cuh,
CUDA_TRACE_VARIANT_MAILBOX): reusing a triangle's test outcome in later cells changes no result.  tools/study/rt_study.c
restates the walk WITH the product's mailbox (4 entries, round robin over the tests actually computed) on top of the
oracle's primitives; here it must (1) return what the oracle's plain walk returns, bit for bit, and (2) count exactly
the tests the GPU reported asking for, and the same share answered from the mailbox (tests/golden/mailbox_gpu_stats.json,
recorded on B200) -- the rays are the ones tests/test_gpu_parity.py::test_mailboxing_mode shoots."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

F = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def study():
    d = os.path.join(ROOT, "tools", "study")
    subprocess.check_call(["bash", os.path.join(d, "build.sh")])
    lib = C.CDLL(os.path.join(d, "librt_study.so"))
    lib.rts_mailbox_walk.restype = C.c_int
    lib.rts_mailbox_walk.argtypes = [C.c_void_p, F, F, C.c_int, F, F, F, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64)]
    return lib


@pytest.mark.parametrize("name,res", [("killeroo", 64), ("room", 64), ("torusknot", 64)])
def test_mailbox_walk_equals_plain_walk_and_gpu_counts(study, port, scene_data, name, res):
    with open(os.path.join(ROOT, "tests", "golden", "mailbox_gpu_stats.json")) as f:
        gpu = json.load(f)["%s_g%d" % (name, res)]
    sd = scene_data(name)
    ps = port.scene(sd.vtx, sd.tri, res, tight_ranges=True)
    rs = np.random.RandomState(11)
    n = 30000
    g = ps.grid()
    lo, hi = np.asarray(g["aabb_min"], np.float32), np.asarray(g["aabb_max"], np.float32)
    o = (rs.uniform(-1.0, 1.0, (n, 3)) * 2.0 * (hi - lo) + (lo + hi) / 2).astype(np.float32)
    target = rs.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    for variant in (0, 1):
        tri, t, u, v = ps.intersect_rays(o, d, variant)
        asked, reused = C.c_uint64(0), C.c_uint64(0)
        for i in range(n):
            ct, cu, cv, idx = C.c_float(0), C.c_float(0), C.c_float(0), C.c_uint32(0xFFFFFFFF)
            oi, di = np.ascontiguousarray(o[i]), np.ascontiguousarray(d[i])
            hit = study.rts_mailbox_walk(C.byref(ps.s), oi.ctypes.data_as(F), di.ctypes.data_as(F), variant, C.byref(ct),
                                         C.byref(cu), C.byref(cv), C.byref(idx), C.byref(asked), C.byref(reused))
            if hit:
                got = np.array([ct.value, cu.value, cv.value], np.float32).view(np.uint32)
                want = np.array([t[i], u[i], v[i]], np.float32).view(np.uint32)
                assert idx.value == tri[i] and np.array_equal(got, want), (name, variant, i)
            else:
                assert tri[i] == 0xFFFFFFFF, (name, variant, i)
        assert asked.value == gpu[str(variant)]["asked"]
        assert round(100.0 * reused.value / asked.value, 1) == gpu[str(variant)]["percent"]
