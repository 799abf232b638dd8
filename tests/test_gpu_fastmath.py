"""GPU check of the range-check-free arithmetic the trace kernel uses for operands of ordinary magnitude
(csrc/rt_device.cuh: rcp_normal, div_normal, sqrt_normal): the same bits as the IEEE intrinsics
__frcp_rn / __fdiv_rn / __fsqrt_rn -- which are what the reference's `1.0f / x`, `a / b` and `sqrt` compile to on
x86-64 (aabb.h:49-51, grid.cpp:199-214, lin_alg.h:151-156) -- over 2^32 pseudo-random operands."""
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exp_lo,exp_hi,seed", [(-40, 2, 1), (-40, 40, 2), (-20, 21, 3), (-1, 1, 4)])
def test_fast_sequences_equal_ieee_intrinsics(exp_lo, exp_hi, seed):
    # exponent windows: direction components (2^-40 .. 1) against cell-sized numerators; the launcher's full
    # validity window; frame constants (generate_ray); the densest case, operands of the same magnitude
    bad = pkg("capi").check_fast_arith(1 << 30, seed, exp_lo, exp_hi)
    assert bad == (0, 0, 0), "mismatches (rcp, div, sqrt) = %r" % (bad,)


def test_frames_with_and_without_fast_math_are_identical(monkeypatch, scene_data):
    """RTM_FAST_MATH=0 forces the intrinsics everywhere: images and per-sample records must not change."""
    import numpy as np
    capi, hostapi = pkg("capi"), pkg("hostapi")
    sd = scene_data("killeroo")
    w, h, spp = 320, 200, 4
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("RTM_FAST_MATH", flag)
        ct = capi.CudaTrace(1)
        ct.upload_scene(sd.vtx, sd.tri, 64)
        fov_xs, aspect = hostapi.host_api().camera_constants(sd.fov, w, h)
        img = ct.trace_tiles(ct.make_frame(w, h, spp, sd.cam16, fov_xs, aspect, keep_hits=True))
        out.append((img, ) + tuple(ct.download_hits(w, h, spp)))
        ct.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
