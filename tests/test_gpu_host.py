"""The C++ host mirror (Renderer : Framebuffer, Scene, Grid -- host/*.h) driven on one GPU and compared with the
unmodified reference's classes of the same names (oracle/_ref), or with the committed digests when it is absent.

Everything here goes Mesh -> Scene -> Renderer -> Resize/StartRendering -> WaitRendering -> tiles, i.e. the route a
viewer using the reference's API takes (INTEGRATION.md route A)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _host_renderer(name, grid_res=64):
    hostapi, scenes = pkg("hostapi"), pkg("scenes")
    m, fov, cam = scenes.build(hostapi.host_api(), name)
    return hostapi.HostRenderer(m, fov, cam, grid_res), fov, cam


def _ref_renderer(ref, name, grid_res=64):
    m, fov, cam = pkg("scenes").build(ref.api, name)
    return ref.renderer(m, fov, cam, grid_res)


@pytest.mark.parametrize("name,w,h,spp", [("cornell", 200, 120, 4), ("killeroo", 333, 217, 2), ("head", 128, 128, 16)])
def test_renderer_frame_matches_reference(ref, name, w, h, spp):
    hr, _, _ = _host_renderer(name)
    sec, img = hr.render(w, h, spp)
    assert sec > 0 and hr.last_kernel_ms() > 0
    _, want = _ref_renderer(ref, name).render(w, h, spp)
    assert np.array_equal(img, want)
    # a second frame at another size and sample count through the same Renderer (Resize path)
    _, img2 = hr.render(w // 2 + 1, h // 2 + 3, 1)
    _, want2 = _ref_renderer(ref, name).render(w // 2 + 1, h // 2 + 3, 1)
    assert np.array_equal(img2, want2)


def test_renderer_digest_without_reference():
    """Same route, checked against tests/golden/ref_digests.json -- runs on boxes without oracle/_ref."""
    with open(os.path.join(HERE, "golden", "ref_digests.json")) as f:
        digests = json.load(f)
    key = sorted(k for k, d in digests.items() if d["scene"] == "cornell")[0]
    d = digests[key]
    hr, _, _ = _host_renderer(d["scene"], d["grid_res"])
    _, img = hr.render(d["width"], d["height"], d["spp"])
    assert hashlib.md5(np.ascontiguousarray(img).tobytes()).hexdigest() == d["image_md5"]
    gi = hr.grid_info()
    assert [int(x) for x in gi["dim"]] == d["dim"] and gi["num_refs"] == d["refs"]
    assert int(np.float32(gi["cell_wdh"]).view(np.uint32)) == d["cell_wdh_bits"]


def test_save_bmp_byte_identical(ref, tmp_path):
    hr, _, _ = _host_renderer("cornell")
    hr.render(96, 64, 2)
    rr = _ref_renderer(ref, "cornell")
    rr.render(96, 64, 2)
    a, b = str(tmp_path / "host.bmp"), str(tmp_path / "ref.bmp")
    hr.save_bmp(a)
    rr.save_bmp(b)
    with open(a, "rb") as fa, open(b, "rb") as fb:
        assert fa.read() == fb.read()


def test_grid_intersect_and_ray_march_single_ray(ref):
    """Grid::Intersect (grid.cpp:208-334) and Renderer::RayMarch (renderer.cpp:24-41) as one-ray queries."""
    hr, _, _ = _host_renderer("cornell")
    rr = _ref_renderer(ref, "cornell")
    o3, d3 = rr.generate_rays(16, 12, 1, 0, 12)
    o, d = o3.reshape(-1, 3), d3.reshape(-1, 3)
    idx, t, u, v = rr.intersect_rays(o, d)
    mh, mt = rr.ray_march(o, d)
    hits = 0
    for i in range(0, len(o), 7):
        ok, tuv, tri = hr.intersect(o[i], d[i])
        assert ok == (idx[i] != 0xFFFFFFFF)
        if ok:
            hits += 1
            assert tri == idx[i]
            assert np.array_equal(tuv.view(np.uint32), np.array([t[i], u[i], v[i]], np.float32).view(np.uint32))
        got_hit, got_t = hr.ray_march(o[i], d[i])
        assert got_hit == bool(mh[i]) and np.float32(got_t).view(np.uint32) == mt[i].view(np.uint32)
    assert hits > 5


def test_renderer_alternates_match_port(port):
    """Renderer::SetOrthographicWidth / SetShadingMode through the C++ mirror == the port's restatement of
    camera.h:25-36 and renderer.cpp:116,118 (itself pinned to the reference on the CPU)."""
    hostapi, scenes = pkg("hostapi"), pkg("scenes")
    m, fov, cam = scenes.build(hostapi.host_api(), "cornell")
    vtx, tri = m.arrays()
    ps = port.scene(vtx, tri, 64)
    hr = hostapi.HostRenderer(m, fov, cam)
    w, h, spp = 150, 90, 4
    for ortho, mode in ((1.6, 0), (0.0, 1), (0.0, 2), (1.2, 2), (0.0, 0)):
        hr.set_alternates(ortho, mode)
        _, img = hr.render(w, h, spp)
        want = ps.render(cam, fov, w, h, spp, ortho_width=ortho if ortho else None, shade_mode=mode)["bgra"]
        assert np.array_equal(img, want), (ortho, mode)


def test_grid_info_matches_reference(ref):
    for name, res in (("killeroo", 64), ("room", 32)):
        hr, _, _ = _host_renderer(name, res)
        gi = hr.grid_info()
        g = _ref_renderer(ref, name, res).grid()
        assert np.array_equal(gi["dim"], g["dim"]) and gi["num_refs"] == len(g["tri_index"])
        assert np.float32(gi["cell_wdh"]).view(np.uint32) == np.float32(g["cell_wdh"]).view(np.uint32)
        for k in ("aabb_min", "aabb_max"):
            assert np.array_equal(gi[k].view(np.uint32), np.asarray(g[k], np.float32).view(np.uint32))
