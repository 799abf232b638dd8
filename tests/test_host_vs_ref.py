"""CPU tests: the host library's Mesh / Matrix44f mirror against the unmodified reference
(oracle/_ref), bit for bit.  Skipped where the reference could not be built."""
import os

import numpy as np
import pytest

from conftest import pkg


@pytest.fixture(scope="module")
def host():
    return pkg("hostapi").host_api()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_matrix_builders(host, ref):
    r = ref.api
    for deg in (0.0, 30.0, 45.0, 90.0, -17.5, 123.456, 360.0):
        assert np.array_equal(bits(host.rotation_x(deg)), bits(r.rotation_x(deg)))
        assert np.array_equal(bits(host.rotation_y(deg)), bits(r.rotation_y(deg)))
        assert np.array_equal(bits(host.rotation_z(deg)), bits(r.rotation_z(deg)))
    assert np.array_equal(bits(host.identity()), bits(r.identity()))
    assert np.array_equal(bits(host.scaling(0.25)), bits(r.scaling(0.25)))
    assert np.array_equal(bits(host.translation(-0.065, -0.1, 0.05)), bits(r.translation(-0.065, -0.1, 0.05)))
    for eye, at in [((-1.00001, 1.0, 1.0), (0.0, -0.2, 0.0)), ((0, 0, -2), (0, 0, 0)), ((-0.47, 0.15, -0.3), (1.0, -0.7, 0.9)),
                    ((-1.6, 1.2, -1.0), (0.0, 0.0, -0.1)), ((0.5, 0.5, 0.0), (0, 0, 0))]:
        assert np.array_equal(bits(host.look_at(eye, at)), bits(r.look_at(eye, at)))


def test_matrix_multiply_and_invert_random(host, ref):
    rs = np.random.RandomState(3)
    for i in range(300):
        a = rs.normal(size=16).astype(np.float32) * np.float32(10.0 ** rs.randint(-3, 3))
        b = rs.normal(size=16).astype(np.float32)
        assert np.array_equal(bits(host.multiply(a, b)), bits(ref.api.multiply(a, b)))
        ok_h, inv_h = host.invert(a)
        ok_r, inv_r = ref.api.invert(a)
        assert ok_h == ok_r
        assert np.array_equal(bits(inv_h), bits(inv_r))
    sing = np.zeros(16, np.float32)
    assert host.invert(sing)[0] is False and ref.api.invert(sing)[0] is False


def test_camera_constants(host, ref, port):
    for fov in (30.0, 45.0, 51.0, 60.0, 75.0, 90.0, 1.0, 179.0):
        for w, h in ((512, 512), (1920, 1080), (3840, 2160), (67, 45)):
            a = host.camera_constants(fov, w, h)
            b = ref.api.camera_constants(fov, w, h)
            c = port.camera_constants(fov, w, h)
            assert bits(a[0]) == bits(b[0]) == bits(c[0]) and bits(a[1]) == bits(b[1]) == bits(c[1])


def test_cornell_box_and_quads(host, ref):
    hv, ht = host.mesh().cornell_box().arrays()
    rv, rt = ref.api.mesh().cornell_box().arrays()
    assert np.array_equal(hv.view(np.uint32), rv.view(np.uint32)) and np.array_equal(ht, rt)
    assert len(ht) == 32 and len(hv) == 64


@pytest.mark.parametrize("name", list(range(10)) + ["tiger_soup_small"])
def test_scene_presets_bitwise(host, ref, name):
    scenes = pkg("scenes")
    hm, hf, hc = scenes.build(host, name)
    rm, rf, rc = scenes.build(ref.api, name)
    assert hf == rf and np.array_equal(bits(hc), bits(rc))
    hv, ht = hm.arrays()
    rv, rt = rm.arrays()
    assert np.array_equal(ht, rt)
    assert np.array_equal(hv.view(np.uint32), rv.view(np.uint32))
    ha, hb = hm.compute_aabb()
    ra, rb = rm.compute_aabb()
    assert np.array_equal(bits(ha), bits(ra)) and np.array_equal(bits(hb), bits(rb))


REF_MESH_DIR = "/root/reference/meshes"


@pytest.mark.skipif(not os.path.isdir(REF_MESH_DIR), reason="reference meshes not present")
def test_ascii_reader_matches_reference_and_assets(host, ref):
    meshapi = pkg("meshapi")
    names = sorted(f[:-4] for f in os.listdir(REF_MESH_DIR) if f.endswith(".dat"))
    assert len(names) == 17
    for name in names:
        for flip in ((False, True) if name == "table_chair" else (False,)):
            path = os.path.join(REF_MESH_DIR, name + ".dat")
            hm, rm = host.mesh(), ref.api.mesh()
            assert hm.read_file(path, flip) and rm.read_file(path, flip)
            hv, ht = hm.arrays()
            rv, rt = rm.arrays()
            assert np.array_equal(ht, rt), name
            assert np.array_equal(hv.view(np.uint32), rv.view(np.uint32)), name
            av, at = meshapi.load_meshbin(name, flip)
            assert np.array_equal(at, rt) and np.array_equal(av.view(np.uint32), rv.view(np.uint32)), name


def test_float_parsing_matches_reference_on_hard_decimals(host, ref, tmp_path):
    """The host reader parses plain decimals on a fast exact path and leaves the rest to strtof (host/mesh.cpp);
    the reference uses fscanf("%f").  Positions written in every notation that occurs in practice plus the hard
    cases: float rounding midpoints spelled out exactly, one digit above / below them, 17+ significant digits,
    exponents, subnormals, signed zeros, huge values.  Non-indexed position-only mesh: three floats per line."""
    rs = np.random.RandomState(11)
    vals = []
    f32 = rs.standard_normal(600).astype(np.float32) * np.float32(10.0) ** rs.randint(-6, 7, 600).astype(np.float32)
    for x in f32:
        lo, hi = float(x), float(np.nextafter(x, np.float32(np.inf)))
        mid = (lo + hi) / 2                                   # exactly representable in double
        vals += ["%.6f" % lo, "%.9g" % lo, "%.17g" % lo, "%.30f" % mid, "%.25e" % mid,
                 "%.17g" % np.nextafter(mid, np.inf), "%.17g" % np.nextafter(mid, -np.inf)]
    vals += ["0", "-0", "0.0", "-0.000", "1e-45", "1.4e-45", "7e-46", "1e-40", "-3.4028235e38", "3.4028234e+38",
             "1e22", "1e23", "123456789012345678", "1234567890123456789012", "0.1", ".5", "5.", "+2.5", "1E3", "1e+3",
             "100000000000000000000e-20", "0.000000000000000000001e21", "9007199254740993", "16777217", "33554433"]
    while len(vals) % 9:
        vals.append("1")
    lines = [" ".join(vals[i:i + 3]) for i in range(0, len(vals), 3)]
    path = tmp_path / "hard.dat"
    path.write_text("\n".join(lines) + "\n")
    hm, rm = host.mesh(), ref.api.mesh()
    assert hm.read_file(str(path)) and rm.read_file(str(path))
    hv, _ = hm.arrays()
    rv, _ = rm.arrays()
    assert len(hv) == len(vals) // 3
    assert np.array_equal(hv[:, :3].view(np.uint32), rv[:, :3].view(np.uint32))


def test_reader_errors(host, tmp_path):
    m = host.mesh()
    assert not m.read_file(str(tmp_path / "missing.dat"))
    for i, text in enumerate(["", "1 2\n", "3\n\n0 0 0\n1 0 0\n", "0 0 0\n1 0 0\n0 1 0\n0 0 1\n",
                              "3\n\n0 0 0\n1 0 0\n0 1 0\n\n3\n\n0 1 7\n", "3\n\n0 0 0\n1 0 0\n0 1 0\n\n4\n\n0 1 2 0\n"]):
        p = tmp_path / ("bad%d.dat" % i)
        p.write_text(text)
        assert not m.read_file(str(p)), text
    good = tmp_path / "good.dat"
    good.write_text("0 0 0\r\n1 0 0\r\n0 1 0\r\n")  # CRLF, position only -> face normals
    assert m.read_file(str(good))
    v, t = m.arrays()
    assert len(t) == 1 and np.allclose(v[:, 3:], [[0, 0, 1]] * 3)
