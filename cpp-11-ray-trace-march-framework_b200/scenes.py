"""Scene presets of the reference viewer, restated as data + calls on a MeshApi backend.

The numbers are those of ``Application::InitializeScene`` (application.cpp:304-517); each preset
is built by calling the *backend's own* Mesh / Matrix44f operations (meshapi.MeshApi), so the
host library and the reference oracle each produce their scene through their own arithmetic.

``CONFIGS`` maps the BASELINE.json configurations C1..C5 to
(scene, width, height, spp, grid_res).
"""
import numpy as np


def _quad_y(half, y):
    # ground plane quads of presets 3, 8, 9: (-h, y, +d), (+h, y, +d), (+h, y, -d), (-h, y, -d)
    hx, hz = half
    return [-hx, y, hz, hx, y, hz, hx, y, -hz, -hx, y, -hz]


def scene_0(b):  # application.cpp:312-317
    m = b.mesh().read_asset("torusknot_column_teapot_plane").normalize_dimensions()
    return m, 51.0, b.look_at((-1.00001, 1.0, 1.0), (0.0, -0.2, 0.0))


def scene_1(b):  # application.cpp:319-341
    m = b.mesh().cornell_box().normalize_dimensions()
    cube = b.mesh().read_asset("cube").normalize_dimensions()
    cube.transform(b.multiply(b.scaling(0.25), b.rotation_x(45.0), b.rotation_y(45.0),
                              b.translation(0.0, 0.3, 0.0)))
    m.add_mesh(cube)
    return m, 51.0, b.look_at((0.0, 0.0, -2.0), (0.0, 0.0, 0.0))


def scene_2(b):  # application.cpp:343-348
    m = b.mesh().read_asset("room_table_chair_tv").normalize_dimensions()
    return m, 90.0, b.look_at((-0.47, 0.15, -0.3), (1.0, -0.7, 0.9))


def scene_3(b):  # application.cpp:350-366
    m = b.mesh().read_asset("table_chair", flip_winding=True).normalize_dimensions()
    m.add_quad(_quad_y((1.2, 1.2), -0.219097))
    return m, 45.0, b.look_at((1.001, 1.002, -1.0), (0.0, 0.0, -0.3))


def scene_4(b):  # application.cpp:368-378
    m = b.mesh().read_asset("head").normalize_dimensions()
    m.transform(b.rotation_y(30.0))
    return m, 75.0, b.look_at((0.0, 0.0, -1.0), (0.0, 0.0, 0.0))


def scene_5(b):  # application.cpp:380-400
    m = b.mesh().read_asset("room_three_windows_two_columns").normalize_dimensions()
    cat = b.mesh().read_asset("cat").normalize_dimensions()
    cat.transform(b.multiply(b.scaling(0.25), b.rotation_y(30.0), b.translation(-0.065, -0.1, 0.05)))
    m.add_mesh(cat)
    return m, 90.0, b.look_at((-0.2, 0.0, -0.33), (0.0, 0.0, 0.0))


def scene_6(b):  # application.cpp:402-420
    m = b.mesh().read_asset("water_surface").normalize_dimensions()
    knot = b.mesh().read_asset("torus_knot").normalize_dimensions()
    knot.transform(b.multiply(b.scaling(0.25), b.translation(-0.0, 0.2, 0.0)))
    m.add_mesh(knot)
    return m, 30.0, b.look_at((-1.0, 2.0, -1.0), (0.0, 0.0, 0.0))


def scene_7(b):  # application.cpp:422-441
    m = b.mesh().read_asset("griebel").normalize_dimensions()
    pot = b.mesh().read_asset("teapot").normalize_dimensions()
    pot.transform(b.multiply(b.scaling(0.3), b.rotation_y(90.0), b.translation(0.0, 0.1, 0.0)))
    m.add_mesh(pot)
    return m, 75.0, b.look_at((0.5, 0.5, 0.0), (0.0, 0.0, 0.0))


def scene_8(b):  # application.cpp:443-459
    m = b.mesh().read_asset("killeroo").normalize_dimensions()
    m.add_quad(_quad_y((0.75, 0.75), -0.229267))
    return m, 30.0, b.look_at((-1.6, 1.2, -1.0), (0.0, 0.0, -0.1))


def scene_9(b):  # application.cpp:461-500
    m = b.mesh()
    dwarf = b.mesh().read_asset("d3d_dwarf").normalize_dimensions()
    dwarf.transform(b.translation(0.0, 0.500100, 0.0))
    m.add_mesh(dwarf)
    hand = b.mesh().read_asset("hand").normalize_dimensions()
    hand.transform(b.multiply(b.rotation_x(90.0), b.rotation_y(90.0), b.translation(0.7, 0.490801, 0.0)))
    m.add_mesh(hand)
    blob = b.mesh().read_asset("blob").normalize_dimensions()
    blob.transform(b.multiply(b.scaling(0.6), b.translation(-0.8, 0.278176, 0.0)))
    m.add_mesh(blob)
    m.add_quad([-1.5, 0.0, 1.0, 1.5, 0.0, 1.0, 1.5, 0.0, -1.0, -1.5, 0.0, -1.0])
    return m, 60.0, b.look_at((0.0, 1.5, -2.0), (0.0, 0.0, 0.0))


SOUP_SEED = 20261018


def soup_params(nx, ny, nz, seed=SOUP_SEED):
    """Instance parameters of the synthetic tiger soup (config C5; builder-defined because the
    reference has no such scene -- SURVEY.md D3).  n = nx*ny*nz instances on a jittered lattice
    in [-0.5, 0.5]^3, x fastest; per instance, drawn in this order from numpy's legacy
    ``RandomState(seed)`` (a frozen stream): rot_y deg, rot_x deg, jitter x, y, z.
    Row = {scale, rot_y, rot_x, tx, ty, tz} for Scaling * RotationY * RotationX * Translation."""
    rs = np.random.RandomState(seed)
    n = nx * ny * nz
    draws = rs.random_sample((n, 5))
    idx = np.arange(n)
    ix, iy, iz = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    pitch = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
    centre = np.stack([ix + 0.5, iy + 0.5, iz + 0.5], 1) * pitch - 0.5
    jitter = (draws[:, 2:5] - 0.5) * 0.5 * pitch
    p = np.empty((n, 6), np.float32)
    p[:, 0] = 0.8 / max(nx, ny, nz)
    p[:, 1] = draws[:, 0] * 360.0
    p[:, 2] = draws[:, 1] * 360.0
    p[:, 3:6] = centre + jitter
    return p


def tiger_soup(b, nx=44, ny=44, nz=43):
    """44 x 44 x 43 = 83 248 tigers = 50 115 296 triangles."""
    base = b.mesh().read_asset("tiger").normalize_dimensions()
    m = b.mesh()
    m.add_instances(base, soup_params(nx, ny, nz))
    m.normalize_dimensions()
    return m, 45.0, b.look_at((-1.6, 1.2, -1.0), (0.0, 0.0, 0.0))


PRESETS = {0: scene_0, 1: scene_1, 2: scene_2, 3: scene_3, 4: scene_4,
           5: scene_5, 6: scene_6, 7: scene_7, 8: scene_8, 9: scene_9}

NAMED = {
    "torusknot": scene_0, "cornell": scene_1, "room": scene_2, "table_chair": scene_3,
    "head": scene_4, "room_cat": scene_5, "water_knot": scene_6, "griebel_teapot": scene_7,
    "killeroo": scene_8, "dwarf_hand_blob": scene_9,
    "tiger_soup": tiger_soup,
    "tiger_soup_small": lambda b: tiger_soup(b, 6, 6, 5),
    "tiger_soup_medium": lambda b: tiger_soup(b, 16, 16, 15),
}

# BASELINE.json configs -> (scene, width, height, spp, grid_res)
CONFIGS = {
    "C1": ("cornell", 512, 512, 1, 64),
    "C2": ("killeroo", 1920, 1080, 4, 64),
    "C3": ("torusknot", 1920, 1080, 16, 64),
    "C4": ("room", 3840, 2160, 16, 64),
    "C5": ("tiger_soup", 3840, 2160, 16, 768),  # optimum of the density sweep (profiles/r02_c5_grid_density_sweep.txt)
    "killeroo4k": ("killeroo", 3840, 2160, 16, 64),
}


def build(b, name):
    """-> (MeshHandle, fov_degrees, cam_mat16) built through backend ``b``."""
    fn = PRESETS[name] if isinstance(name, int) else NAMED[name]
    return fn(b)
