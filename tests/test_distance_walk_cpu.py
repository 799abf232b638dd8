"""CPU checks of the host-side arithmetic behind the large-grid traversal (no GPU needed):

* the property the distance walk of K1 (mode 3) / K7 rests on -- from a cell of city-block distance v to the nearest
  occupied or padding cell, ANY chain of face-adjacent steps meets only empty cells during its first v - 1 steps -- on
  the ORACLE's grid of a real scene and on the actual cells the reference's 3D-DDA visits;
* the grid density heuristic cuda_trace_suggest_grid_res (pure host arithmetic in the C ABI library)."""
import numpy as np

from conftest import pkg


def cityblock(occ):
    d = np.where(occ, 0, 10 ** 6).astype(np.int64)
    for axis in range(3):
        d = np.moveaxis(d, axis, 0)
        for i in range(1, d.shape[0]):
            np.minimum(d[i], d[i - 1] + 1, out=d[i])
        for i in range(d.shape[0] - 2, -1, -1):
            np.minimum(d[i], d[i + 1] + 1, out=d[i])
        d = np.moveaxis(d, 0, axis)
    return np.minimum(d, 255)


def padded_occupancy(grid):
    dx, dy, dz = [int(v) for v in grid["dim"]]
    off = np.asarray(grid["cell_offset"]).astype(np.int64)
    occ = np.ones((dy + 2, dz + 2, dx + 2), bool)  # [y][z][x]: cell = x + z*dx + y*dx*dz (grid.h:41-42); padding counts as occupied
    occ[1:-1, 1:-1, 1:-1] = (off[1:] != off[:-1]).reshape(dy, dz, dx)
    return occ


def test_blind_steps_never_skip_an_occupied_cell(port, scene_data):
    sd = scene_data("killeroo")
    g = port.scene(sd.vtx, sd.tri, 40).grid()
    occ = padded_occupancy(g)
    dist = cityblock(occ)
    assert dist[occ].max() == 0 and dist[~occ].min() >= 1 and dist.max() > 4
    rs = np.random.RandomState(3)
    empties = np.argwhere(~occ)
    steps = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])
    checked = 0
    for start in empties[rs.choice(len(empties), 4000, replace=False)]:
        v = int(dist[tuple(start)])
        # a DDA walk is monotone per axis: pick a sign per axis, then a random order of axes
        signs = rs.choice([-1, 1], 3)
        pos = start.copy()
        for j in range(1, v):
            a = rs.randint(3)
            pos[a] += signs[a]
            assert not occ[tuple(pos)], "a blind step landed on an occupied cell"
            checked += 1
    assert checked > 5000
    # neighbouring cells differ by at most one: what bounds the look-ups a ray saves
    for axis in range(3):
        assert np.abs(np.diff(dist, axis=axis)).max() <= 1
    del steps


def test_blind_steps_on_the_cells_the_reference_walk_visits(port, scene_data):
    """The same on real rays: replay the reference's 3D-DDA (oracle port, work counters) is not needed -- the cells a
    ray visits form a face-adjacent chain, so it suffices that the chain property holds from every empty cell; here
    the walk is restated in numpy for a few hundred camera rays and every blind step is checked against the grid."""
    sd = scene_data("cornell")
    ps = port.scene(sd.vtx, sd.tri, 32)
    g = ps.grid()
    occ = padded_occupancy(g)
    dist = cityblock(occ)
    mn, cw = np.asarray(g["aabb_min"], np.float64), float(g["cell_wdh"])
    dim = np.asarray(g["dim"], np.int64)
    rs = np.random.RandomState(5)
    blind = looked = 0
    for _ in range(300):
        o = rs.uniform(mn, mn + dim * cw)  # inside the grid
        d = rs.normal(size=3)
        d /= np.linalg.norm(d)
        pos = np.clip(((o - mn) / cw).astype(np.int64), 0, dim - 1)
        step = np.where(d > 0, 1, -1)
        nxt = np.where(d > 0, (mn + (pos + 1) * cw - o) / d, (mn + pos * cw - o) / d)
        delta = np.abs(cw / d)
        k = 0
        while True:
            pidx = (pos[1] + 1, pos[2] + 1, pos[0] + 1)
            if k == 0:
                looked += 1
                k = int(dist[pidx])
                if k == 0:
                    break  # occupied (the walk would test its triangles) or outside
            else:
                assert not occ[pidx]
                blind += 1
            a = int(np.argmin(nxt))
            pos[a] += step[a]
            nxt[a] += delta[a]
            k -= 1
            if (pos < 0).any() or (pos >= dim).any():
                assert k == 0 or dist[(pos[1] + 1, pos[2] + 1, pos[0] + 1)] == 0
                break
    assert blind > looked / 4 and looked > 300


def test_grid_density_heuristic():
    lib = pkg("capi").load_library()
    f = lib.cuda_trace_suggest_grid_res
    assert f(1) == 16 and f(44) == 16                      # floor
    assert f(24336) == 42 and f(370) == 16                 # ~3 cells per triangle while the occupancy map fits in shared memory
    assert f(50115296) == 767                              # ~9 per triangle beyond: the soup sweep's optimum (768)
    assert f(4000000000) == 896                            # cap
    vals = [f(n) for n in (10 ** k for k in range(1, 10))]
    assert vals == sorted(vals)                            # monotone
