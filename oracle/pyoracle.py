"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU oracles.

* ``Ref``  -- oracle/_ref/libref_oracle.so, the UNMODIFIED reference + oracle/ref_driver.cpp.
* ``Port`` -- oracle/librt_oracle.so, the plain-C restatement (oracle/rt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "librt_oracle.so")
MISS = 0xFFFFFFFF

def _meshapi():
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("cpp-11-ray-trace-march-framework_b200.meshapi")


_F32P = C.POINTER(C.c_float)
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(ref=True, port=True):
    """Compile the oracles (the reference half only where /root/reference exists)."""
    targets = []
    if port:
        targets.append("port")
    if ref and os.path.isdir("/root/reference"):
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-s", "-C", HERE] + targets)


def have_ref():
    return os.path.exists(REF_SO)


# ------------------------------------------------------------------------------------------ ref
class Ref:
    """The reference itself.  ``api`` is a MeshApi with the reference's own Mesh/Matrix44f."""

    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        self.lib = lib = C.CDLL(REF_SO)
        self.api = _meshapi().MeshApi(lib, "ref_")
        lib.ref_hardware_threads.restype = C.c_int
        lib.ref_renderer_new.restype = C.c_void_p
        lib.ref_renderer_new.argtypes = [C.c_void_p, C.c_float, _F32P, C.c_uint32]
        lib.ref_renderer_free.argtypes = [C.c_void_p]
        lib.ref_renderer_set_threads.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_renderer_get_threads.argtypes = [C.c_void_p]
        lib.ref_renderer_get_threads.restype = C.c_uint32
        lib.ref_renderer_render.restype = C.c_double
        lib.ref_renderer_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _U32P]
        lib.ref_renderer_render_tile_subset.restype = C.c_double
        lib.ref_renderer_render_tile_subset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                        C.c_uint32, C.c_uint32, C.c_uint32, _U32P, _U64P]
        lib.ref_trace_hits.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_uint32, _U32P, _F32P, _F32P, _F32P]
        lib.ref_intersect_rays.argtypes = [C.c_void_p, C.c_uint32, _F32P, _F32P, _U32P, _F32P, _F32P, _F32P]
        lib.ref_generate_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_uint32, _F32P, _F32P]
        lib.ref_sample_table.argtypes = [C.c_uint32, _F32P]
        lib.ref_grid_info.restype = C.c_uint64
        lib.ref_grid_info.argtypes = [C.c_void_p, _U32P, _F32P, _F32P, _F32P, _F32P]
        lib.ref_grid_dump.argtypes = [C.c_void_p, _U64P, _U32P]
        lib.ref_scene_num_vertices.argtypes = [C.c_void_p]
        lib.ref_scene_num_vertices.restype = C.c_uint32
        lib.ref_scene_num_triangles.argtypes = [C.c_void_p]
        lib.ref_scene_num_triangles.restype = C.c_uint32
        lib.ref_scene_get_mesh.argtypes = [C.c_void_p, _F32P, C.c_void_p]
        lib.ref_save_bmp.argtypes = [C.c_void_p, C.c_char_p]
        lib.ref_tri_test.restype = C.c_int
        lib.ref_tri_test.argtypes = [C.c_int, _F32P, _F32P, _F32P, _F32P, _F32P, _F32P, _F32P]

    def hardware_threads(self):
        return self.lib.ref_hardware_threads()

    # -- the reference's sampling module (sampling.h / sampling.cpp)
    def qmc_sequence(self, kind, scramble, n_begin, count, dim_begin=0, dim_count=1, num_smp=1, bits=0):
        out = np.zeros((count, dim_count), np.float64)
        self.lib.ref_qmc_sequence.argtypes = [C.c_uint32] * 8 + [C.POINTER(C.c_double)]
        self.lib.ref_qmc_sequence(kind, scramble, n_begin, count, dim_begin, dim_count, num_smp, bits,
                                  out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def qmc_tables(self, scramble, n_primes):
        """Permutation tables of the first n_primes primes back to back (scramble 1 BW, 2 Faure, 3 reverse, 4 randomized)."""
        self.lib.ref_qmc_table.restype = C.c_uint32
        self.lib.ref_qmc_table.argtypes = [C.c_uint32, C.c_uint32, _U32P]
        parts = []
        for i in range(n_primes):
            buf = np.zeros(8192, np.uint32)
            n = self.lib.ref_qmc_table(scramble, i, _p(buf, _U32P))
            parts.append(buf[:n].copy())
        return np.concatenate(parts)

    def qmc_prime(self, idx):
        self.lib.ref_qmc_prime.restype = C.c_uint32
        return int(self.lib.ref_qmc_prime(C.c_uint32(idx)))

    def cranley_patterson(self, x, e):
        self.lib.ref_qmc_cranley_patterson.restype = C.c_double
        self.lib.ref_qmc_cranley_patterson.argtypes = [C.c_double, C.c_double]
        return np.array([self.lib.ref_qmc_cranley_patterson(float(v), float(e)) for v in np.ravel(x)], np.float64)

    def sample_table(self, spp):
        xy = np.zeros((spp, 2), np.float32)
        self.lib.ref_sample_table(spp, _p(xy, _F32P))
        return xy

    def tri_test(self, variant, o, d, v0, v1, v2, n):
        a = [np.ascontiguousarray(x, np.float32) for x in (o, d, v0, v1, v2, n)]
        out = np.zeros(3, np.float32)
        hit = self.lib.ref_tri_test(variant, *[_p(x, _F32P) for x in a], _p(out, _F32P))
        return bool(hit), out

    def renderer(self, mesh, fov, cam16, grid_res=64):
        return RefRenderer(self, mesh, fov, cam16, grid_res)


class RefRenderer:
    """Reference ``Renderer`` + ``Scene`` (takes ownership of the MeshHandle)."""

    def __init__(self, ref, mesh, fov, cam16, grid_res=64):
        self.ref = ref
        self.lib = ref.lib
        self.fov = float(fov)
        self.cam16 = np.ascontiguousarray(cam16, np.float32)
        self.h = self.lib.ref_renderer_new(mesh.release(), self.fov, _p(self.cam16, _F32P), grid_res)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_renderer_free(self.h)
            self.h = None

    def set_threads(self, n):
        self.lib.ref_renderer_set_threads(self.h, n)

    def get_threads(self):
        return self.lib.ref_renderer_get_threads(self.h)

    def render(self, width, height, spp, want_image=True):
        """-> (seconds, bgra uint32 [H,W], row 0 = y 0) through the reference's own thread pool."""
        img = np.zeros((height, width), np.uint32) if want_image else None
        sec = self.lib.ref_renderer_render(self.h, width, height, spp, _p(img, _U32P))
        return sec, img

    def render_tile_subset(self, width, height, spp, stride, offset, n_threads):
        tiles, pix = C.c_uint32(), C.c_uint64()
        sec = self.lib.ref_renderer_render_tile_subset(self.h, width, height, spp, stride, offset,
                                                       n_threads, C.byref(tiles), C.byref(pix))
        return sec, tiles.value, pix.value

    def trace_hits(self, width, height, spp, y_begin=0, y_end=None, n_threads=0, want_tuv=True):
        y_end = height if y_end is None else y_end
        n_threads = n_threads or self.ref.hardware_threads()
        shape = (y_end - y_begin, width, spp)
        idx = np.empty(shape, np.uint32)
        t = np.empty(shape, np.float32) if want_tuv else None
        u = np.empty(shape, np.float32) if want_tuv else None
        v = np.empty(shape, np.float32) if want_tuv else None
        self.lib.ref_trace_hits(self.h, width, height, spp, y_begin, y_end, n_threads,
                                _p(idx, _U32P), _p(t, _F32P), _p(u, _F32P), _p(v, _F32P))
        return idx, t, u, v

    def intersect_rays(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        idx, t, u, v = (np.empty(n, np.uint32), np.empty(n, np.float32), np.empty(n, np.float32),
                        np.empty(n, np.float32))
        self.lib.ref_intersect_rays(self.h, n, _p(o, _F32P), _p(d, _F32P), _p(idx, _U32P),
                                    _p(t, _F32P), _p(u, _F32P), _p(v, _F32P))
        return idx, t, u, v

    def intersect_rays_brute_force(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        idx, t, u, v = (np.empty(n, np.uint32), np.empty(n, np.float32), np.empty(n, np.float32),
                        np.empty(n, np.float32))
        self.lib.ref_intersect_rays_brute_force.argtypes = [C.c_void_p, C.c_uint32, _F32P, _F32P, _U32P, _F32P, _F32P, _F32P]
        self.lib.ref_intersect_rays_brute_force(self.h, n, _p(o, _F32P), _p(d, _F32P), _p(idx, _U32P),
                                                _p(t, _F32P), _p(u, _F32P), _p(v, _F32P))
        return idx, t, u, v

    def ray_march(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        hit, t = np.empty(n, np.uint32), np.empty(n, np.float32)
        self.lib.ref_ray_march.argtypes = [C.c_void_p, C.c_uint32, _F32P, _F32P, _U32P, _F32P]
        self.lib.ref_ray_march(self.h, n, _p(o, _F32P), _p(d, _F32P), _p(hit, _U32P), _p(t, _F32P))
        return hit, t

    def generate_rays(self, width, height, spp, y_begin, y_end):
        shape = (y_end - y_begin, width, spp, 3)
        o, d = np.empty(shape, np.float32), np.empty(shape, np.float32)
        self.lib.ref_generate_rays(self.h, width, height, spp, y_begin, y_end, _p(o, _F32P), _p(d, _F32P))
        return o, d

    def generate_rays_ortho(self, width, height, spp, y_begin, y_end, ortho_width):
        shape = (y_end - y_begin, width, spp, 3)
        o, d = np.empty(shape, np.float32), np.empty(shape, np.float32)
        self.lib.ref_generate_rays_ortho.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                     C.c_float, _F32P, _F32P]
        self.lib.ref_generate_rays_ortho(self.h, width, height, spp, y_begin, y_end, ortho_width, _p(o, _F32P), _p(d, _F32P))
        return o, d

    def grid(self):
        """-> dict(dim, aabb_min, aabb_max, cell_wdh, inv_cell_wdh, cell_offset u64, tri_index u32)"""
        dim = np.zeros(3, np.uint32)
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        cw, icw = C.c_float(), C.c_float()
        refs = self.lib.ref_grid_info(self.h, _p(dim, _U32P), _p(mn, _F32P), _p(mx, _F32P),
                                      C.byref(cw), C.byref(icw))
        cells = int(dim[0]) * int(dim[1]) * int(dim[2])
        off = np.zeros(cells + 1, np.uint64)
        tri = np.zeros(max(refs, 1), np.uint32)
        self.lib.ref_grid_dump(self.h, _p(off, _U64P), _p(tri, _U32P))
        return dict(dim=dim, aabb_min=mn, aabb_max=mx, cell_wdh=np.float32(cw.value),
                    inv_cell_wdh=np.float32(icw.value), cell_offset=off, tri_index=tri[:refs])

    def mesh_arrays(self):
        nv = self.lib.ref_scene_num_vertices(self.h)
        nt = self.lib.ref_scene_num_triangles(self.h)
        vtx = np.empty((nv, 6), np.float32)
        tri = np.empty((nt, 6), np.uint32)
        self.lib.ref_scene_get_mesh(self.h, _p(vtx, _F32P), tri.ctypes.data_as(C.c_void_p))
        return vtx, tri

    def save_bmp(self, path):
        self.lib.ref_save_bmp(self.h, os.fsencode(path))


# ----------------------------------------------------------------------------------------- port
class _RenderOptions(C.Structure):
    _fields_ = [("ortho", C.c_int), ("ortho_width", C.c_float), ("shade_mode", C.c_int)]


class _Grid(C.Structure):
    _fields_ = [("dim", C.c_uint32 * 3), ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
                ("cell_wdh", C.c_float), ("inv_cell_wdh", C.c_float), ("num_cells", C.c_uint64),
                ("num_refs", C.c_uint64), ("cell_offset", _U64P), ("tri_index", _U32P)]


class _Scene(C.Structure):
    _fields_ = [("vtx", _F32P), ("tri", _U32P), ("num_vtx", C.c_uint32), ("num_tri", C.c_uint32),
                ("grid", _Grid)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "cells", "tri_tests", "hits", "rej_det", "rej_u",
                                          "rej_v", "full", "box_miss", "nonempty")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Port:
    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = lib = C.CDLL(PORT_SO)
        lib.rto_sample_table.argtypes = [C.c_uint32, _F32P]
        lib.rto_radical_inverse.restype = C.c_double
        lib.rto_radical_inverse.argtypes = [C.c_uint32, C.c_uint32]
        lib.rto_camera_constants.argtypes = [C.c_float, C.c_uint32, C.c_uint32, _F32P, _F32P]
        lib.rto_generate_ray.argtypes = [_F32P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                         C.c_float, C.c_float, C.c_float, _F32P, _F32P]
        lib.rto_ray_tri.restype = C.c_int
        lib.rto_ray_tri.argtypes = [_F32P] * 5 + [_F32P] * 3 + [C.c_void_p]
        lib.rto_ray_tri_bary.restype = C.c_int
        lib.rto_ray_tri_bary.argtypes = [_F32P] * 6 + [_F32P] * 3 + [C.c_void_p]
        lib.rto_grid_build.restype = C.c_int
        lib.rto_grid_build.argtypes = [C.POINTER(_Scene), C.c_uint32, C.c_uint32]
        lib.rto_grid_free.argtypes = [C.POINTER(_Grid)]
        lib.rto_render_rows.argtypes = [C.POINTER(_Scene), _F32P, C.c_float, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                        _U32P, _U32P, _F32P, _F32P, _F32P, C.POINTER(Counters)]
        lib.rto_intersect_rays.argtypes = [C.POINTER(_Scene), C.c_uint32, _F32P, _F32P, C.c_int, _U32P,
                                           _F32P, _F32P, _F32P]
        lib.rto_powf_vs_sqrtf.restype = C.c_uint64
        lib.rto_powf_vs_sqrtf.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _U64P]
        lib.rto_resolve_pixel.restype = C.c_uint32
        lib.rto_resolve_pixel.argtypes = [_F32P, C.c_uint32, C.c_int]

    def sample_table(self, spp):
        xy = np.zeros((spp, 2), np.float32)
        self.lib.rto_sample_table(spp, _p(xy, _F32P))
        return xy

    def camera_constants(self, fov, width, height):
        a, b = C.c_float(), C.c_float()
        self.lib.rto_camera_constants(fov, width, height, C.byref(a), C.byref(b))
        return np.float32(a.value), np.float32(b.value)

    def tri_test(self, variant, o, d, v0, v1, v2, n):
        a = [np.ascontiguousarray(x, np.float32) for x in (o, d, v0, v1, v2, n)]
        t, u, v = C.c_float(), C.c_float(), C.c_float()
        if variant == 0:
            hit = self.lib.rto_ray_tri(*[_p(x, _F32P) for x in a[:5]], C.byref(t), C.byref(u), C.byref(v), None)
        else:
            hit = self.lib.rto_ray_tri_bary(*[_p(x, _F32P) for x in a], C.byref(t), C.byref(u), C.byref(v), None)
        return bool(hit), np.array([t.value, u.value, v.value], np.float32)

    def powf_vs_sqrtf(self, lo_bits, hi_bits, n_threads=0):
        bd = C.c_uint64()
        n = self.lib.rto_powf_vs_sqrtf(lo_bits, hi_bits, n_threads or (os.cpu_count() or 1), C.byref(bd))
        return int(n), int(bd.value)

    def scene(self, vtx, tri, grid_res=64, grid=None, n_threads=0, tight_ranges=False):
        C.c_int.in_dll(self.lib, "rto_grid_build_tight_ranges").value = int(tight_ranges)
        try:
            return PortScene(self, vtx, tri, grid_res, grid, n_threads)
        finally:
            C.c_int.in_dll(self.lib, "rto_grid_build_tight_ranges").value = 0


class PortScene:
    """Post-transform mesh arrays + a grid (built by the port, or injected from elsewhere)."""

    def __init__(self, port, vtx, tri, grid_res=64, grid=None, n_threads=0):
        self.port = port
        self.lib = port.lib
        self.vtx = np.ascontiguousarray(vtx, np.float32)
        self.tri = np.ascontiguousarray(tri, np.uint32)
        self.s = _Scene()
        self.s.vtx = _p(self.vtx, _F32P)
        self.s.tri = _p(self.tri, _U32P)
        self.s.num_vtx = len(self.vtx)
        self.s.num_tri = len(self.tri)
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self._own = grid is None
        if grid is None:
            if self.lib.rto_grid_build(C.byref(self.s), grid_res, self.n_threads) != 0:
                raise ValueError("rto_grid_build failed")
        else:
            g = self.s.grid
            self._keep = (np.ascontiguousarray(grid["cell_offset"], np.uint64),
                          np.ascontiguousarray(grid["tri_index"], np.uint32))
            for k in range(3):
                g.dim[k] = int(grid["dim"][k])
                g.aabb_min[k] = float(grid["aabb_min"][k])
                g.aabb_max[k] = float(grid["aabb_max"][k])
            g.cell_wdh = float(grid["cell_wdh"])
            g.inv_cell_wdh = float(grid["inv_cell_wdh"])
            g.num_cells = len(self._keep[0]) - 1
            g.num_refs = len(self._keep[1])
            g.cell_offset = _p(self._keep[0], _U64P)
            g.tri_index = _p(self._keep[1], _U32P)

    def __del__(self):
        if getattr(self, "_own", False):
            self.lib.rto_grid_free(C.byref(self.s.grid))
            self._own = False

    def grid(self):
        g = self.s.grid
        cells, refs = int(g.num_cells), int(g.num_refs)
        return dict(dim=np.array(list(g.dim), np.uint32), aabb_min=np.array(list(g.aabb_min), np.float32),
                    aabb_max=np.array(list(g.aabb_max), np.float32), cell_wdh=np.float32(g.cell_wdh),
                    inv_cell_wdh=np.float32(g.inv_cell_wdh),
                    cell_offset=np.ctypeslib.as_array(g.cell_offset, (cells + 1,)).copy(),
                    tri_index=(np.ctypeslib.as_array(g.tri_index, (refs,)).copy() if refs
                               else np.zeros(0, np.uint32)))

    def render(self, cam16, fov, width, height, spp, variant=0, gamma=True, y_begin=0, y_end=None,
               want_hits=False, want_tuv=False, n_threads=0, ortho_width=None, shade_mode=0):
        """-> dict(bgra [rows,W], tri [rows,W,spp]?, t/u/v?, counters).  ortho_width: orthographic camera
        (camera.h:25-36); shade_mode 1 / 2: face normal / depth (renderer.cpp:116,118)"""
        y_end = height if y_end is None else y_end
        cam16 = np.ascontiguousarray(cam16, np.float32)
        rows = y_end - y_begin
        bgra = np.zeros((rows, width), np.uint32)
        tri = np.empty((rows, width, spp), np.uint32) if want_hits else None
        t = np.empty((rows, width, spp), np.float32) if want_tuv else None
        u = np.empty((rows, width, spp), np.float32) if want_tuv else None
        v = np.empty((rows, width, spp), np.float32) if want_tuv else None
        cnt = Counters()
        opt = _RenderOptions(int(ortho_width is not None), float(ortho_width or 0.0), int(shade_mode))
        self.lib.rto_render_rows_ex.restype = None
        self.lib.rto_render_rows_ex.argtypes = [C.POINTER(_Scene), _F32P, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.POINTER(_RenderOptions), _U32P, _U32P, _F32P, _F32P, _F32P,
                                                C.POINTER(Counters)]
        self.lib.rto_render_rows_ex(C.byref(self.s), _p(cam16, _F32P), fov, width, height, spp, variant,
                                    int(gamma), y_begin, y_end, n_threads or self.n_threads, C.byref(opt),
                                    _p(bgra, _U32P), _p(tri, _U32P), _p(t, _F32P), _p(u, _F32P),
                                    _p(v, _F32P), C.byref(cnt))
        return dict(bgra=bgra, tri=tri, t=t, u=u, v=v, counters=cnt.as_dict())

    def intersect_rays(self, origins, dirs, variant=0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        idx, t, u, v = (np.empty(n, np.uint32), np.empty(n, np.float32), np.empty(n, np.float32),
                        np.empty(n, np.float32))
        self.lib.rto_intersect_rays(C.byref(self.s), n, _p(o, _F32P), _p(d, _F32P), variant,
                                    _p(idx, _U32P), _p(t, _F32P), _p(u, _F32P), _p(v, _F32P))
        return idx, t, u, v
