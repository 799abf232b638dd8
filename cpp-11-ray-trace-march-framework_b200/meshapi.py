"""ctypes view of a "mesh / matrix" C API with the reference's Mesh + Matrix44f semantics.

Two libraries export the same functions under different prefixes:

* ``rtm_`` -- this package's host library (host/capi.cpp), the drop-in mirror of the reference's
  ``Mesh`` (mesh.h:10-37) and ``Matrix44f`` (lin_alg.h:235-690) classes;
* ``ref_`` -- oracle/ref_driver.cpp around the unmodified reference (test infrastructure).

Scene recipes (scenes.py) are written once against this interface, so the product path and the
oracle build their scenes through their *own* arithmetic and can then be compared bit for bit.
Matrices are 16 float32 in ``Matrix44f::m_mat`` memory order.
"""
import ctypes as C
import os
import struct

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "meshes")

_F32P = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(_F32P)


def load_meshbin(name, flip_winding=False):
    """Read assets/meshes/<name>[.flip].meshbin -> (vtx float32 [V,6], tri uint32 [T,6]).

    The file is the state of the reference's ``Mesh`` right after ``Mesh::Read``
    (oracle/convert_meshes.py documents the layout)."""
    path = os.path.join(ASSET_DIR, name + (".flip" if flip_winding else "") + ".meshbin")
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"RTMMESH1":
        raise ValueError("bad meshbin magic in " + path)
    nv, nt = struct.unpack_from("<II", data, 8)
    vtx = np.frombuffer(data, np.float32, nv * 6, 16).reshape(nv, 6).copy()
    tri = np.frombuffer(data, np.uint32, nt * 6, 16 + nv * 24).reshape(nt, 6).copy()
    return vtx, tri


class MeshHandle:
    """Owns one library-side Mesh object."""

    def __init__(self, api):
        self.api = api
        self.h = api._f("mesh_new")()
        if not self.h:
            raise MemoryError("mesh_new failed")

    def release(self):
        """Give up ownership (the library took it, e.g. renderer_new)."""
        h, self.h = self.h, None
        return h

    def __del__(self):
        if getattr(self, "h", None):
            self.api._f("mesh_free")(self.h)
            self.h = None

    # -- Mesh methods (mesh.h:29-36)
    def set(self, vtx, tri):
        vtx = np.ascontiguousarray(vtx, np.float32)
        tri = np.ascontiguousarray(tri, np.uint32)
        self.api._f("mesh_set")(self.h, _fp(vtx), len(vtx), tri.ctypes.data_as(C.c_void_p), len(tri))
        return self

    def read_asset(self, name, flip_winding=False):
        return self.set(*load_meshbin(name, flip_winding))

    def read_file(self, path, flip_winding=False):
        return bool(self.api._f("mesh_read")(self.h, os.fsencode(path), int(flip_winding)))

    def cornell_box(self):
        self.api._f("mesh_cornell_box")(self.h)
        return self

    def normalize_dimensions(self):
        self.api._f("mesh_normalize_dimensions")(self.h)
        return self

    def transform(self, mat16):
        mat16 = np.ascontiguousarray(mat16, np.float32)
        self.api._f("mesh_transform")(self.h, _fp(mat16))
        return self

    def add_mesh(self, other):
        self.api._f("mesh_add_mesh")(self.h, other.h)
        return self

    def add_quad(self, quad12):
        q = np.ascontiguousarray(quad12, np.float32).reshape(12)
        self.api._f("mesh_add_quad")(self.h, _fp(q))
        return self

    def add_instances(self, base, params):
        p = np.ascontiguousarray(params, np.float32).reshape(-1, 6)
        self.api._f("mesh_add_instances")(self.h, base.h, len(p), _fp(p))
        return self

    def compute_aabb(self):
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.api._f("mesh_compute_aabb")(self.h, _fp(mn), _fp(mx))
        return mn, mx

    @property
    def num_vertices(self):
        return self.api._f("mesh_num_vertices")(self.h)

    @property
    def num_triangles(self):
        return self.api._f("mesh_num_triangles")(self.h)

    def arrays(self):
        vtx = np.empty((self.num_vertices, 6), np.float32)
        tri = np.empty((self.num_triangles, 6), np.uint32)
        self.api._f("mesh_get")(self.h, _fp(vtx), tri.ctypes.data_as(C.c_void_p))
        return vtx, tri


class MeshApi:
    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        self._cache = {}
        p = prefix
        sigs = {
            "mat_identity": (None, [_F32P]),
            "mat_translation": (None, [C.c_float, C.c_float, C.c_float, _F32P]),
            "mat_scaling": (None, [C.c_float, _F32P]),
            "mat_rotation_x": (None, [C.c_float, _F32P]),
            "mat_rotation_y": (None, [C.c_float, _F32P]),
            "mat_rotation_z": (None, [C.c_float, _F32P]),
            "mat_multiply": (None, [_F32P, _F32P, _F32P]),
            "mat_invert": (C.c_int, [_F32P, _F32P]),
            "mat_look_at": (None, [_F32P, _F32P, _F32P]),
            "mesh_new": (C.c_void_p, []),
            "mesh_free": (None, [C.c_void_p]),
            "mesh_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
            "mesh_set": (None, [C.c_void_p, _F32P, C.c_uint32, C.c_void_p, C.c_uint32]),
            "mesh_num_vertices": (C.c_uint32, [C.c_void_p]),
            "mesh_num_triangles": (C.c_uint32, [C.c_void_p]),
            "mesh_get": (None, [C.c_void_p, _F32P, C.c_void_p]),
            "mesh_cornell_box": (None, [C.c_void_p]),
            "mesh_normalize_dimensions": (None, [C.c_void_p]),
            "mesh_transform": (None, [C.c_void_p, _F32P]),
            "mesh_add_mesh": (None, [C.c_void_p, C.c_void_p]),
            "mesh_add_quad": (None, [C.c_void_p, _F32P]),
            "mesh_add_instances": (None, [C.c_void_p, C.c_void_p, C.c_uint32, _F32P]),
            "mesh_compute_aabb": (None, [C.c_void_p, _F32P, _F32P]),
            "camera_constants": (None, [C.c_float, C.c_uint32, C.c_uint32, _F32P, _F32P]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(lib, p + name)
            fn.restype = res
            fn.argtypes = args
            self._cache[name] = fn

    def _f(self, name):
        return self._cache[name]

    # -- matrices (lin_alg.h:317-467)
    def _mat(self, name, *args):
        out = np.zeros(16, np.float32)
        self._f(name)(*args, _fp(out))
        return out

    def identity(self):
        return self._mat("mat_identity")

    def translation(self, x, y, z):
        return self._mat("mat_translation", x, y, z)

    def scaling(self, f):
        return self._mat("mat_scaling", f)

    def rotation_x(self, deg):
        return self._mat("mat_rotation_x", deg)

    def rotation_y(self, deg):
        return self._mat("mat_rotation_y", deg)

    def rotation_z(self, deg):
        return self._mat("mat_rotation_z", deg)

    def multiply(self, a, *rest):
        """a * b * c ... left-associative like the C++ expression (lin_alg.h:304-305)."""
        acc = np.ascontiguousarray(a, np.float32)
        for b in rest:
            b = np.ascontiguousarray(b, np.float32)
            out = np.zeros(16, np.float32)
            self._f("mat_multiply")(_fp(acc), _fp(b), _fp(out))
            acc = out
        return acc

    def invert(self, a):
        a = np.ascontiguousarray(a, np.float32)
        out = np.zeros(16, np.float32)
        ok = self._f("mat_invert")(_fp(a), _fp(out))
        return bool(ok), out

    def look_at(self, eye, at):
        eye = np.asarray(eye, np.float32)
        at = np.asarray(at, np.float32)
        out = np.zeros(16, np.float32)
        self._f("mat_look_at")(_fp(eye), _fp(at), _fp(out))
        return out

    def camera_constants(self, fov, width, height):
        a, b = C.c_float(), C.c_float()
        self._f("camera_constants")(fov, width, height, C.byref(a), C.byref(b))
        return np.float32(a.value), np.float32(b.value)

    def mesh(self):
        return MeshHandle(self)
