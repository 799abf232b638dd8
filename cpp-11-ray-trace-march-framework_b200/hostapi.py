"""ctypes binding of librtm_host.so: the host-side mirror of the reference's Mesh / Matrix44f /
Scene / Renderer classes (host/*.cpp), via the flat C wrappers of host/capi.cpp.

``host_api()`` returns a meshapi.MeshApi (mesh + matrix operations, no GPU needed);
``HostRenderer`` is the reference's ``Renderer`` usage pattern: construct from a mesh + camera,
``render(width, height, spp)`` = SetSampleCount -> Resize/StartRendering -> wait -> tiles.
"""
import ctypes as C
import os

import numpy as np

from . import capi, meshapi

LIB_PATH = os.path.join(capi.PKG, "librtm_host.so")
_F32P = C.POINTER(C.c_float)
_U32P = C.POINTER(C.c_uint32)

_lib = None
_api = None


def load_host_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(LIB_PATH + " is missing: run `python cpp-11-ray-trace-march-framework_b200/build.py`")
        capi.load_library()  # librtm_host.so links against libcuda_trace.so
        lib = C.CDLL(LIB_PATH)
        lib.rtm_last_error.restype = C.c_char_p
        lib.rtm_mesh_read_binary.restype = C.c_int
        lib.rtm_mesh_read_binary.argtypes = [C.c_void_p, C.c_char_p]
        lib.rtm_renderer_new.restype = C.c_void_p
        lib.rtm_renderer_new.argtypes = [C.c_void_p, C.c_float, _F32P, C.c_uint32, C.c_int]
        lib.rtm_renderer_free.argtypes = [C.c_void_p]
        lib.rtm_renderer_render.restype = C.c_double
        lib.rtm_renderer_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _U32P]
        lib.rtm_renderer_last_kernel_ms.restype = C.c_float
        lib.rtm_renderer_last_kernel_ms.argtypes = [C.c_void_p]
        lib.rtm_renderer_save_bmp.argtypes = [C.c_void_p, C.c_char_p]
        lib.rtm_renderer_start.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.rtm_renderer_wait.argtypes = [C.c_void_p]
        lib.rtm_renderer_stop.argtypes = [C.c_void_p]
        lib.rtm_renderer_finished_tiles.argtypes = [C.c_void_p]
        lib.rtm_renderer_finished_tiles.restype = C.c_uint32
        lib.rtm_renderer_last_render_seconds.argtypes = [C.c_void_p]
        lib.rtm_renderer_last_render_seconds.restype = C.c_double
        lib.rtm_renderer_copy_bitmap.argtypes = [C.c_void_p, _U32P]
        lib.rtm_renderer_device_context.restype = C.c_void_p
        lib.rtm_renderer_device_context.argtypes = [C.c_void_p]
        lib.rtm_renderer_intersect.restype = C.c_int
        lib.rtm_renderer_intersect.argtypes = [C.c_void_p, _F32P, _F32P, _F32P, _U32P]
        lib.rtm_renderer_set_alternates.argtypes = [C.c_void_p, C.c_float, C.c_uint32]
        lib.rtm_renderer_ray_march.restype = C.c_int
        lib.rtm_renderer_ray_march.argtypes = [C.c_void_p, _F32P, _F32P, _F32P]
        lib.rtm_renderer_grid_info.argtypes = [C.c_void_p, _U32P, _F32P, _F32P, _F32P, C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def host_api():
    """MeshApi over the host library (Mesh + Matrix44f mirror)."""
    global _api
    if _api is None:
        _api = meshapi.MeshApi(load_host_library(), "rtm_")
    return _api


class HostRenderer:
    """Mesh -> Scene(grid_res) -> Renderer on ``n_gpus`` GPUs.  Takes ownership of the mesh."""

    def __init__(self, mesh, fov, cam16, grid_res=64, n_gpus=1):
        self.lib = load_host_library()
        self.cam16 = np.ascontiguousarray(cam16, np.float32)
        self.h = self.lib.rtm_renderer_new(mesh.release(), float(fov), self.cam16.ctypes.data_as(_F32P), grid_res, n_gpus)
        if not self.h:
            raise RuntimeError("Renderer construction failed: " + self.lib.rtm_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rtm_renderer_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def render(self, width, height, spp, variant=0, gamma=True, out=None, want_image=True):
        """-> (seconds, image [H, W] uint32 or None)"""
        if want_image and out is None:
            out = np.zeros((height, width), np.uint32)
        sec = self.lib.rtm_renderer_render(self.h, width, height, spp, variant, int(gamma),
                                           out.ctypes.data_as(_U32P) if want_image else None)
        if sec < 0:
            raise RuntimeError("render failed: " + self.lib.rtm_last_error().decode())
        return sec, out

    # -- the asynchronous form (Resize / StartRendering return at once, framebuffer.cpp:94-134)
    def start(self, width, height, spp):
        if self.lib.rtm_renderer_start(self.h, width, height, spp):
            raise RuntimeError("start failed: " + self.lib.rtm_last_error().decode())

    def wait(self):
        """WaitRendering(); raises if the frame failed.  -> seconds from the start call to the last tile"""
        if self.lib.rtm_renderer_wait(self.h):
            raise RuntimeError("render failed: " + self.lib.rtm_last_error().decode())
        return float(self.lib.rtm_renderer_last_render_seconds(self.h))

    def stop(self):
        self.lib.rtm_renderer_stop(self.h)

    def finished_tiles(self):
        return int(self.lib.rtm_renderer_finished_tiles(self.h))

    def copy_bitmap(self, width, height):
        out = np.zeros((height, width), np.uint32)
        self.lib.rtm_renderer_copy_bitmap(self.h, out.ctypes.data_as(_U32P))
        return out

    def set_alternates(self, ortho_width=0.0, shade_mode=0):
        """Renderer::SetOrthographicWidth / SetShadingMode (0 width = perspective; mode 0 = the live shading)."""
        self.lib.rtm_renderer_set_alternates(self.h, float(ortho_width), int(shade_mode))

    def last_kernel_ms(self):
        return float(self.lib.rtm_renderer_last_kernel_ms(self.h))

    def save_bmp(self, path):
        self.lib.rtm_renderer_save_bmp(self.h, os.fsencode(path))

    def intersect(self, origin, direction):
        o = np.ascontiguousarray(origin, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        tuv = np.zeros(3, np.float32)
        idx = C.c_uint32(capi.MISS)
        rc = self.lib.rtm_renderer_intersect(self.h, o.ctypes.data_as(_F32P), d.ctypes.data_as(_F32P),
                                             tuv.ctypes.data_as(_F32P), C.byref(idx))
        if rc < 0:
            raise RuntimeError(self.lib.rtm_last_error().decode())
        return bool(rc), tuv, idx.value

    def ray_march(self, origin, direction):
        """Renderer::RayMarch (renderer.cpp:24-41) for one ray -> (hit, t)."""
        o = np.ascontiguousarray(origin, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        t = C.c_float(0.0)
        rc = self.lib.rtm_renderer_ray_march(self.h, o.ctypes.data_as(_F32P), d.ctypes.data_as(_F32P), C.byref(t))
        return bool(rc), np.float32(t.value)

    def grid_info(self):
        dim = np.zeros(3, np.uint32)
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        cw, refs = C.c_float(), C.c_uint64()
        self.lib.rtm_renderer_grid_info(self.h, dim.ctypes.data_as(_U32P), mn.ctypes.data_as(_F32P),
                                        mx.ctypes.data_as(_F32P), C.byref(cw), C.byref(refs))
        return dict(dim=dim, aabb_min=mn, aabb_max=mx, cell_wdh=np.float32(cw.value), num_refs=int(refs.value))
