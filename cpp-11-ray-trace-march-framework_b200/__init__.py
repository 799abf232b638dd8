"""B200-native tile tracer: drop-in for the per-tile tracing hot path of
blitzcode/cpp-11-ray-trace-march-framework (Renderer::RenderTile and everything under it).

* ``capi``    -- ctypes binding of the C ABI (include/cuda_trace.h, libcuda_trace.so)
* ``meshapi`` -- ctypes view of the host library's Mesh / Matrix44f mirror
* ``scenes``  -- the reference viewer's scene presets as data
* ``build``   -- in-tree nvcc / g++ build of the native libraries

The package directory name is not a Python identifier; import it with
``importlib.import_module("cpp-11-ray-trace-march-framework_b200")``.
"""
