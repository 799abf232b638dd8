#!/usr/bin/env python
"""Convert the reference's ASCII meshes (/root/reference/meshes/*.dat) to `.meshbin` assets.

/root/reference does not exist on the GPU box, so the scenes BASELINE.json names must travel
with the repo.  Each asset is the in-memory state of the reference's `Mesh` right after
`Mesh::Read` (mesh.cpp:138-391), produced by the reference's own parser through
oracle/_ref/libref_oracle.so -- i.e. before NormalizeDimensions / Transform / AddMesh, which each
implementation (reference, oracle port, host library) then applies itself.

Layout (little endian):  8-byte magic b"RTMMESH1", uint32 num_vertices, uint32 num_triangles,
num_vertices x {float32 px,py,pz,nx,ny,nz}   (= Mesh::Vertex,   mesh.h:20-24, 24 B)
num_triangles x {uint32 v0,v1,v2, float32 nx,ny,nz} (= Mesh::Triangle, mesh.h:12-18, 24 B)

`table_chair` is also stored with flip_winding=true (scene 3, application.cpp:352).
Run:  python oracle/convert_meshes.py        (needs `make -C oracle ref` first)
"""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MESHES = "/root/reference/meshes"
OUT = os.path.join(ROOT, "assets", "meshes")


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so"))
    lib.ref_mesh_new.restype = C.c_void_p
    lib.ref_mesh_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.ref_mesh_num_vertices.argtypes = [C.c_void_p]
    lib.ref_mesh_num_triangles.argtypes = [C.c_void_p]
    lib.ref_mesh_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_mesh_free.argtypes = [C.c_void_p]
    os.makedirs(OUT, exist_ok=True)
    jobs = [(f[:-4], False) for f in sorted(os.listdir(REF_MESHES)) if f.endswith(".dat")]
    jobs.append(("table_chair", True))
    total = 0
    for name, flip in jobs:
        m = lib.ref_mesh_new()
        ok = lib.ref_mesh_read(m, os.path.join(REF_MESHES, name + ".dat").encode(), int(flip))
        if not ok:
            print("FAILED to read", name)
            sys.exit(1)
        nv, nt = lib.ref_mesh_num_vertices(m), lib.ref_mesh_num_triangles(m)
        vtx = np.empty((nv, 6), np.float32)
        tri = np.empty((nt, 6), np.uint32)
        lib.ref_mesh_get(m, vtx.ctypes.data, tri.ctypes.data)
        lib.ref_mesh_free(m)
        out = os.path.join(OUT, name + (".flip" if flip else "") + ".meshbin")
        with open(out, "wb") as f:
            f.write(b"RTMMESH1" + struct.pack("<II", nv, nt))
            f.write(vtx.tobytes())
            f.write(tri.tobytes())
        total += os.path.getsize(out)
        print(f"{name:40s} flip={int(flip)} V={nv:7d} T={nt:7d} -> {os.path.getsize(out)} B")
    print("total bytes", total)


if __name__ == "__main__":
    main()
