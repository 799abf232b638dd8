#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref; needs /root/reference built).

The reference ships no tests or golden vectors, so these are outputs of the
reference itself: for every scene preset at a reduced size, and for BASELINE config C1 at full size,
the md5 of the rendered BGRA image (reference Renderer + worker pool), the md5 of the per-sample hit
triangle indices and of t/u/v (reference GenerateRay + Grid::Intersect), hit counts, and the grid
digest (dims, cell width bits, md5 of offsets / triangle lists).  One small case is stored in full
(cornell 48x48x2: image + hit indices) so a failure can be localised.

    python tests/golden/make_golden.py      ->  tests/golden/ref_digests.json, tests/golden/cornell_48x48x2.npz
    python tests/golden/make_golden.py --missing   only the cases ref_digests.json does not hold yet

Also records the md5 of the BMP file the reference's own SaveToBMP writes for each case (`bmp_md5`): SURVEY.md
section 8c quotes those for C1-C4 and killeroo 4K.
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

scenes = importlib.import_module("cpp-11-ray-trace-march-framework_b200.scenes")

CASES = [(n, 160, 96, 4, 64) for n in ("torusknot", "cornell", "room", "table_chair", "head", "room_cat",
                                       "water_knot", "griebel_teapot", "killeroo", "dwarf_hand_blob")]
CASES += [("cornell", 512, 512, 1, 64),          # BASELINE config C1, full size
          ("killeroo", 480, 270, 4, 64),          # C2 at quarter resolution
          ("torusknot", 240, 135, 16, 64),        # C3 scaled
          ("room", 240, 135, 16, 64),             # C4 scaled
          ("tiger_soup_small", 160, 90, 4, 48),   # C5 construction at test size
          ("killeroo", 97, 61, 3, 33),            # odd everything
          ("killeroo", 1920, 1080, 4, 64),        # BASELINE config C2, FULL size
          ("torusknot", 1920, 1080, 16, 64),      # BASELINE config C3, FULL size
          ("room", 3840, 2160, 16, 64),           # BASELINE config C4, FULL size
          ("killeroo", 3840, 2160, 16, 64),       # the north-star scene = bench.py's default workload, FULL size
          ("tiger_soup_medium", 1920, 1080, 4, 160)]  # C5's construction at 3 840 instances / 2.3 M triangles


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = po.Ref.get()
    out = {}
    path = os.path.join(ROOT, "tests", "golden", "ref_digests.json")
    if "--missing" in sys.argv and os.path.exists(path):
        with open(path) as f:
            out = json.load(f)
    for name, w, h, spp, res in CASES:
        if "%s_%dx%dx%d_g%d" % (name, w, h, spp, res) in out:
            continue
        m, fov, cam = scenes.build(ref.api, name)
        r = ref.renderer(m, fov, cam, res)
        _, img = r.render(w, h, spp)
        bmp = "/tmp/_golden_%d.bmp" % os.getpid()
        r.save_bmp(bmp)  # the reference's own Framebuffer::SaveToBMP of the frame just rendered
        with open(bmp, "rb") as f:
            bmp_md5 = hashlib.md5(f.read()).hexdigest()
        os.remove(bmp)
        idx, t, u, v = r.trace_hits(w, h, spp)
        g = r.grid()
        key = "%s_%dx%dx%d_g%d" % (name, w, h, spp, res)
        out[key] = dict(scene=name, width=w, height=h, spp=spp, grid_res=res,
                        image_md5=md5(img), bmp_md5=bmp_md5, tri_md5=md5(idx), t_md5=md5(t), u_md5=md5(u), v_md5=md5(v),
                        hits=int((idx != 0xFFFFFFFF).sum()), dim=[int(x) for x in g["dim"]],
                        cell_wdh_bits=int(np.float32(g["cell_wdh"]).view(np.uint32)),
                        refs=int(len(g["tri_index"])), cell_offset_md5=md5(g["cell_offset"]),
                        tri_index_md5=md5(g["tri_index"]))
        print(key, out[key]["image_md5"], out[key]["hits"])
        if (name, w, spp) == ("cornell", 160, 4):
            pass
    m, fov, cam = scenes.build(ref.api, "cornell")
    r = ref.renderer(m, fov, cam, 64)
    _, img = r.render(48, 48, 2)
    idx, t, u, v = r.trace_hits(48, 48, 2)
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    np.savez_compressed(os.path.join(gdir, "cornell_48x48x2.npz"), image=img, tri=idx, t=t, u=u, v=v)
    with open(os.path.join(gdir, "ref_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
