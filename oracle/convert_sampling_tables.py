#!/usr/bin/env python
"""Dump the Braaten-Weller digit permutations (first 16 primes, 381 entries) from the reference's
sampling module (sampling.cpp:55-98, via oracle/_ref) to assets/sampling/braaten_weller_16.u32.
The table is published data (Vandewoestyne & Cools); it is an asset like the meshes, because the
device QMC generator takes permutation tables as input.   python oracle/convert_sampling_tables.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

if __name__ == "__main__":
    tab = po.Ref.get().qmc_tables(1, 16)
    assert len(tab) == 381
    out = os.path.join(ROOT, "assets", "sampling", "braaten_weller_16.u32")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tab.astype(np.uint32).tofile(out)
    print("wrote", out, len(tab), "entries")
