#!/usr/bin/env python3
"""Make a girth-8 triangular-form LDPC code (the library's counterpart of
Matlab/Hgen_irregularDegree_no6cycles_systematic_encoding.m) and save it as the MAT-v5 `H_sparse` file the loader reads.

  python tools/hgen.py --preset n2000_k1000 -o /tmp/n2000.mat
  python tools/hgen.py --checks 1000x6 --vars 2000x3 --seed 4 -o my.mat      # count x degree, degrees descending

Needs no GPU."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ldpc_erasure_codes_b200 import hgen

PRESETS = {   # the profiles listed in the reference script (:22-42)
    "n204_k102": ([(102, 6)], [(204, 3)]),
    "n1200_k600": ([(600, 8)], [(1200, 4)]),
    "n2000_k1500": ([(500, 12)], [(2000, 3)]),
    "n2000_k1000": ([(1000, 6)], [(2000, 3)]),
    "n4000_k2000": ([(2000, 6)], [(4000, 3)]),
    "n2040_k1530": ([(479, 13), (31, 12)], [(145, 10), (1389, 3), (476, 2), (30, 1)]),
    "n4080_k3060": ([(682, 15), (338, 14)], [(422, 12), (2638, 3), (964, 2), (56, 1)]),
}


def prof(text):
    return [tuple(int(x) for x in part.split("x")) for part in text.split(",")]


ap = argparse.ArgumentParser()
ap.add_argument("--preset", choices=sorted(PRESETS))
ap.add_argument("--checks", type=prof, help="check profile, e.g. 479x13,31x12")
ap.add_argument("--vars", type=prof, help="variable profile, e.g. 145x10,1389x3,476x2,30x1")
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-tries", type=int, default=100)
ap.add_argument("-o", "--out", required=True)
a = ap.parse_args()
c, v = PRESETS[a.preset] if a.preset else (a.checks, a.vars)
t0 = time.time()
H, tries = hgen.generate(c, v, seed=a.seed, max_tries=a.max_tries)
n4, n6 = hgen.count_short_cycles(H)
hgen.save_mat(a.out, H)
rw = np.bincount(np.diff(H.indptr))
cw = np.bincount(np.asarray(H.sum(axis=0)).ravel().astype(int))
print(f"{a.out}: H {H.shape[0]} x {H.shape[1]}, nnz {H.nnz}, {tries} tries, {time.time() - t0:.1f} s; variables on 4-cycles {n4}, on 6-cycles {n6}; "
      f"row weights {dict((i, int(x)) for i, x in enumerate(rw) if x)}, column weights {dict((i, int(x)) for i, x in enumerate(cw) if x)}")
