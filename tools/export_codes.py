#!/usr/bin/env python3
"""Re-export the reference's committed code definitions into this repo.

Run in the authoring container (where /root/reference is mounted).  The GPU
box has no /root/reference, so everything the tests / bench / library need at
run time is written here:

  ldpc_erasure_codes_b200/codes/<name>.mat   MAT-v5 files holding `H_sparse`
        (sparse double, zlib-compressed) -- same variable, class and values as
        the reference's Matlab/*.mat, re-serialised by scipy (header text and
        zlib stream differ; content identical).  libldpc_cuda's loader reads
        both these and the originals.
  tests/golden/gf256_tables.npz              the reference's GF(2^8) add / mul /
        inverse tables (Matlab/GF_256_add_mult_inv_tables.mat), used to pin the
        oracle's and the library's field (polynomial 0x171).
  tests/golden/codes_digest.json             shape / nnz / crc32 of the CSR
        arrays, so tests can tell a stale export.
"""
import json
import os
import sys
import zlib

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

REF = "/root/reference/Matlab"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = {
    # name in this repo                 reference file                                             k
    "n2000_k1000": ("n2000_k1000_no6cycles_triangleForm_OpenCL_H.mat", 1000),
    "n2040_k1530": ("n2040_k1530_irreg_H_no6cycles_triangleForm.mat", 1530),
    "n4000_k2000": ("n4000_k2000_no6cycles_triangleForm.mat", 2000),
}


def main():
    out_dir = os.path.join(ROOT, "ldpc_erasure_codes_b200", "codes")
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(gold, exist_ok=True)
    digest = {}
    for name, (ref_file, k) in CODES.items():
        H = sio.loadmat(os.path.join(REF, ref_file), spmatrix=True)["H_sparse"]
        H = sp.csc_matrix(H)
        H.sort_indices()
        assert np.all(H.data == 1.0)
        m, n = H.shape
        assert n - m == k
        sio.savemat(os.path.join(out_dir, name + ".mat"), {"H_sparse": H.astype(np.float64)},
                    do_compression=True, format="5")
        R = H.tocsr()
        R.sort_indices()
        digest[name] = {
            "n": int(n), "k": int(k), "m": int(m), "nnz": int(H.nnz),
            "crc32_row_ptr": zlib.crc32(R.indptr.astype("<i4").tobytes()),
            "crc32_col_idx": zlib.crc32(R.indices.astype("<i4").tobytes()),
            "reference_file": "Matlab/" + ref_file,
        }
    g = sio.loadmat(os.path.join(REF, "GF_256_add_mult_inv_tables.mat"))
    np.savez_compressed(os.path.join(gold, "gf256_tables.npz"),
                        add=g["GF_add_lookup"].astype(np.uint8),
                        mul=g["GF_mult_lookup"].astype(np.uint8),
                        inv=g["GF_inv_lookup"].astype(np.uint8).reshape(-1))
    with open(os.path.join(gold, "codes_digest.json"), "w") as f:
        json.dump(digest, f, indent=1, sort_keys=True)
    print(json.dumps(digest, indent=1))


if __name__ == "__main__":
    sys.exit(main())
