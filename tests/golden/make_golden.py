"""Generates tests/golden/csim_*.npz by running the UNMODIFIED reference C simulation
(oracle/_ref, built from /root/reference by oracle/Makefile) on small seeded matrices.

Run here (the reference tree must be present):  python tests/golden/make_golden.py
Each fixture holds the CSR, the dense vector, and for every implementation the packed result
words that the reference's `top_wrapper` (spmv_csim/csim.cpp:22-136) wrote, together with the
value words / channel geometry used. The GPU parity tests replay them through the C ABI.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hisparse_b200 import matgen  # noqa: E402
from oracle import hsoracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_reference(ref, port, impl, mat, x, skip):
    rows, cols, indptr, indices, data = mat
    IF, P, NCH = ref.INTERLEAVE_FACTOR, ref.PACK_SIZE, ref.NUM_HBM_CHANNELS
    OB, VB = ref.LOGICAL_OB_SIZE, ref.LOGICAL_VB_SIZE
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, P * NCH * IF, P)
    # value conversion and formatting are done by the reference's own code
    words = ref.val_from_float(data)
    cp = ref.csr2cpsr(rows, cols, indptr, indices, data, skip)
    # channel images: oracle restatement of sw/host.cpp:163-231 over blocks that were verified
    # identical to the reference's (tests/test_oracle_vs_ref.py)
    kind = hsoracle.VAL_Q824 if impl == "fixed" else hsoracle.VAL_FLOAT_BITS
    m = port.csr2cpsr(r2, c2, ip2, indices, words, P, OB, VB, NCH * IF, skip, kind)
    for j in range(cp.n_row_parts):
        for i in range(cp.n_col_parts):
            for c in range(NCH * IF):
                assert np.array_equal(m.block(j, i, c)[0], cp.block(j, i, c)[0])
                assert np.array_equal(m.block(j, i, c)[1], cp.block(j, i, c)[1])
    images = m.channel_images(IF)
    xpad = np.zeros(c2, np.float32)
    xpad[:cols] = x
    xw = ref.val_from_float(xpad)
    y = np.zeros(r2, np.uint32)
    for rp in range(m.n_row_parts):
        rows_here = OB if (rp < m.n_row_parts - 1 or r2 % OB == 0) else r2 % OB
        ref.top_wrapper(images, xw, y, rp, rows_here // NCH, m.n_col_parts, m.n_col_parts * m.n_row_parts, c2)
    out = dict(rows_padded=r2, cols_padded=c2, y=y)
    if impl == "fixed":           # float words are just the IEEE bits of the inputs
        out.update(val_words=words, x_words=xw)
    return out


def main():
    hsoracle.build()
    port = hsoracle.Port()
    refs = {i: hsoracle.Ref(i) for i in hsoracle.IMPLS}
    cases = {}
    rng = np.random.default_rng(20261017)
    # 1: random, U(0,1) values: exercises product rounding
    cases["rand"] = (matgen.random_csr(320, 1500, 0.02, 101), "u01", False)
    # 2: power law with empty rows, skip-empty-rows markers, two column partitions
    cases["rmat_skip"] = (matgen.rmat_csr(1500, 12000, 102)[:5], "u01", True)
    big = matgen.random_csr(200, 36000, 0.002, 103)
    cases["two_col_parts"] = (big, "u01", False)
    # 3: rows whose fixed-point sum saturates at 255.99999994 and products that saturate
    r, c, ip, ix, d = matgen.random_csr(128, 1024, 0.1, 104)
    rowid = np.repeat(np.arange(r), np.diff(ip.astype(np.int64)))
    d = (d * np.where(rowid % 2 == 0, 40.0, 0.02)).astype(np.float32)   # even rows saturate, odd rows do not
    cases["saturating"] = ((r, c, ip, ix, d), "big", False)
    for name, (mat, xkind, skip) in cases.items():
        rows, cols, indptr, indices, data = mat
        x = rng.random(cols, dtype=np.float32)
        if xkind == "big":
            x = (x * 30).astype(np.float32)
        out = dict(rows=rows, cols=cols, indptr=indptr, indices=indices, data=data, x=x, skip=int(skip))
        for impl in hsoracle.IMPLS:
            xi = x
            di = data
            if impl != "fixed" and name != "saturating":
                # float variants take signed data
                di = (data - np.float32(0.5)).astype(np.float32)
                xi = (x * 2 - 1).astype(np.float32)
                out["data_" + impl] = di
                out["x_" + impl] = xi
            res = run_reference(refs[impl], port, impl, (rows, cols, indptr, indices, di), xi, skip)
            for k, v in res.items():
                out["%s_%s" % (impl, k)] = v
        path = os.path.join(OUT, "csim_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
