"""CPU: the oracle restatement reproduces the committed fixtures that the reference C simulation
generated (tests/golden/make_golden.py). Bit-exact for fixed, 1e-5 norm-wise for float."""
import glob
import os

import numpy as np
import pytest

from hisparse_b200 import matgen
from oracle import hsoracle

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "csim_*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_port_reproduces_reference_fixture(port, path):
    g = np.load(path)
    rows, cols = int(g["rows"]), int(g["cols"])
    indptr, indices = g["indptr"], g["indices"]
    # fixed: quantisation, then closed-form Q8.24 SpMV
    words = port.quantize(g["data"])
    assert np.array_equal(words, g["fixed_val_words"])
    r2, c2 = int(g["fixed_rows_padded"]), int(g["fixed_cols_padded"])
    _, _, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    xw = np.zeros(c2, np.uint32)
    xw[:cols] = port.quantize(g["x"])
    assert np.array_equal(xw, g["fixed_x_words"])
    y = port.spmv_q824(ip2, indices, words, xw)
    assert y.size == r2 and np.array_equal(y, g["fixed_y"])
    if "saturating" in path:
        assert (y == 0xFFFFFFFF).any() and (y != 0xFFFFFFFF).any()
    for impl in ("float_pob", "float_stall"):
        d = g["data_" + impl] if ("data_" + impl) in g else g["data"]
        x = g["x_" + impl] if ("x_" + impl) in g else g["x"]
        rp = int(g[impl + "_rows_padded"])
        _, _, ipp = matgen.pad_csr(rows, cols, indptr, rp - rows + 1 if rp == rows else rp, 8) if False else matgen.pad_csr(rows, cols, indptr, 1024 if impl == "float_stall" else 128, 8)
        y64, sa = port.spmv_f64(ipp, indices, d, np.concatenate([x, np.zeros(8, np.float32)]))
        yref = g[impl + "_y"].view(np.float32).astype(np.float64)
        assert np.all(np.abs(yref - y64) <= 1e-5 * sa + 1e-30)
