"""The oracle against the reference's OWN known-answer vectors.

Golden vectors: /root/reference/unit_tests/test_io.cpp (written for GraphLily's gtest suite;
the file is not buildable in the reference repo any more, but the vectors are the only
known-answer fixtures the reference holds for the CSR -> CPSR path):
    :45-65, :361-367   the three tiny matrices
    :121-130           util_round_csr_matrix_dim
    :143-174           util_convert_csr_to_dds (column partitioning)
    :206-245           util_pack_rows
    :248-306           csr2cpsr, column partitioning
    :309-349           csr2cpsr, row partitioning
    :370-398           csr2cpsr, skip-empty-rows markers
Each vector is checked against (a) our C restatement (oracle/hsoracle.c) and, when oracle/_ref
exists, (b) the unmodified reference formatter compiled at the same pack_size (Ref2).
The goldens store the end-of-row marker VALUE as the number k ("advance k slots"); that is what
the current formatter produces for an integer DataT (sw/data_formatter.h:69-74), so the integer
instantiation is the one compared bit-for-bit.
"""
import os

import numpy as np
import pytest

from oracle import hsoracle

M = 0xFFFFFFFF

CSR1 = dict(rows=4, cols=4, data=[1, 2, 3, 4, 5, 6, 7, 8], indices=[0, 1, 2, 3, 0, 2, 1, 3], indptr=[0, 4, 6, 7, 8])
CSR2 = dict(rows=4, cols=8, data=[1, 2, 3, 4, 1, 2, 3, 4, 5, 6, 5, 6, 7, 7, 8, 8],
            indices=[0, 1, 2, 3, 4, 5, 6, 7, 0, 2, 4, 6, 1, 5, 3, 7], indptr=[0, 8, 12, 14, 16])
CSR3 = dict(rows=8, cols=4, data=[1, 2, 3, 4, 5], indices=[0, 2, 0, 1, 0], indptr=[0, 0, 2, 3, 3, 3, 4, 5, 5])


def _port_blocks(port, m, ob, vb, nch, skip):
    h = port.csr2cpsr(m["rows"], m["cols"], m["indptr"], m["indices"], m["data"], 2, ob, vb, nch, skip,
                      hsoracle.VAL_INT)
    return h


def _ref2_available():
    return os.path.exists(os.path.join(hsoracle.HERE, "_ref", "libref_fmt2.so"))


def _check(block, data, indices, indptr):
    idx, val, ptr = block
    assert idx.tolist() == indices
    assert val.astype(np.int64).tolist() == data
    assert ptr.tolist() == indptr


GOLD_COL = {  # test_io.cpp:262-278 : (j, i, c) -> data, indices, indptr
    (0, 0, 0): ([[1, 5], [2, 6], [3, 1], [4, 0], [1, 0]], [[0, 0], [1, 2], [2, M], [3, 0], [M, 0]], [[0, 0], [5, 3]]),
    (0, 0, 1): ([[7, 8], [1, 1]], [[1, 3], [M, M]], [[0, 0], [2, 2]]),
}
GOLD_COL[(0, 1, 0)] = GOLD_COL[(0, 0, 0)]
GOLD_COL[(0, 1, 1)] = GOLD_COL[(0, 0, 1)]

GOLD_ROW = {  # test_io.cpp:323-335
    (0, 0, 0): ([[1, 5], [2, 6], [3, 1], [4, 0], [1, 0]], [[0, 0], [1, 2], [2, M], [3, 0], [M, 0]], [[0, 0], [5, 3]]),
    (1, 0, 0): ([[7, 8], [1, 1]], [[1, 3], [M, M]], [[0, 0], [2, 2]]),
}

GOLD_SKIP = {  # test_io.cpp:385-390
    (0, 0, 0): ([[1, 1], [3, 2], [2, 2], [5, 4], [1, 2]], [[M, 0], [0, 2], [M, M], [0, 1], [M, M]],
                [[0, 0], [1, 3], [3, 3], [3, 5], [5, 5]]),
}

CASES = [("col", CSR2, 4, 4, 2, False, GOLD_COL), ("row", CSR1, 2, 4, 1, False, GOLD_ROW),
         ("skip", CSR3, 8, 4, 1, True, GOLD_SKIP)]


@pytest.mark.parametrize("name,m,ob,vb,nch,skip,gold", CASES, ids=[c[0] for c in CASES])
def test_port_csr2cpsr_golden(port, name, m, ob, vb, nch, skip, gold):
    h = _port_blocks(port, m, ob, vb, nch, skip)
    for (j, i, c), (data, indices, indptr) in gold.items():
        _check(h.block(j, i, c), data, indices, indptr)


@pytest.mark.skipif(not _ref2_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,m,ob,vb,nch,skip,gold", CASES, ids=[c[0] for c in CASES])
def test_reference_csr2cpsr_golden(name, m, ob, vb, nch, skip, gold):
    """The reference formatter itself still reproduces its golden vectors (integer DataT)."""
    r2 = hsoracle.Ref2()
    h = r2.csr2cpsr(m["rows"], m["cols"], m["indptr"], m["indices"], m["data"], ob, vb, nch, skip, "i32")
    for (j, i, c), (data, indices, indptr) in gold.items():
        _check(r2.block(h, j, i, c, "i32"), data, indices, indptr)


@pytest.mark.skipif(not _ref2_available(), reason="oracle/_ref not built")
def test_reference_float_marker_is_bit_pattern():
    """With DataT=float the marker slots hold the BIT PATTERN of k (data_formatter.h:69-74),
    which the float loader decodes with val2bit (spmv-fp/libfpga/spmv_cluster.h:104)."""
    r2 = hsoracle.Ref2()
    m = CSR3
    h = r2.csr2cpsr(m["rows"], m["cols"], m["indptr"], m["indices"], m["data"], 8, 4, 1, True, "f32")
    idx, val, _ = r2.block(h, 0, 0, 0, "f32")
    bits = val.view(np.uint32)
    want = np.array(GOLD_SKIP[(0, 0, 0)][0])
    mk = idx == M
    assert bits[mk].tolist() == want[mk].tolist()
    assert val[~mk].tolist() == want[~mk].astype(np.float32).tolist()


def test_round_dims_golden(port):
    assert port.round_dims(4, 4, 3, 5) == (6, 5)          # test_io.cpp:121-130
    if _ref2_available():
        assert hsoracle.Ref2().round_dims(4, 4, 3, 5) == (6, 5)


@pytest.mark.skipif(not _ref2_available(), reason="oracle/_ref not built")
def test_reference_dds_and_pack_rows_golden():
    r2 = hsoracle.Ref2()
    m = CSR1
    d0, i0, p0 = r2.csr_to_dds(4, 4, m["indptr"], m["indices"], m["data"], 3, 0)   # test_io.cpp:160-165
    d1, i1, p1 = r2.csr_to_dds(4, 4, m["indptr"], m["indices"], m["data"], 3, 1)
    assert d0.tolist() == [1, 2, 3, 5, 6, 7] and i0.tolist() == [0, 1, 2, 0, 2, 1] and p0.tolist() == [0, 3, 5, 6, 6]
    assert d1.tolist() == [4, 8] and i1.tolist() == [0, 0] and p1.tolist() == [0, 1, 1, 1, 2]
    idx, val, ptr = r2.pack_rows(m["indptr"], m["indices"], m["data"], 2, 0)        # test_io.cpp:226-231
    assert val.tolist() == [[1, 5], [2, 6], [3, 0], [4, 0]] and idx.tolist() == [[0, 0], [1, 2], [2, 0], [3, 0]]
    assert ptr.tolist() == [[0, 0], [4, 2]]
    idx, val, ptr = r2.pack_rows(m["indptr"], m["indices"], m["data"], 2, 1)
    assert val.tolist() == [[7, 8]] and idx.tolist() == [[1, 3]] and ptr.tolist() == [[0, 0], [1, 1]]
