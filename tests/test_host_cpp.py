"""CPU: the C++ host mirror of the reference's spmv::io API (hisparse_b200/host/) against the
reference's golden vectors (unit_tests/test_io.cpp) and against the oracle restatement (itself
pinned to the reference formatter): CSR->CPSR blocks, per-channel packet images, .npz loader."""
import os
import subprocess

import numpy as np
import pytest

from hisparse_b200 import matgen
from oracle import hsoracle
from tests.test_oracle_golden import CASES, M

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hisparse_b200", "host")


@pytest.fixture(scope="module")
def host_bins():
    from hisparse_b200 import capi
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "bin")


def run_format_check(bins, rows, cols, indptr, indices, words, P, C, IF, OB, VB, skip, kind):
    toks = [rows, cols, len(indices), P, C, IF, OB, VB, int(skip), kind]
    text = " ".join(map(str, toks)) + "\n" + " ".join(map(str, indptr)) + "\n" + " ".join(map(str, indices)) + \
        "\n" + " ".join(map(str, words)) + "\n"
    out = subprocess.run([os.path.join(bins, "format_check")], input=text, capture_output=True, text=True, check=True).stdout
    lines = out.split("\n")
    nrp, ncp = map(int, lines[0].split())
    blocks, images = {}, {}
    k = 1
    while k < len(lines) and lines[k]:
        f = lines[k].split()
        if f[0] == "B":
            j, i, c, n, nptr = map(int, f[1:])
            body = np.array(lines[k + 1].split(), dtype=np.uint64).astype(np.uint32).reshape(n, P, 2) if n else np.zeros((0, P, 2), np.uint32)
            ptr = np.array(lines[k + 2].split(), dtype=np.uint64).astype(np.uint32).reshape(nptr, P)
            blocks[(j, i, c)] = (body[:, :, 0], body[:, :, 1], ptr)
            k += 3
        else:
            c, n = int(f[1]), int(f[2])
            images[c] = np.array(lines[k + 1].split(), dtype=np.uint64).astype(np.uint32).reshape(n, 16) if n else np.zeros((0, 16), np.uint32)
            k += 2
    return nrp, ncp, blocks, images


@pytest.mark.parametrize("name,m,ob,vb,nch,skip,gold", CASES, ids=[c[0] for c in CASES])
def test_cpp_csr2cpsr_golden(host_bins, name, m, ob, vb, nch, skip, gold):
    _, _, blocks, _ = run_format_check(host_bins, m["rows"], m["cols"], m["indptr"], m["indices"], m["data"], 2, nch, 1,
                                       ob, vb, skip, 0)
    for key, (data, indices, indptr) in gold.items():
        idx, val, ptr = blocks[key]
        assert idx.tolist() == indices and val.tolist() == data and ptr.tolist() == indptr


@pytest.mark.parametrize("impl,IF", [("fixed", 1), ("float_pob", 1), ("float_stall", 8)])
@pytest.mark.parametrize("skip", [False, True])
def test_cpp_formatter_and_channel_images_match_oracle(host_bins, port, impl, IF, skip):
    rows, cols, indptr, indices, data = matgen.rmat_csr(1500, 9000, 33)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    OB, VB = 128 * IF * 4, 512                      # several row and column partitions
    if impl == "fixed":
        words, kind, okind = port.quantize(data), 2, hsoracle.VAL_Q824
    else:
        words, kind, okind = data.view(np.uint32), 1, hsoracle.VAL_FLOAT_BITS
    nrp, ncp, blocks, images = run_format_check(host_bins, r2, c2, ip2, indices, words, 8, 16, IF, OB, VB, skip, kind)
    want = port.csr2cpsr(r2, c2, ip2, indices, words, 8, OB, VB, 16 * IF, skip, okind)
    assert (nrp, ncp) == (want.n_row_parts, want.n_col_parts)
    for (j, i, c), (idx, val, ptr) in blocks.items():
        a, b, p = want.block(j, i, c)
        assert np.array_equal(idx, a) and np.array_equal(val, b) and np.array_equal(ptr, p)
    for c, im in enumerate(want.channel_images(IF)):
        assert np.array_equal(images[c], im), c


def test_cpp_npz_loader(host_bins, tmp_path):
    rows, cols, indptr, indices, data = matgen.random_csr(300, 500, 0.05, 5)
    for compressed, idt in ((False, np.int32), (True, np.int64)):
        path = str(tmp_path / ("m%d.npz" % compressed))
        save = np.savez_compressed if compressed else np.savez
        save(path, shape=np.array([rows, cols], np.int64), data=data, indices=indices.astype(idt),
             indptr=indptr.astype(idt), format=np.array("csr"))
        out = subprocess.run([os.path.join(host_bins, "npz_check"), path], capture_output=True, text=True).stdout.split()
        assert out[0] != "ERROR", out
        assert (int(out[0]), int(out[1]), int(out[2])) == (rows, cols, indices.size)
        assert abs(float(out[3]) - float(data.astype(np.float64).sum())) < 1e-2
        assert int(out[4]) == int(indices.astype(np.uint64).sum()) and int(out[5]) == int(indptr.astype(np.uint64).sum())


def test_csim_harness_binaries_call_the_library():
    """oracle/_ref/csim_gpu_*: the unmodified reference harness whose top_wrapper is ours. Without a GPU the
    call must fail loudly (no CPU fallback behind the drop-in symbol)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "csim_gpu_fixed")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/csim_gpu_fixed not built")
    syms = subprocess.run(["nm", "-u", exe], capture_output=True, text=True, check=True).stdout
    assert "hsb_top_wrapper" in syms
    from hisparse_b200 import capi
    if capi.device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "top_wrapper on the GPU failed" in r.stderr


def test_cpp_sharding_matches_python(tmp_path):
    """hisparse_b200/host/sharding.h (the C++ multi-GPU driver's row-block cut) == hisparse_b200/sharding.py"""
    from hisparse_b200 import sharding
    rows, cols, indptr, indices, data = matgen.rmat_csr(5000, 90000, 23)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    src = tmp_path / "shard.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include "sharding.h"
int main(int argc, char **argv) {
    std::vector<uint32_t> ip;
    for (unsigned v; std::scanf("%u", &v) == 1;) ip.push_back(v);
    for (int world = 1; world <= 5; world++) {
        for (uint32_t b : spmv::shard::shard_bounds(ip, world)) std::printf("%u ", b);
        std::printf("\n");
    }
    spmv::io::CSRMatrix<float> m;
    m.num_rows = (uint32_t)ip.size() - 1; m.num_cols = 8; m.adj_indptr = ip;
    m.adj_indices.assign(ip.back(), 1u); m.adj_data.assign(ip.back(), 2.0f);
    auto s = spmv::shard::extract_shard(m, 128, 384);
    std::printf("%u %u %u %zu\n", s.num_rows, s.adj_indptr.front(), s.adj_indptr.back(), s.adj_indices.size());
    return 0;
}
''')
    exe = tmp_path / "shard"
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", HOST, str(src), "-o", str(exe), "-lz"], check=True)
    out = subprocess.run([str(exe)], input=" ".join(str(int(v)) for v in ip2), capture_output=True, text=True,
                         check=True).stdout.splitlines()
    for world in range(1, 6):
        assert [int(v) for v in out[world - 1].split()] == sharding.shard_bounds(ip2, world)
    n = int(ip2[384]) - int(ip2[128])
    assert [int(v) for v in out[5].split()] == [256, 0, n, n]
