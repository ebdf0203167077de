"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
include/hisparse_b200.h declares, the tile-stream format round-trips, and reference channel
images decode back to the CSR they were built from. No compute calls (no GPU here)."""
import numpy as np
import pytest

from hisparse_b200 import capi, matgen
from oracle import hsoracle


def test_library_exports_every_declared_symbol():
    capi.build()
    L = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.hsb_version()


def test_config_matches_reference_constants(refs):
    for name, impl in capi.IMPL_BY_NAME.items():
        cfg = capi.get_config(impl)
        r = refs[name]
        assert (cfg.pack_size, cfg.num_hbm_channels, cfg.interleave_factor, cfg.logical_ob_size,
                cfg.logical_vb_size) == (r.PACK_SIZE, r.NUM_HBM_CHANNELS, r.INTERLEAVE_FACTOR,
                                         r.LOGICAL_OB_SIZE, r.LOGICAL_VB_SIZE)


def test_no_cpu_fallback():
    """Without a usable GPU the product refuses to run instead of computing on the CPU."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.HsbError):
        capi.Context(0, capi.IMPL_FIXED)


def _canon(rows, indptr, indices, vals):
    """sort entries inside every row by (column, value) for order-insensitive comparison"""
    r = np.repeat(np.arange(rows, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    o = np.lexsort((vals, indices, r))
    return indices[o], vals[o]


CASES = [
    ("empty", (8, 8, np.zeros(9, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32)), 0, 0),
    ("dense", matgen.dense_csr(64, 72), 0, 0),
    ("uniform_unsorted_cols", matgen.uniform_sparse_csr(1000, 1024, 10), 0, 0),
    ("rand_multi_tile", matgen.random_csr(300, 70000, 0.004, 3), 0, 0),
    ("rand_small_tiles", matgen.random_csr(2000, 3000, 0.01, 4), 512, 64),
    ("rmat_row_parts", matgen.rmat_csr(6000, 90000, 5), 1024, 1024),
    ("one_long_row", (4, 40000, np.array([0, 0, 40000, 40000, 40001], np.uint32),
                      np.concatenate([np.arange(40000), [7]]).astype(np.uint32),
                      np.arange(40001, dtype=np.float32)), 0, 0),
]


@pytest.mark.parametrize("name,mat,rpp,tile", CASES, ids=[c[0] for c in CASES])
def test_tile_format_round_trip(name, mat, rpp, tile):
    rows, cols, indptr, indices, data = mat
    f = capi.Format(rows, cols, indptr, indices, data, rpp, tile)
    st = f.stats()
    assert st["nnz"] == indices.size and st["rows"] == rows
    ip, ix, vv = f.expand()
    assert np.array_equal(ip, indptr)
    a = _canon(rows, indptr, indices, np.ascontiguousarray(data).view(np.uint32))
    b = _canon(rows, ip, ix, vv)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # bytes: 6 per stored slot + 4 per lane stream slot + 8 per slice (+ tile table)
    assert st["format_bytes"] >= 6 * st["nnz"] and st["n_elems"] >= st["nnz"]
    if st["layout"] == 1:       # narrow: one row unit per slice, padding only up to the longest of 32 sorted streams
        assert st["n_elems"] <= 2 * st["nnz"] + 64 * st["n_slices"]
    assert st["n_streams"] <= max(1, rows * st["n_col_tiles"]) + st["nnz"] // 64 + 1


def test_format_rejects_malformed():
    ip = np.array([0, 2, 1], np.uint32)
    with pytest.raises(capi.HsbError):
        capi.Format(2, 4, ip, np.zeros(2, np.uint32), np.zeros(2, np.float32))
    ip = np.array([0, 1, 2], np.uint32)
    with pytest.raises(capi.HsbError):
        capi.Format(2, 4, ip, np.array([0, 9], np.uint32), np.zeros(2, np.float32))


@pytest.mark.parametrize("impl", hsoracle.IMPLS)
@pytest.mark.parametrize("skip", [False, True])
def test_cpsr_images_decode_to_original_csr(port, impl, skip):
    """reference-format channel images (built by the oracle's restatement of sw/host.cpp:163-231,
    itself pinned against the reference) -> hsb_cpsr_to_csr == the CSR we started from."""
    cfg = capi.get_config(capi.IMPL_BY_NAME[impl])
    IF = cfg.interleave_factor
    rows, cols, indptr, indices, data = matgen.rmat_csr(4000, 50000, 17)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    words = port.quantize(data) if impl == "fixed" else data.view(np.uint32)
    kind = hsoracle.VAL_Q824 if impl == "fixed" else hsoracle.VAL_FLOAT_BITS
    m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, cfg.logical_ob_size, cfg.logical_vb_size, 16 * IF, skip, kind)
    images = m.channel_images(IF)
    ip, ix, vv = capi.cpsr_to_csr(impl, images, m.n_row_parts, m.n_col_parts, r2, c2)
    assert np.array_equal(ip, ip2)
    a = _canon(r2, ip2, indices, words)
    b = _canon(r2, ip, ix, vv)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_cpsr_decode_rejects_truncated_image(port):
    cfg = capi.get_config(0)
    rows, cols, indptr, indices, data = matgen.random_csr(256, 512, 0.05, 9)
    m = port.csr2cpsr(rows, cols, indptr, indices, port.quantize(data), 8, cfg.logical_ob_size,
                      cfg.logical_vb_size, 16, False, hsoracle.VAL_Q824)
    images = m.channel_images(1)
    images[3] = images[3][:-1]
    with pytest.raises(capi.HsbError):
        capi.cpsr_to_csr(0, images, 1, 1, rows, cols)


def test_quantize_helper_matches_oracle(port):
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.random(50000, dtype=np.float32) * 300 - 10,
                        np.array([0, 1, 255.99999, 256, 1e9, -1, 2 ** -25, 3 * 2 ** -25, np.nan, np.inf], np.float32)])
    assert np.array_equal(matgen.quantize_q824(v), port.quantize(v))


# ------------------------------------------------------------------------------------------
# launch planner (host side of hsb_spmv): every step exactly once, balanced, interleaved
# ------------------------------------------------------------------------------------------
def _steps_per_tile(fmt, rec):
    """tile -> sorted list of [first, end) step ranges handed out by the plan"""
    out = {}
    for cta, tile, a, b in rec:
        out.setdefault(int(tile), []).append((int(a), int(b)))
    return {t: sorted(v) for t, v in out.items()}


@pytest.mark.parametrize("name,make,tile_cols,ctas", [
    ("one_tile", lambda: matgen.rmat_csr(3000, 60000, 5), 0, 148),
    ("few_tiles", lambda: matgen.random_csr(900, 70000, 0.004, 7), 16384, 148),
    ("many_tiles", lambda: matgen.random_csr(2000, 400000, 0.0004, 9), 1024, 37),
    ("tiny", lambda: matgen.random_csr(64, 64, 0.2, 3), 0, 148),
])
def test_plan_covers_every_step_once(name, make, tile_cols, ctas):
    rows, cols, indptr, indices, data = make()
    fmt = capi.Format(rows, cols, indptr, indices, hsoracle.Port().quantize(data), 0, tile_cols)
    st = fmt.stats()
    rec = fmt.plan(ctas)
    assert rec.shape[0] > 0 and int(rec[:, 0].max()) < ctas
    total = 0
    for tile, ranges in _steps_per_tile(fmt, rec).items():
        # contiguous, non-overlapping, starting at 0
        pos = 0
        for a, b in ranges:
            assert a == pos and b > a, (name, tile, ranges[:4])
            pos = b
        total += pos
    assert total * (32 if st["layout"] == 1 else 128) == st["n_elems"]      # all stored slots, nothing twice
    # balance: no CTA gets more than ~1.5x the mean number of steps (+ one slice of slack)
    per_cta = np.bincount(rec[:, 0], weights=(rec[:, 3] - rec[:, 2]).astype(np.float64), minlength=ctas)
    busy = per_cta[per_cta > 0]
    assert busy.max() <= 1.5 * busy.mean() + 40


def test_plan_interleaves_tiles_when_there_are_many():
    """with >= 2 tiles per CTA the k-th tile of CTA b is tile k * ctas + b (up to the drift of the equal-cost
    cuts, which can hand a CTA the tail of its neighbour's run first): concurrent CTAs work on neighbouring
    tiles (hypersparse shards)"""
    rows, cols, indptr, indices, data = matgen.random_csr(1500, 600000, 0.0005, 11)
    ctas = 20
    fmt = capi.Format(rows, cols, indptr, indices, hsoracle.Port().quantize(data), 0, 2048)
    assert fmt.stats()["n_col_tiles"] >= 8 * ctas
    rec = fmt.plan(ctas)
    # every CTA walks its tiles upwards in steps of `ctas`
    for c in (0, ctas // 2, ctas - 1):
        tiles = []
        for cta, tile, a, b in rec:
            if int(cta) == c and (not tiles or tiles[-1] != int(tile)):
                tiles.append(int(tile))
        d = np.diff(tiles)
        assert len(tiles) >= 6 and np.median(d) == ctas, (c, tiles[:8])


# ------------------------------------------------------------------------------------------
# the C header as a C compiler sees it: plain C, and the same struct layouts the ctypes glue assumes
# ------------------------------------------------------------------------------------------
def test_header_is_plain_c_and_layouts_match_ctypes(tmp_path):
    import ctypes
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "hisparse_b200.h"
int main(void) {
    printf("stats %zu %zu %zu %zu %zu\n", sizeof(hsb_stats), offsetof(hsb_stats, n_slices), offsetof(hsb_stats, format_bytes),
           offsetof(hsb_stats, kernel_launches), offsetof(hsb_stats, preprocess_seconds));
    printf("config %zu\n", sizeof(hsb_config));
    printf("dcsr %zu %zu %zu %zu\n", sizeof(hsb_device_csr), offsetof(hsb_device_csr, nnz), offsetof(hsb_device_csr, d_indptr),
           offsetof(hsb_device_csr, device));
    printf("blob %d\n", HSB_PEER_BLOB_BYTES);
    return 0;
}
''')
    exe = tmp_path / "abi"
    inc = capi.HEADER.rsplit("/", 1)[0]
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict((l.split()[0], [int(v) for v in l.split()[1:]]) for l in
               subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    S, D = capi.Stats, capi.DeviceCsrStruct
    assert out["stats"] == [ctypes.sizeof(S), S.n_slices.offset, S.format_bytes.offset, S.kernel_launches.offset,
                            S.preprocess_seconds.offset]
    assert out["config"] == [ctypes.sizeof(capi.Config)]
    assert out["dcsr"] == [ctypes.sizeof(D), D.nnz.offset, D.d_indptr.offset, D.device.offset]
    assert out["blob"] == [capi.PEER_BLOB_BYTES]


def test_tile_width_rule():
    """one tile while x fits shared memory; <= 44,000 columns for ordinary wide matrices; up to the full 57,344 when
    there is less than one entry per row and tile (hypersparse: narrow layout)"""
    q = hsoracle.Port().quantize
    r, c, ip, ix, d = matgen.random_csr(256, 57344, 0.002, 1)
    assert capi.Format(r, c, ip, ix, q(d)).stats()["n_col_tiles"] == 1
    r, c, ip, ix, d = matgen.random_csr(600, 107616, 0.01, 2)                  # ~1076 entries per row: dense enough
    st = capi.Format(r, c, ip, ix, q(d)).stats()
    assert st["n_col_tiles"] == 3 and st["tile_cols"] <= 44000
    r, c, ip, ix, d = matgen.random_csr(4000, 400000, 0.000005, 3)             # ~2 entries per row over 10+ tiles
    st = capi.Format(r, c, ip, ix, q(d)).stats()
    assert st["n_col_tiles"] == 7 and st["tile_cols"] <= 57344 and st["layout"] == 1


# ------------------------------------------------------------------------------------------
# the formatter / planner / kernel contract, walked on the host (hsb_format_emulate_fixed)
# ------------------------------------------------------------------------------------------
EMU_CASES = [
    ("dense128", lambda: matgen.dense_csr(128, 128), 0, 0, 148),
    ("uniform_unsorted", lambda: matgen.uniform_sparse_csr(1000, 1024, 10), 0, 0, 148),
    ("rmat_one_tile", lambda: matgen.rmat_csr(6000, 150000, 5), 0, 0, 148),
    ("rmat_few_ctas", lambda: matgen.rmat_csr(6000, 150000, 5), 0, 0, 7),
    ("rmat_two_ctas_per_sm", lambda: matgen.rmat_csr(20000, 500000, 6), 0, 0, 296),
    ("multi_tile", lambda: matgen.random_csr(900, 70000, 0.004, 7), 0, 16384, 148),
    ("many_tiles_interleaved", lambda: matgen.random_csr(2000, 400000, 0.0004, 9), 0, 1024, 37),
    ("row_partitions", lambda: matgen.rmat_csr(20000, 300000, 9), 4096, 0, 148),
    ("hypersparse_singletons", lambda: matgen.random_csr(30000, 3000000, 0.000004, 13), 0, 32768, 148),
    ("long_row_two_tiles", lambda: (4, 70000, np.array([0, 0, 70000, 70000, 70001], np.uint32),
                                    np.concatenate([np.arange(70000), [3]]).astype(np.uint32),
                                    np.full(70001, 0.001, np.float32)), 0, 0, 148),
    ("saturating", lambda: matgen.random_csr(256, 4096, 0.3, 15), 0, 0, 64),
]


@pytest.mark.parametrize("name,make,rpp,tile_cols,ctas", EMU_CASES, ids=[c[0] for c in EMU_CASES])
def test_plan_walk_matches_oracle(port, name, make, rpp, tile_cols, ctas):
    rows, cols, indptr, indices, data = make()
    scale = 40.0 if name == "saturating" else 1.0          # row sums beyond 2^32 - 1: the clamp must show
    words = port.quantize((data * np.float32(scale)).astype(np.float32))
    x = port.quantize(np.random.default_rng(1).random(cols, dtype=np.float32) * np.float32(scale))
    fmt = capi.Format(rows, cols, indptr, indices, words, rpp, tile_cols)
    want = port.spmv_q824(indptr, indices, words, x)
    assert np.array_equal(fmt.emulate_fixed(ctas, x), want)
    if name == "saturating":
        assert (want == 0xFFFFFFFF).any()


def test_plan_walk_bench_matrix(port):
    """the bench line's own matrix (C2 stand-in, 13.6 M non-zeros) on the bench line's own plan (148 CTAs, three
    x tiles, the mixed-slice cost rule): the host walk reproduces the oracle bit for bit, saturated row included"""
    import bench
    bench.WORKLOAD = "c2"
    r2, c2, ip2, indices, data, x = bench.workload(0)
    words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    fmt = capi.Format(r2, c2, ip2, indices, words)
    assert fmt.stats()["n_col_tiles"] == 3
    want = port.spmv_q824(ip2, indices, words, xw)
    assert np.array_equal(fmt.emulate_fixed(148, xw), want)


# ------------------------------------------------------------------------------------------
# narrow layout (hypersparse matrices): chosen automatically, and forced either way on the same inputs
# ------------------------------------------------------------------------------------------
def test_layout_choice():
    q = hsoracle.Port().quantize
    r, c, ip, ix, d = matgen.random_csr(30000, 3000000, 0.000004, 13)      # ~12 entries per row over 92 tiles: singletons
    assert capi.Format(r, c, ip, ix, q(d), 0, 32768).stats()["layout"] == 1
    r, c, ip, ix, d = matgen.rmat_csr(6000, 150000, 5)                      # 25 per row in one tile
    assert capi.Format(r, c, ip, ix, q(d)).stats()["layout"] == 0


NARROW_CASES = [
    ("singletons", lambda: matgen.random_csr(30000, 3000000, 0.000004, 13), 0, 32768, 148),
    ("rmat_forced", lambda: matgen.rmat_csr(6000, 150000, 5), 0, 0, 148),            # streams up to 32, hub rows split
    ("dense_forced", lambda: matgen.dense_csr(128, 128), 0, 0, 37),
    ("row_partitions_forced", lambda: matgen.rmat_csr(20000, 300000, 9), 4096, 8192, 148),
    ("one_long_row_forced", lambda: (4, 70000, np.array([0, 0, 70000, 70000, 70001], np.uint32),
                                     np.concatenate([np.arange(70000), [3]]).astype(np.uint32),
                                     np.full(70001, 0.001, np.float32)), 0, 0, 148),
    ("empty_rows_and_tiles", lambda: matgen.random_csr(500, 200000, 0.00002, 21), 0, 8192, 5),
]


@pytest.mark.parametrize("narrow", [1, 0])
@pytest.mark.parametrize("name,make,rpp,tile_cols,ctas", NARROW_CASES, ids=[c[0] for c in NARROW_CASES])
def test_both_layouts_round_trip_and_walk(port, monkeypatch, name, make, rpp, tile_cols, ctas, narrow):
    monkeypatch.setenv("HSB_NARROW", str(narrow))
    rows, cols, indptr, indices, data = make()
    words = port.quantize(data)
    x = port.quantize(np.random.default_rng(2).random(cols, dtype=np.float32))
    fmt = capi.Format(rows, cols, indptr, indices, words, rpp, tile_cols)
    st = fmt.stats()
    assert st["layout"] == narrow
    ip, ix, vv = fmt.expand()
    assert np.array_equal(ip, indptr)
    a, b = _canon(rows, indptr, indices, words), _canon(rows, ip, ix, vv)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the plan hands out every unit exactly once, and (narrow) only whole slices
    rec = fmt.plan(ctas)
    total = sum(r[-1][1] for r in _steps_per_tile(fmt, rec).values())
    assert total * (32 if narrow else 128) == st["n_elems"]
    assert np.array_equal(fmt.emulate_fixed(ctas, x), port.spmv_q824(indptr, indices, words, x))


def test_binding_refuses_float64_words():
    """value / vector arguments are 32-bit words: float64 (scipy's default dtype) must be refused, not truncated;
    wide integer index arrays are narrowed after a range check"""
    with pytest.raises(TypeError):
        capi._words(np.array([0.5, 1.5]))
    assert capi._words(np.array([1, 2, 3], np.int64)).dtype == np.uint32
    with pytest.raises(ValueError):
        capi._words(np.array([1, -2], np.int64))
    with pytest.raises(ValueError):
        capi._words(np.array([1 << 33], np.int64))
    a = np.array([0.5], np.float32)
    assert capi._words(a).view(np.float32)[0] == np.float32(0.5)


def test_traffic_evidence_is_keyed_to_the_current_kernel_sources():
    """profiles/traffic.json (the ncu DRAM bytes bench.py quotes as roofline.traffic) carries the hash of the kernel and
    format sources it was captured on; bench.py drops the figure when the hash differs. This keeps the committed
    evidence and the committed sources in step: after a kernel edit, re-run tools/final_capture.sh."""
    import json
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = json.load(open(os.path.join(here, "profiles", "traffic.json")))
    assert t["source_sha256"] == capi.source_hash()
    assert t["dram_bytes_per_launch"] == t["dram_read"] + t["dram_write"]
