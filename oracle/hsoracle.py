"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy bindings for the oracle.

Two things live here:

* ``Port``  -- ``oracle/libhsoracle.so``: the plain-C restatement of the reference hot path
  (``oracle/hsoracle.c``; every function cites the reference file:line it follows).
* ``Ref``   -- ``oracle/_ref/libref_<impl>.so``: the UNMODIFIED reference C simulation
  (``/root/reference/spmv_csim/csim.cpp`` and what it includes) compiled against the clean-room
  HLS shim by ``oracle/Makefile``. ``Ref2`` is the reference formatter at pack_size 2, the
  shape the reference's own golden vectors (``unit_tests/test_io.cpp``) use.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module. The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MARKER = 0xFFFFFFFF
VAL_INT, VAL_FLOAT_BITS, VAL_Q824 = 0, 1, 2
IMPLS = ("fixed", "float_pob", "float_stall")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C port (always) and, when /root/reference is present, oracle/_ref."""
    args = ["make", "-s", "-C", HERE, "all"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Port:
    """oracle/libhsoracle.so"""

    def __init__(self):
        path = os.path.join(HERE, "libhsoracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.hso_quantize_q824.argtypes = [f32p, C.c_size_t, u32p]
        L.hso_q824_mac_chain.argtypes = [u32p, u32p, C.c_size_t]
        L.hso_q824_mac_chain.restype = C.c_uint32
        L.hso_spmv_q824_csr.argtypes = [C.c_uint32, u32p, u32p, u32p, u32p, u32p]
        L.hso_spmv_f32_csr.argtypes = [C.c_uint32, u32p, u32p, f32p, f32p, f32p]
        L.hso_spmv_f64_csr.argtypes = [C.c_uint32, u32p, u32p, f32p, f32p, f64p, f64p]
        L.hso_time_spmv_f32_csr.argtypes = [C.c_uint32, u32p, u32p, f32p, f32p, f32p, C.c_int]
        L.hso_time_spmv_f32_csr.restype = C.c_double
        L.hso_time_spmv_q824_csr.argtypes = [C.c_uint32, u32p, u32p, u32p, u32p, u32p, C.c_int]
        L.hso_time_spmv_q824_csr.restype = C.c_double
        L.hso_round_dims.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
        L.hso_csr2cpsr.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.hso_csr2cpsr.restype = C.c_void_p
        L.hso_cpsr_dims.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        for n in ("hso_cpsr_block_len", "hso_cpsr_block_slots"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
            getattr(L, n).restype = C.c_uint32
        L.hso_cpsr_block_get.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hso_cpsr_free.argtypes = [C.c_void_p]
        L.hso_channel_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.hso_channel_image.restype = C.c_size_t
        L.hso_top_wrapper.argtypes = [C.POINTER(C.c_void_p), u32p, u32p, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.hso_top_wrapper.restype = C.c_int

    # ---- arithmetic -------------------------------------------------------------------
    def quantize(self, v):
        v = _f32(v).ravel()
        out = np.empty(v.size, np.uint32)
        self.L.hso_quantize_q824(v, v.size, out)
        return out

    def mac_chain(self, a, b):
        a, b = _u32(a), _u32(b)
        return int(self.L.hso_q824_mac_chain(a, b, a.size))

    # ---- SpMV on CSR ------------------------------------------------------------------
    def spmv_q824(self, indptr, indices, val, x):
        indptr, indices, val, x = _u32(indptr), _u32(indices), _u32(val), _u32(x)
        y = np.empty(indptr.size - 1, np.uint32)
        self.L.hso_spmv_q824_csr(y.size, indptr, indices, val, x, y)
        return y

    def spmv_f32(self, indptr, indices, val, x):
        indptr, indices, val, x = _u32(indptr), _u32(indices), _f32(val), _f32(x)
        y = np.empty(indptr.size - 1, np.float32)
        self.L.hso_spmv_f32_csr(y.size, indptr, indices, val, x, y)
        return y

    def spmv_f64(self, indptr, indices, val, x):
        indptr, indices, val, x = _u32(indptr), _u32(indices), _f32(val), _f32(x)
        y = np.empty(indptr.size - 1, np.float64)
        sa = np.empty(indptr.size - 1, np.float64)
        self.L.hso_spmv_f64_csr(y.size, indptr, indices, val, x, y, sa)
        return y, sa

    def time_spmv_f32(self, indptr, indices, val, x, runs):
        indptr, indices, val, x = _u32(indptr), _u32(indices), _f32(val), _f32(x)
        y = np.empty(indptr.size - 1, np.float32)
        return float(self.L.hso_time_spmv_f32_csr(y.size, indptr, indices, val, x, y, runs))

    def time_spmv_q824(self, indptr, indices, val, x, runs):
        indptr, indices, val, x = _u32(indptr), _u32(indices), _u32(val), _u32(x)
        y = np.empty(indptr.size - 1, np.uint32)
        return float(self.L.hso_time_spmv_q824_csr(y.size, indptr, indices, val, x, y, runs))

    # ---- formatting -------------------------------------------------------------------
    def round_dims(self, rows, cols, row_div, col_div):
        r, c = C.c_uint32(rows), C.c_uint32(cols)
        self.L.hso_round_dims(C.byref(r), C.byref(c), row_div, col_div)
        return r.value, c.value

    def csr2cpsr(self, rows, cols, indptr, indices, val_words, pack_size, ob, vb, channels, skip, val_kind):
        """rows/cols must already be rounded; indptr padded accordingly. Returns a PortCPSR."""
        indptr, indices, val_words = _u32(indptr), _u32(indices), _u32(val_words)
        assert indptr.size == rows + 1
        h = self.L.hso_csr2cpsr(rows, cols, indptr, indices, val_words, pack_size, ob, vb, channels,
                                int(bool(skip)), val_kind)
        if not h:
            raise ValueError("hso_csr2cpsr: dimensions not rounded (sw/data_formatter.h:475-490)")
        return PortCPSR(self, h, pack_size, channels)

    def top_wrapper(self, images, x, y, impl, interleave, ob, vb, row_part_id, part_len, ncp, nparts, num_cols):
        arr = (C.c_void_p * 16)(*[im.ctypes.data for im in images])
        rc = self.L.hso_top_wrapper(arr, _u32(x), y, impl, interleave, ob, vb, row_part_id, part_len, ncp,
                                    nparts, num_cols)
        if rc:
            raise RuntimeError("hso_top_wrapper: malformed channel image (%d)" % rc)


class PortCPSR:
    def __init__(self, port, h, P, Cn):
        self.port, self.h, self.P, self.C = port, h, P, Cn
        d = (C.c_uint32 * 4)()
        port.L.hso_cpsr_dims(h, d)
        self.rows, self.cols, self.n_row_parts, self.n_col_parts = list(d)

    def block(self, j, i, c):
        """-> (idx[n,P], val[n,P], indptr[slots+1,P])"""
        L = self.port.L
        n = L.hso_cpsr_block_len(self.h, j, i, c)
        s = L.hso_cpsr_block_slots(self.h, j, i, c)
        idx = np.zeros((n, self.P), np.uint32)
        val = np.zeros((n, self.P), np.uint32)
        ptr = np.zeros((s + 1, self.P), np.uint32)
        L.hso_cpsr_block_get(self.h, j, i, c, idx.ctypes.data, val.ctypes.data, ptr.ctypes.data)
        return idx, val, ptr

    def channel_images(self, interleave=1):
        """16 arrays of shape (n_packets, 16) uint32: the per-HBM-channel images (host.cpp:163-231)."""
        L = self.port.L
        out = []
        for pc in range(self.C // interleave):
            n = L.hso_channel_image(self.h, pc, interleave, None)
            im = np.zeros((max(n, 1), 16), np.uint32)
            L.hso_channel_image(self.h, pc, interleave, im.ctypes.data)
            out.append(im[:n] if n else im[:0])
        return out

    def __del__(self):
        try:
            self.port.L.hso_cpsr_free(self.h)
        except Exception:
            pass


def ref_available(impl="fixed"):
    return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % impl))


class Ref:
    """oracle/_ref/libref_<impl>.so -- the reference's own code."""

    def __init__(self, impl="fixed"):
        path = os.path.join(HERE, "_ref", "libref_%s.so" % impl)
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        cfg = (C.c_uint * 8)()
        L.ref_config(cfg)
        (self.PACK_SIZE, self.NUM_HBM_CHANNELS, self.INTERLEAVE_FACTOR, self.LOGICAL_OB_SIZE,
         self.LOGICAL_VB_SIZE, self.PKT_BYTES, self.VAL_BYTES, self.impl_id) = list(cfg)
        self.impl = impl
        L.ref_val_from_float.argtypes = [f32p, C.c_size_t, u32p]
        L.ref_mac_chain.argtypes = [u32p, u32p, C.c_size_t]
        L.ref_mac_chain.restype = C.c_uint32
        L.ref_csr2cpsr.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f32p, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.ref_csr2cpsr.restype = C.c_void_p
        L.ref_cpsr_dims.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.ref_cpsr_len.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ref_cpsr_len.restype = C.c_size_t
        L.ref_cpsr_get.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_cpsr_free.argtypes = [C.c_void_p]
        L.ref_top_wrapper.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p] + [C.c_uint] * 5
        L.ref_compute_ref.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f32p, f32p, f32p]
        L.ref_time_compute_ref.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f32p, f32p, f32p, C.c_int]
        L.ref_time_compute_ref.restype = C.c_double
        L.ref_time_compute_ref_mt.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f32p, f32p, f32p, C.c_int, C.c_int]
        L.ref_time_compute_ref_mt.restype = C.c_double
        L.ref_test_harness.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f32p, C.c_int, C.c_uint]
        L.ref_selftest.restype = C.c_int

    def val_from_float(self, v):
        v = _f32(v).ravel()
        out = np.empty(v.size, np.uint32)
        self.L.ref_val_from_float(v, v.size, out)
        return out

    def mac_chain(self, a, b):
        a, b = _u32(a), _u32(b)
        return int(self.L.ref_mac_chain(a, b, a.size))

    def csr2cpsr(self, rows, cols, indptr, indices, data_f32, skip, ob=None, vb=None, channels=None):
        ob = ob or self.LOGICAL_OB_SIZE
        vb = vb or self.LOGICAL_VB_SIZE
        channels = channels or self.NUM_HBM_CHANNELS * self.INTERLEAVE_FACTOR
        h = self.L.ref_csr2cpsr(rows, cols, _u32(indptr), _u32(indices), _f32(data_f32),
                                self.PACK_SIZE * self.NUM_HBM_CHANNELS * self.INTERLEAVE_FACTOR,
                                self.PACK_SIZE, ob, vb, channels, int(bool(skip)))
        return RefCPSR(self, h, channels)

    def top_wrapper(self, images, x_words, y_words, row_part_id, part_len, ncp, nparts, num_cols):
        arr = (C.c_void_p * 16)(*[im.ctypes.data for im in images])
        assert x_words.dtype == np.uint32 and y_words.dtype == np.uint32
        self.L.ref_top_wrapper(arr, x_words.ctypes.data, y_words.ctypes.data, row_part_id, part_len, ncp,
                               nparts, num_cols)

    def compute_ref(self, rows, cols, indptr, indices, data, x):
        y = np.empty(rows, np.float32)
        self.L.ref_compute_ref(rows, cols, _u32(indptr), _u32(indices), _f32(data), _f32(x), y)
        return y

    def time_compute_ref(self, rows, cols, indptr, indices, data, x, runs):
        y = np.empty(rows, np.float32)
        return float(self.L.ref_time_compute_ref(rows, cols, _u32(indptr), _u32(indices), _f32(data),
                                                 _f32(x), y, runs))

    def time_compute_ref_mt(self, rows, cols, indptr, indices, data, x, runs, threads, y=None):
        """compute_ref on `threads` nnz-balanced row blocks at once (every host core): seconds per SpMV"""
        if y is None:
            y = np.empty(rows, np.float32)
        return float(self.L.ref_time_compute_ref_mt(rows, cols, _u32(indptr), _u32(indices), _f32(data),
                                                    _f32(x), y, runs, threads))

    def test_harness(self, rows, cols, indptr, indices, data, skip, seed=1):
        return bool(self.L.ref_test_harness(rows, cols, _u32(indptr), _u32(indices), _f32(data),
                                            int(bool(skip)), seed))

    def selftest(self):
        return bool(self.L.ref_selftest())


class RefCPSR:
    def __init__(self, ref, h, channels):
        self.ref, self.h, self.C, self.P = ref, h, channels, ref.PACK_SIZE
        d = (C.c_uint32 * 4)()
        ref.L.ref_cpsr_dims(h, d)
        self.rows, self.cols, self.n_row_parts, self.n_col_parts = list(d)

    def block(self, j, i, c):
        """-> (idx[n,P], val[n,P], lane_lengths[P])"""
        n = self.ref.L.ref_cpsr_len(self.h, j, i, c)
        idx = np.zeros((n, self.P), np.uint32)
        val = np.zeros((n, self.P), np.uint32)
        lens = np.zeros(self.P, np.uint32)
        self.ref.L.ref_cpsr_get(self.h, j, i, c, idx.ctypes.data, val.ctypes.data, lens.ctypes.data)
        return idx, val, lens

    def __del__(self):
        try:
            self.ref.L.ref_cpsr_free(self.h)
        except Exception:
            pass


class Ref2:
    """oracle/_ref/libref_fmt2.so -- reference formatter at pack_size 2 (test_io.cpp shapes)."""

    P = 2

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_fmt2.so")
        if not os.path.exists(path):
            build()
        self.L = C.CDLL(path)
        for n in ("ref2_csr2cpsr_i32", "ref2_csr2cpsr_f32"):
            getattr(self.L, n).restype = C.c_void_p
        for n in ("ref2_get_i32", "ref2_get_f32", "ref2_pack_rows_f32"):
            getattr(self.L, n).restype = C.c_size_t

    def csr2cpsr(self, rows, cols, indptr, indices, data, ob, vb, nch, skip, kind="i32"):
        indptr, indices = _u32(indptr), _u32(indices)
        data = np.ascontiguousarray(data, dtype=np.int32 if kind == "i32" else np.float32)
        fn = getattr(self.L, "ref2_csr2cpsr_" + kind)
        h = C.c_void_p(fn(C.c_uint32(rows), C.c_uint32(cols), indptr.ctypes.data_as(C.c_void_p),
                          indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                          C.c_uint32(ob), C.c_uint32(vb), C.c_uint32(nch), C.c_int(int(skip))))
        return h

    def block(self, h, j, i, c, kind="i32"):
        get = getattr(self.L, "ref2_get_" + kind)
        nptr = C.c_size_t()
        n = get(h, C.c_uint32(j), C.c_uint32(i), C.c_uint32(c), None, None, None, C.byref(nptr))
        idx = np.zeros((n, 2), np.uint32)
        val = np.zeros((n, 2), np.int32 if kind == "i32" else np.float32)
        ptr = np.zeros((nptr.value, 2), np.uint32)
        get(h, C.c_uint32(j), C.c_uint32(i), C.c_uint32(c), idx.ctypes.data_as(C.c_void_p),
            val.ctypes.data_as(C.c_void_p), ptr.ctypes.data_as(C.c_void_p), C.byref(nptr))
        return idx, val, ptr

    def pack_rows(self, indptr, indices, data, nch, c):
        indptr, indices, data = _u32(indptr), _u32(indices), _f32(data)
        rows = indptr.size - 1
        idx = np.zeros((indices.size + 1, 2), np.uint32)
        val = np.zeros((indices.size + 1, 2), np.float32)
        ptr = np.zeros((rows + 2, 2), np.uint32)
        nptr = C.c_size_t()
        n = self.L.ref2_pack_rows_f32(C.c_uint32(rows), indptr.ctypes.data_as(C.c_void_p),
                                      indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                                      C.c_uint32(nch), C.c_uint32(c), idx.ctypes.data_as(C.c_void_p),
                                      val.ctypes.data_as(C.c_void_p), ptr.ctypes.data_as(C.c_void_p), C.byref(nptr))
        return idx[:n], val[:n], ptr[:nptr.value]

    def csr_to_dds(self, rows, cols, indptr, indices, data, cols_per_part, part):
        indptr, indices, data = _u32(indptr), _u32(indices), _f32(data)
        od = np.zeros(indices.size, np.float32)
        oi = np.zeros(indices.size, np.uint32)
        op = np.zeros(rows + 1, np.uint32)
        nnz = C.c_uint32()
        self.L.ref2_csr_to_dds_f32(C.c_uint32(rows), C.c_uint32(cols), indptr.ctypes.data_as(C.c_void_p),
                                   indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                                   C.c_uint32(cols_per_part), C.c_uint32(part), od.ctypes.data_as(C.c_void_p),
                                   oi.ctypes.data_as(C.c_void_p), op.ctypes.data_as(C.c_void_p), C.byref(nnz))
        return od[:nnz.value], oi[:nnz.value], op

    def round_dims(self, rows, cols, rd, cd):
        r, c = C.c_uint32(rows), C.c_uint32(cols)
        self.L.ref2_round_dims(C.byref(r), C.byref(c), C.c_uint32(rd), C.c_uint32(cd))
        return r.value, c.value


def axpb_q824(alpha_word, y_words, beta_word):
    """alpha (*) y (+) beta in ap_ufixed<32,8,AP_RND,AP_SAT> on raw words: the PE's rounded saturating product
    (spmv/libfpga/pe.h:64: VAL_T incr = mat * vec) followed by its saturating add (pe.h:72) -- the checker for
    hsb_axpb_to_vector / hsb_axpb_to_peers / hsb_iterate (fixed point)."""
    q = (np.uint64(alpha_word) * np.asarray(y_words).astype(np.uint64) + np.uint64(1 << 23)) >> np.uint64(24)
    q = np.minimum(q, np.uint64(0xFFFFFFFF)) + np.uint64(beta_word)
    return np.minimum(q, np.uint64(0xFFFFFFFF)).astype(np.uint32)
