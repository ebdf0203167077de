"""ctypes binding of the C ABI declared in include/hisparse_b200.h.

This is glue for tests and bench.py (the reference's host side is C++; the C++ host mirror lives
in hisparse_b200/host/). There is NO fallback: if libhisparse_b200.so is missing or the CUDA
device cannot be used, the calls raise.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("HSB_LIB") or os.path.join(HERE, "libhisparse_b200.so")   # HSB_LIB: tuning builds
HEADER = os.path.join(ROOT, "include", "hisparse_b200.h")

IMPL_FIXED, IMPL_FLOAT_POB, IMPL_FLOAT_STALL = 0, 1, 2
IMPL_BY_NAME = {"fixed": 0, "float_pob": 1, "float_stall": 2}
NUM_HBM_CHANNELS, PACK_SIZE = 16, 8
PEER_BLOB_BYTES = 192


class HsbError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("pack_size", "num_hbm_channels", "interleave_factor",
                                          "logical_ob_size", "logical_vb_size")]


class Stats(C.Structure):
    _fields_ = [("nnz", C.c_uint64), ("rows", C.c_uint32), ("cols", C.c_uint32), ("n_row_parts", C.c_uint32),
                ("n_col_tiles", C.c_uint32), ("tile_cols", C.c_uint32), ("n_slices", C.c_uint64),
                ("n_streams", C.c_uint64), ("n_elems", C.c_uint64), ("format_bytes", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("sm_count", C.c_uint32), ("grid", C.c_uint32),
                ("replicas", C.c_uint32), ("layout", C.c_uint32), ("preprocess_seconds", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DeviceCsrStruct(C.Structure):
    _fields_ = [("rows", C.c_uint32), ("cols", C.c_uint32), ("nnz", C.c_uint64), ("d_indptr", C.c_void_p),
                ("d_indices", C.c_void_p), ("d_vals", C.c_void_p), ("device", C.c_int)]


def build(force=False):
    """Compile hisparse_b200/libhisparse_b200.so for sm_100a with nvcc (in-tree)."""
    args = ["make", "-s", "-C", os.path.join(HERE, "csrc")]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def source_hash():
    """SHA-256 over the CODE that decides what a launch reads from HBM -- the kernels, the format and the planner
    (csrc/spmv_kernels.*, tile_format.*, gpu_format.cu), comments and white space stripped: identifies a build across
    recompilations, comment edits and changes to the host shim; profiles/traffic.json is keyed by it."""
    import hashlib
    h = hashlib.sha256()
    for name in ("spmv_kernels.cu", "spmv_kernels.cuh", "tile_format.cpp", "tile_format.h", "gpu_format.cu"):
        text = open(os.path.join(HERE, "csrc", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        h.update(name.encode())
        h.update("".join(text.split()).encode())
    return h.hexdigest()


def declared_symbols():
    """Every function name include/hisparse_b200.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsbError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    L.hsb_version.restype = C.c_char_p
    L.hsb_last_error.restype = C.c_char_p
    L.hsb_get_config.argtypes = [C.c_int, C.POINTER(Config)]
    L.hsb_create.argtypes = [C.c_int, C.c_int]
    L.hsb_create.restype = vp
    L.hsb_destroy.argtypes = [vp]
    L.hsb_destroy.restype = None
    L.hsb_host_alloc.argtypes = [sz]
    L.hsb_host_alloc.restype = vp
    L.hsb_host_free.argtypes = [vp]
    L.hsb_host_free.restype = None
    L.hsb_upload_matrix_cpsr.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    L.hsb_upload_matrix_csr.argtypes = [vp, u32, u32, vp, vp, vp, u32]
    L.hsb_upload_matrix_csr_gpu.argtypes = [vp, u32, u32, vp, vp, vp, u32]
    L.hsb_upload_matrix_csr_device.argtypes = [vp, u32, u32, C.c_uint64, vp, vp, vp, u32]
    L.hsb_upload_vector.argtypes = [vp, vp, C.c_uint]
    L.hsb_spmv_row_partition.argtypes = [vp] + [C.c_uint] * 5
    L.hsb_spmv.argtypes = [vp]
    L.hsb_sync.argtypes = [vp]
    L.hsb_download_result.argtypes = [vp, vp, C.c_uint]
    L.hsb_download_result_async.argtypes = [vp, vp, C.c_uint]
    L.hsb_top_wrapper.argtypes = [C.c_int, C.POINTER(vp), vp, vp] + [C.c_uint] * 5
    L.hsb_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.hsb_set_replicas.argtypes = [vp, C.c_int]
    L.hsb_time_spmv.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.hsb_time_e2e.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.c_uint, C.c_uint, C.c_int, C.c_int,
                               C.POINTER(C.c_double)]
    for n in ("hsb_device_x", "hsb_device_y", "hsb_stream"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = vp
    L.hsb_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.hsb_axpb_to_vector.argtypes = [vp, u32, u32, u32]
    L.hsb_device_x_next.argtypes = [vp]
    L.hsb_device_x_next.restype = vp
    L.hsb_vector_commit.argtypes = [vp]
    L.hsb_iterate.argtypes = [vp, C.c_int, u32, u32]
    L.hsb_iterate_peers.argtypes = [vp, C.c_int, u32, u32, u32]
    L.hsb_peer_export.argtypes = [vp, vp]
    L.hsb_peer_connect.argtypes = [vp, C.c_int, C.c_int, vp]
    L.hsb_axpb_to_peers.argtypes = [vp, u32, u32, u32]
    L.hsb_gather_export.argtypes = [vp, u32, C.c_int, vp]
    L.hsb_gather_connect.argtypes = [vp, C.c_int, C.c_int, u32, vp]
    L.hsb_gather_wait.argtypes = [vp]
    L.hsb_device_y_gathered.argtypes = [vp]
    L.hsb_device_y_gathered.restype = vp
    L.hsb_download_gathered.argtypes = [vp, vp, u32]
    L.hsb_device_numa_node.argtypes = [C.c_int]
    L.hsb_device_l2_bytes.argtypes = [C.c_int]
    L.hsb_device_l2_bytes.restype = sz
    L.hsb_debug_trace.argtypes = [vp, vp, sz]
    L.hsb_debug_profiler.argtypes = [vp, C.c_int]
    L.hsb_debug_timeline.argtypes = [vp, vp, sz]
    L.hsb_debug_plan.argtypes = [vp, vp, vp, sz]
    L.hsb_format_build.argtypes = [u32, u32, vp, vp, vp, u32, u32]
    L.hsb_format_build.restype = vp
    L.hsb_format_from_context.argtypes = [vp]
    L.hsb_format_from_context.restype = vp
    L.hsb_format_stats.argtypes = [vp, C.POINTER(Stats)]
    L.hsb_format_expand.argtypes = [vp, vp, vp, vp]
    L.hsb_format_plan.argtypes = [vp, u32, vp, sz]
    L.hsb_format_plan.restype = C.c_longlong
    L.hsb_format_free.argtypes = [vp]
    L.hsb_format_free.restype = None
    L.hsb_cpsr_to_csr.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(sz), C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                  vp, vp, vp, sz, C.POINTER(sz)]
    L.hsb_synth_powerlaw_csr_device.argtypes = [C.c_int, u32, u32, C.c_uint64, C.c_double, C.c_double, u32, u32,
                                                C.c_double, C.c_uint64, C.c_int, C.c_float, C.POINTER(DeviceCsrStruct)]
    L.hsb_device_csr_download.argtypes = [C.POINTER(DeviceCsrStruct), vp, vp, vp]
    L.hsb_device_csr_free.argtypes = [C.POINTER(DeviceCsrStruct)]
    L.hsb_device_csr_free.restype = None
    L.hsb_synth_last_error.restype = C.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise HsbError("hisparse_b200 error %d: %s" % (rc, lib().hsb_last_error().decode()))


def _words(a):
    """32-bit words for the C ABI: float32 and (u)int32 arrays are reinterpreted, wider INTEGER arrays (index
    arrays such as scipy's int64 indptr) are range-checked and narrowed. Anything else -- float64 above all,
    scipy's default value dtype -- is refused: converting it here would silently truncate 0.5 to 0."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return a
    if a.dtype in (np.float32, np.int32):
        return a.view(np.uint32)
    if np.issubdtype(a.dtype, np.integer):
        if a.size and (int(a.min()) < 0 or int(a.max()) > 0xFFFFFFFF):
            raise ValueError("index array does not fit 32 bits")
        return a.astype(np.uint32)
    raise TypeError("expected 32-bit words (float32 values, or raw uint32 Q8.24 words), got %s: convert explicitly "
                    "(x.astype(np.float32), or matgen.quantize_q824 for the fixed-point path)" % a.dtype)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def get_config(impl):
    cfg = Config()
    _check(lib().hsb_get_config(impl, C.byref(cfg)))
    return cfg


def device_count():
    return lib().hsb_device_count()


def device_l2_bytes(device=0):
    return int(lib().hsb_device_l2_bytes(device))


def _images_args(images):
    imgs = [np.ascontiguousarray(im, dtype=np.uint32) for im in images]
    arr = (C.c_void_p * 16)(*[im.ctypes.data for im in imgs])
    lens = (C.c_size_t * 16)(*[im.size // 16 for im in imgs])
    return imgs, arr, lens


class PinnedArray:
    """numpy view over page-locked host memory from hsb_host_alloc."""

    def __init__(self, n_words, dtype=np.uint32):
        self._p = lib().hsb_host_alloc(n_words * 4)
        if not self._p:
            raise HsbError("hsb_host_alloc failed")
        buf = (C.c_uint32 * n_words).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=n_words)

    def __del__(self):
        try:
            self.array = None
            lib().hsb_host_free(self._p)
        except Exception:
            pass


class DeviceCsr:
    """A CSR that lives in device memory (hsb_device_csr): here always a synthetic power-law shard
    generated on the GPU (BASELINE config C5)."""

    def __init__(self, st):
        self.st = st

    @classmethod
    def powerlaw(cls, device, rows, cols, first_global_row=0, mean_degree=20.0, alpha=2.1, max_degree=10 ** 6,
                 band_half_width=1 << 20, band_fraction=0.8, seed=0xC0FFEE05, q824=False, value_scale=1.0):
        st = DeviceCsrStruct()
        rc = lib().hsb_synth_powerlaw_csr_device(device, rows, cols, first_global_row, mean_degree, alpha, max_degree,
                                                 band_half_width, band_fraction, seed, 1 if q824 else 0, value_scale,
                                                 C.byref(st))
        if rc != 0:
            raise HsbError("hsb_synth_powerlaw_csr_device error %d: %s" % (rc, lib().hsb_synth_last_error().decode()))
        return cls(st)

    rows = property(lambda self: self.st.rows)
    cols = property(lambda self: self.st.cols)
    nnz = property(lambda self: self.st.nnz)

    def download(self):
        indptr = np.empty(self.rows + 1, np.uint32)
        indices = np.empty(max(self.nnz, 1), np.uint32)
        vals = np.empty(max(self.nnz, 1), np.uint32)
        rc = lib().hsb_device_csr_download(C.byref(self.st), _ptr(indptr), _ptr(indices), _ptr(vals))
        if rc != 0:
            raise HsbError("hsb_device_csr_download: " + lib().hsb_synth_last_error().decode())
        return indptr, indices[:self.nnz], vals[:self.nnz]

    def download_rows(self, r0, r1):
        """CSR of rows [r0, r1) only (indptr rebased to 0): a window of a shard too big to bring to the host whole"""
        assert 0 <= r0 <= r1 <= self.rows
        win = DeviceCsrStruct()
        win.rows, win.cols, win.device = r1 - r0, self.cols, self.st.device
        win.d_indptr = self.st.d_indptr + 4 * r0
        indptr = np.empty(r1 - r0 + 1, np.uint32)
        if lib().hsb_device_csr_download(C.byref(win), _ptr(indptr), None, None) != 0:
            raise HsbError("hsb_device_csr_download: " + lib().hsb_synth_last_error().decode())
        e0, e1 = int(indptr[0]), int(indptr[-1])
        win.nnz = e1 - e0
        win.d_indices, win.d_vals = self.st.d_indices + 4 * e0, self.st.d_vals + 4 * e0
        indices = np.empty(max(win.nnz, 1), np.uint32)
        vals = np.empty(max(win.nnz, 1), np.uint32)
        if lib().hsb_device_csr_download(C.byref(win), None, _ptr(indices), _ptr(vals)) != 0:
            raise HsbError("hsb_device_csr_download: " + lib().hsb_synth_last_error().decode())
        return (indptr - np.uint32(e0)).astype(np.uint32), indices[:win.nnz], vals[:win.nnz]

    def free(self):
        if self.st.d_indptr:
            lib().hsb_device_csr_free(C.byref(self.st))

    __del__ = free


class Context:
    """One GPU, one implementation (fixed / float_pob / float_stall): mirrors the reference's
    cl_runtime struct (sw/host.cpp:120-128)."""

    def __init__(self, device=0, impl=IMPL_FIXED):
        if isinstance(impl, str):
            impl = IMPL_BY_NAME[impl]
        self.impl = impl
        self.h = lib().hsb_create(device, impl)
        if not self.h:
            raise HsbError("hsb_create failed: " + lib().hsb_last_error().decode())
        self.rows = self.cols = 0

    def close(self):
        if self.h:
            lib().hsb_destroy(self.h)
            self.h = None

    __del__ = close

    def upload_matrix_csr(self, rows, cols, indptr, indices, vals, rows_per_partition=0, on_gpu=False):
        """on_gpu=True: the tile-stream format is built by device kernels (hsb_upload_matrix_csr_gpu)"""
        indptr, indices, vals = _words(indptr), _words(indices), _words(vals)
        assert indptr.size == rows + 1
        fn = lib().hsb_upload_matrix_csr_gpu if on_gpu else lib().hsb_upload_matrix_csr
        _check(fn(self.h, rows, cols, _ptr(indptr), _ptr(indices), _ptr(vals), rows_per_partition))
        self.rows, self.cols = rows, cols

    def upload_matrix_csr_device(self, dcsr, rows_per_partition=0):
        """a CSR already in device memory (DeviceCsr): formatted on the GPU, never visits the host"""
        st = dcsr.st
        _check(lib().hsb_upload_matrix_csr_device(self.h, st.rows, st.cols, st.nnz, st.d_indptr, st.d_indices,
                                                  st.d_vals, rows_per_partition))
        self.rows, self.cols = st.rows, st.cols

    def upload_matrix_cpsr(self, images, n_row_parts, n_col_parts, rows, cols):
        imgs, arr, lens = _images_args(images)
        _check(lib().hsb_upload_matrix_cpsr(self.h, arr, lens, n_row_parts, n_col_parts, rows, cols))
        self.rows, self.cols = rows, cols

    def upload_vector(self, x):
        x = _words(x)
        _check(lib().hsb_upload_vector(self.h, _ptr(x), x.size))

    def spmv_row_partition(self, row_part_id, part_len, ncp, nparts, num_cols):
        _check(lib().hsb_spmv_row_partition(self.h, row_part_id, part_len, ncp, nparts, num_cols))

    def spmv(self):
        _check(lib().hsb_spmv(self.h))

    def sync(self):
        _check(lib().hsb_sync(self.h))

    def download_result(self, out=None, dtype=np.uint32):
        if out is None:
            out = np.empty(self.rows, dtype)
        _check(lib().hsb_download_result(self.h, _ptr(out), out.size))
        return out

    def download_result_async(self, out):
        _check(lib().hsb_download_result_async(self.h, _ptr(out), out.size))
        return out

    def stats(self):
        s = Stats()
        _check(lib().hsb_get_stats(self.h, C.byref(s)))
        return s.asdict()

    def set_replicas(self, n):
        _check(lib().hsb_set_replicas(self.h, n))

    def time_spmv(self, warmup, steps, kernel=True):
        a, b = C.c_float(), C.c_float()
        _check(lib().hsb_time_spmv(self.h, warmup, steps, C.byref(a), C.byref(b) if kernel else None))
        return a.value, (b.value if kernel else None)

    def time_e2e(self, x_host, y_host, iters, async_download=True):
        """hsb_time_e2e: x_host / y_host are two page-locked arrays each (PinnedArray.array); seconds per SpMV"""
        xs = (C.c_void_p * 2)(*[a.ctypes.data for a in x_host])
        ys = (C.c_void_p * 2)(*[a.ctypes.data for a in y_host])
        sec = C.c_double()
        _check(lib().hsb_time_e2e(self.h, xs, ys, x_host[0].size, y_host[0].size, iters, 1 if async_download else 0,
                                  C.byref(sec)))
        return sec.value

    def axpb_to_vector(self, alpha_word, beta_word, col_offset=0):
        _check(lib().hsb_axpb_to_vector(self.h, int(alpha_word), int(beta_word), col_offset))

    def device_x_next(self):
        return lib().hsb_device_x_next(self.h)

    def vector_commit(self):
        _check(lib().hsb_vector_commit(self.h))

    def peer_export(self):
        blob = np.zeros(PEER_BLOB_BYTES, np.uint8)
        _check(lib().hsb_peer_export(self.h, _ptr(blob)))
        return blob

    def peer_connect(self, world, rank, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        assert blobs.size == world * PEER_BLOB_BYTES
        _check(lib().hsb_peer_connect(self.h, world, rank, _ptr(blobs)))

    def axpb_to_peers(self, alpha_word, beta_word, col_offset):
        _check(lib().hsb_axpb_to_peers(self.h, int(alpha_word), int(beta_word), col_offset))

    def gather_export(self, total_rows, want_buffer=True):
        blob = np.zeros(PEER_BLOB_BYTES, np.uint8)
        _check(lib().hsb_gather_export(self.h, total_rows, 1 if want_buffer else 0, _ptr(blob)))
        self.gather_rows = total_rows
        return blob

    def gather_connect(self, world, rank, row_offset, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        assert blobs.size == world * PEER_BLOB_BYTES
        _check(lib().hsb_gather_connect(self.h, world, rank, row_offset, _ptr(blobs)))

    def gather_wait(self):
        _check(lib().hsb_gather_wait(self.h))

    def device_y_gathered(self):
        return lib().hsb_device_y_gathered(self.h)

    def download_gathered(self, out=None):
        if out is None:
            out = np.empty(self.gather_rows, np.uint32)
        _check(lib().hsb_download_gathered(self.h, _ptr(out), out.size))
        return out

    def iterate(self, iters, alpha_word, beta_word):
        """iters x { x <- alpha (*) A x (+) beta } on the device (PageRank-style power iteration)"""
        _check(lib().hsb_iterate(self.h, iters, int(alpha_word), int(beta_word)))

    def iterate_peers(self, iters, alpha_word, beta_word, col_offset):
        """hsb_iterate_peers: the multi-GPU iteration as one resident kernel per GPU (after peer_connect)"""
        _check(lib().hsb_iterate_peers(self.h, iters, int(alpha_word), int(beta_word), int(col_offset)))

    def set_option(self, name, value):
        _check(lib().hsb_set_option(self.h, name.encode(), int(value)))

    def trace(self, arm=None):
        """arm=True/False switches tracing; arm=None returns the [sm_count, 34] stamps of the last launch"""
        if arm is not None:
            return lib().hsb_debug_trace(self.h, None, 1 if arm else 0)
        n = lib().hsb_debug_trace(self.h, None, 1)
        out = np.zeros(n, np.uint64)
        rc = lib().hsb_debug_trace(self.h, _ptr(out), n)
        if rc < 0:
            _check(rc)
        return out.reshape(self.stats()["sm_count"], -1)

    def timeline(self, arm=None):
        """arm=True/False switches it; arm=None -> (last launch number, [256, 8] globaltimer stamps in ns)"""
        if arm is not None:
            return lib().hsb_debug_timeline(self.h, None, 1 if arm else 0)
        out = np.zeros(256 * 8, np.uint64)
        rc = lib().hsb_debug_timeline(self.h, _ptr(out), out.size)
        if rc < 0:
            _check(rc)
        return rc, out.reshape(256, 8)

    def profiler(self, on):
        _check(lib().hsb_debug_profiler(self.h, 1 if on else 0))

    def plan(self):
        n = self.stats()["sm_count"]
        a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rc = lib().hsb_debug_plan(self.h, _ptr(a), _ptr(b), n)
        if rc < 0:
            _check(rc)
        return a[:rc], b[:rc]

    def device_x(self):
        return lib().hsb_device_x(self.h)

    def device_y(self):
        return lib().hsb_device_y(self.h)

    def stream(self):
        return lib().hsb_stream(self.h)


def top_wrapper(impl, images, x, y, row_part_id, part_len, ncp, nparts, num_cols):
    """spmv_csim/csim.cpp:22-46 with host buffers; y (uint32 words) is written in place."""
    if isinstance(impl, str):
        impl = IMPL_BY_NAME[impl]
    imgs, arr, _ = _images_args(images)
    x = _words(x)
    assert y.dtype == np.uint32 and y.flags.c_contiguous
    _check(lib().hsb_top_wrapper(impl, arr, _ptr(x), _ptr(y), row_part_id, part_len, ncp, nparts, num_cols))


class Format:
    """Host-side view of the tile-stream format (no GPU)."""

    @classmethod
    def from_context(cls, ctx):
        """the matrix resident on the device, downloaded for inspection"""
        self = cls.__new__(cls)
        self.h = lib().hsb_format_from_context(ctx.h)
        if not self.h:
            raise HsbError("hsb_format_from_context: " + lib().hsb_last_error().decode())
        st = self.stats()
        self.rows, self.nnz = st["rows"], st["nnz"]
        return self

    def __init__(self, rows, cols, indptr, indices, vals, rows_per_partition=0, tile_cols=0):
        indptr, indices, vals = _words(indptr), _words(indices), _words(vals)
        self.rows, self.nnz = rows, int(indptr[-1]) if indptr.size else 0
        self._csr = (rows, cols, indptr, indices, vals, rows_per_partition, tile_cols)
        self.h = lib().hsb_format_build(rows, cols, _ptr(indptr), _ptr(indices), _ptr(vals), rows_per_partition,
                                        tile_cols)
        if not self.h:
            raise HsbError("hsb_format_build rejected the matrix")

    def stats(self):
        s = Stats()
        _check(lib().hsb_format_stats(self.h, C.byref(s)))
        return s.asdict()

    def plan(self, ctas):
        """-> [n, 4] uint32 records {cta, tile, first step, end step}: the warp shares of the launch plan"""
        n = lib().hsb_format_plan(self.h, ctas, None, 0)
        if n < 0:
            _check(int(n))
        rec = np.zeros((max(n, 1), 4), np.uint32)
        lib().hsb_format_plan(self.h, ctas, _ptr(rec), n)
        return rec[:n]

    def emulate_fixed(self, ctas, x_words):
        """host walk of the launch plan the way the kernel executes it (test aid) -> y words"""
        x_words = _words(x_words)
        y = np.zeros(self.rows, np.uint32)
        rows, cols, indptr, indices, vals, rpp, tile_cols = self._csr
        rc = _plancheck().hsbt_emulate_fixed(rows, cols, _ptr(indptr), _ptr(indices), _ptr(vals), rpp, tile_cols, ctas,
                                             _ptr(x_words), _ptr(y))
        if rc != 0:
            raise HsbError("the host walk of the launch plan found an inconsistency (%d)" % rc)
        return y

    def expand(self):
        indptr = np.zeros(self.rows + 1, np.uint32)
        indices = np.zeros(max(self.nnz, 1), np.uint32)
        vals = np.zeros(max(self.nnz, 1), np.uint32)
        _check(lib().hsb_format_expand(self.h, _ptr(indptr), _ptr(indices), _ptr(vals)))
        return indptr, indices[:self.nnz], vals[:self.nnz]

    def __del__(self):
        try:
            lib().hsb_format_free(self.h)
        except Exception:
            pass


_plan_lib = None


def _plancheck():
    """test aid built by hisparse_b200/host/Makefile (not part of the product library)"""
    global _plan_lib
    if _plan_lib is None:
        path = os.path.join(HERE, "host", "bin", "libhsb_plancheck.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host"), path], check=True, stdout=subprocess.DEVNULL)
        _plan_lib = C.CDLL(path)
        u32, vp = C.c_uint32, C.c_void_p
        _plan_lib.hsbt_emulate_fixed.argtypes = [u32, u32, vp, vp, vp, u32, u32, u32, vp, vp]
    return _plan_lib


def cpsr_to_csr(impl, images, n_row_parts, n_col_parts, rows, cols):
    if isinstance(impl, str):
        impl = IMPL_BY_NAME[impl]
    imgs, arr, lens = _images_args(images)
    nnz = C.c_size_t()
    _check(lib().hsb_cpsr_to_csr(impl, arr, lens, n_row_parts, n_col_parts, rows, cols, None, None, None, 0,
                                 C.byref(nnz)))
    indptr = np.zeros(rows + 1, np.uint32)
    indices = np.zeros(max(nnz.value, 1), np.uint32)
    vals = np.zeros(max(nnz.value, 1), np.uint32)
    _check(lib().hsb_cpsr_to_csr(impl, arr, lens, n_row_parts, n_col_parts, rows, cols, _ptr(indptr), _ptr(indices),
                                 _ptr(vals), indices.size, C.byref(nnz)))
    return indptr, indices[:nnz.value], vals[:nnz.value]
