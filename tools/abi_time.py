"""Device-resident and isolated-launch SpMV time of one workload for ANY build of the library (HSB_LIB), through
the handful of C-ABI entry points every round's build has -- the tool for bisecting a timing change across
commits (build each commit's csrc/ into hisparse_b200/libhsb_<tag>.so and run this once per library).
    HSB_LIB=... python tools/abi_time.py c1|c2|c3|t95|c4"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hisparse_b200 import matgen  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
bench.WORKLOAD = wl
cache = "/tmp/hsb_%s_cache.npz" % wl
if os.path.exists(cache):
    d = np.load(cache)
    r2, c2, ip2, indices, words, xw = int(d["r2"]), int(d["c2"]), d["ip2"], d["indices"], d["words"], d["xw"]
else:
    r2, c2, ip2, indices, data, x = bench.workload(0)
    if bench.WORKLOADS[wl][1] == "fixed":
        words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    else:
        words, xw = data.view(np.uint32), x.view(np.uint32)
    np.savez(cache, r2=r2, c2=c2, ip2=ip2, indices=indices, words=words, xw=xw)
if os.environ.get("ABI_SHARD"):                          # "world:rank": one nnz-balanced row block of the workload
    from hisparse_b200 import sharding
    world, rank = [int(v) for v in os.environ["ABI_SHARD"].split(":")]
    bounds = sharding.shard_bounds(ip2, world)
    ip2, indices, words = sharding.extract_shard(ip2, indices, words, bounds[rank], bounds[rank + 1])
    r2 = bounds[rank + 1] - bounds[rank]
L = C.CDLL(os.environ.get("HSB_LIB") or os.path.join(ROOT, "hisparse_b200", "libhisparse_b200.so"))
vp, u32 = C.c_void_p, C.c_uint32
L.hsb_create.restype = vp
L.hsb_create.argtypes = [C.c_int, C.c_int]
L.hsb_upload_matrix_csr.argtypes = [vp, u32, u32, vp, vp, vp, u32]
L.hsb_set_replicas.argtypes = [vp, C.c_int]
L.hsb_upload_vector.argtypes = [vp, vp, C.c_uint]
L.hsb_time_spmv.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
L.hsb_destroy.argtypes = [vp]
L.hsb_last_error.restype = C.c_char_p
impl = {"fixed": 0, "float_pob": 1}[bench.WORKLOADS[wl][1]]
h = L.hsb_create(0, impl)
assert h, L.hsb_last_error()
p = lambda a: a.ctypes.data_as(vp)
ip2, indices, words, xw = [np.ascontiguousarray(a, dtype=np.uint32) for a in (ip2, indices, words, xw)]
assert L.hsb_upload_matrix_csr(h, r2, c2, p(ip2), p(indices), p(words), 0) == 0, L.hsb_last_error()
fmt_guess = 6.5 * indices.size
assert L.hsb_set_replicas(h, max(2, int(np.ceil(2.5 * 126 * 2 ** 20 / fmt_guess)))) == 0
assert L.hsb_upload_vector(h, p(xw), c2) == 0
a, b = C.c_float(), C.c_float()
n = 2048 if wl in ("c1", "c2", "c3", "t95") else 400
res = []
for _ in range(3):
    assert L.hsb_time_spmv(h, n // 8, n, C.byref(a), None) == 0, L.hsb_last_error()
    res.append(a.value * 1e3)
assert L.hsb_time_spmv(h, 16, 256, C.byref(a), C.byref(b)) == 0
print("%-4s %-28s resident us/spmv %s   isolated launch %.2f us" % (wl, os.path.basename(os.environ.get("HSB_LIB", "default")),
                                                                     " ".join("%.2f" % t for t in res), b.value * 1e3))
L.hsb_destroy(h)
