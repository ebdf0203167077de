"""Measurement aid (uses the oracle to build the reference channel images, hence under tests/): time from host CSR and from
the reference CPSR channel images of the C2-sized matrix to a resident matrix. HSB_DEBUG_PLAN=1 prints the phases."""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from hisparse_b200 import capi, matgen
from oracle import hsoracle
port = hsoracle.Port()
rows, cols, indptr, indices, data = matgen.rmat_csr(107614, 13_670_000, 0xC0FFEE02)
r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
words = port.quantize((data * np.float32(0.05)).astype(np.float32))
cfg = capi.get_config(0)
m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, cfg.logical_ob_size, cfg.logical_vb_size, 16, True, hsoracle.VAL_Q824)
images = m.channel_images(1)
ctx = capi.Context(0, 0)
for rep in range(4):
    t0 = time.perf_counter(); ctx.upload_matrix_csr(r2, c2, ip2, indices, words); t1 = time.perf_counter()
    ctx.upload_matrix_cpsr(images, m.n_row_parts, m.n_col_parts, r2, c2); t2 = time.perf_counter()
    print("csr %.1f ms  cpsr %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1)), flush=True)
