/*
 * hsoracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement ("port") of the reference SpMV
 * hot path of cornell-zhang/HiSparse, in plain C. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (hisparse_b200/, include/) never does.
 *
 * PARITY STATUS: pinned. Every function below is checked (tests/test_oracle_*.py) against
 *   (a) the golden vectors of the reference's own tests (unit_tests/test_io.cpp:110-398) and
 *   (b) the UNMODIFIED reference sources compiled here into oracle/_ref/ (oracle/Makefile,
 *       oracle/ref_driver.cpp) -- formatter, top_wrapper dataflow simulation, compute_ref.
 * One caveat is inherited from (b): Xilinx's ap_fixed.h is not available, so Q8.24
 * rounding/saturation corner cases are pinned to the documented AP_RND/AP_SAT behaviour
 * (oracle/shim/ap_fixed.h), which the reference's own {0,1}-valued tests cannot distinguish.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define HSO_MARKER 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------
 * Q8.24 arithmetic: VAL_T = ap_ufixed<32, 8, AP_RND, AP_SAT>      spmv/libfpga/common.h:35-38
 * ---------------------------------------------------------------------------------------- */

/* float -> VAL_T, as the element-wise std::copy in sw/data_loader.h:76-84 and the vector
 * packing in sw/host.cpp:242 do it: round half up at 2^-24, clamp to [0, 2^32-1]. */
uint32_t hso_q824_from_float(float v) {
    double d = (double)v;
    if (!(d > 0.0)) return 0u;
    double s = floor(ldexp(d, 24) + 0.5);
    if (s >= 4294967296.0) return 0xFFFFFFFFu;
    return (uint32_t)s;
}
void hso_quantize_q824(const float *in, size_t n, uint32_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = hso_q824_from_float(in[i]);
}
/* `VAL_T incr = mat_val * vec_val`                                   spmv/libfpga/pe.h:64 */
static inline uint32_t q824_mul(uint32_t a, uint32_t b) {
    uint64_t p = ((uint64_t)a * (uint64_t)b + (1ull << 23)) >> 24;
    return p > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)p;
}
/* `VAL_T new_q = q_fwd + incr`                                       spmv/libfpga/pe.h:72 */
static inline uint32_t q824_add(uint32_t a, uint32_t b) {
    uint64_t s = (uint64_t)a + (uint64_t)b;
    return s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s;
}
uint32_t hso_q824_mul(uint32_t a, uint32_t b) { return q824_mul(a, b); }
uint32_t hso_q824_mac_chain(const uint32_t *a, const uint32_t *b, size_t n) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; i++) acc = q824_add(acc, q824_mul(a[i], b[i]));
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * SpMV on CSR
 * ---------------------------------------------------------------------------------------- */

/* Fixed-point SpMV, PE semantics (spmv/libfpga/pe.h:63-81) applied per row. Because every
 * term is >= 0 and saturation clamps at a constant, the result is independent of the order
 * in which the shuffle delivers updates; this sequential form is therefore THE fixed result. */
void hso_spmv_q824_csr(uint32_t rows, const uint32_t *indptr, const uint32_t *indices,
                       const uint32_t *val, const uint32_t *x, uint32_t *y) {
    for (uint32_t r = 0; r < rows; r++) {
        uint32_t acc = 0;
        for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++)
            acc = q824_add(acc, q824_mul(val[i], x[indices[i]]));
        y[r] = acc;
    }
}

/* The reference CPU SpMV `compute_ref`: fp32, row by row, separate multiply and add
 * (sw/host.cpp:33-48 == spmv_csim/csim.cpp:143-158). Built with -ffp-contract=off. */
void hso_spmv_f32_csr(uint32_t rows, const uint32_t *indptr, const uint32_t *indices,
                      const float *val, const float *x, float *y) {
    for (uint32_t r = 0; r < rows; r++) {
        float acc = 0.0f;
        for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++) {
            float p = val[i] * x[indices[i]];
            acc = acc + p;
        }
        y[r] = acc;
    }
}
double hso_time_spmv_f32_csr(uint32_t rows, const uint32_t *indptr, const uint32_t *indices,
                             const float *val, const float *x, float *y, int runs) {
    struct timespec t0, t1;
    hso_spmv_f32_csr(rows, indptr, indices, val, x, y);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < runs; k++) hso_spmv_f32_csr(rows, indptr, indices, val, x, y);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / runs;
}
double hso_time_spmv_q824_csr(uint32_t rows, const uint32_t *indptr, const uint32_t *indices,
                              const uint32_t *val, const uint32_t *x, uint32_t *y, int runs) {
    struct timespec t0, t1;
    hso_spmv_q824_csr(rows, indptr, indices, val, x, y);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < runs; k++) hso_spmv_q824_csr(rows, indptr, indices, val, x, y);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / runs;
}

/* fp64 SpMV + sum of |a_i x_i| per row: error budget for the float tolerance tests (not in
 * the reference; used to state "within 1e-5 relative" norm-wise). */
void hso_spmv_f64_csr(uint32_t rows, const uint32_t *indptr, const uint32_t *indices,
                      const float *val, const float *x, double *y, double *sumabs) {
    for (uint32_t r = 0; r < rows; r++) {
        double acc = 0.0, sa = 0.0;
        for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++) {
            double p = (double)val[i] * (double)x[indices[i]];
            acc += p;
            sa += fabs(p);
        }
        y[r] = acc;
        if (sumabs) sumabs[r] = sa;
    }
}

/* ------------------------------------------------------------------------------------------
 * Host-side formatting: CSR -> CPSR                                sw/data_formatter.h
 * ---------------------------------------------------------------------------------------- */

/* util_round_csr_matrix_dim                                  sw/data_formatter.h:15-29 */
void hso_round_dims(uint32_t *rows, uint32_t *cols, uint32_t row_div, uint32_t col_div) {
    if (*rows % row_div) *rows += row_div - *rows % row_div;
    if (*cols % col_div) *cols += col_div - *cols % col_div;
}

/* marker value encodings                                      sw/data_formatter.h:69-74,154-158 */
enum { HSO_VAL_INT = 0, HSO_VAL_FLOAT_BITS = 1, HSO_VAL_Q824 = 2 };
static uint32_t marker_word(uint32_t k, int kind) {
    switch (kind) {
    case HSO_VAL_FLOAT_BITS: return k;                     /* reinterpret_cast<float&>(k) */
    case HSO_VAL_Q824: return k >= 256u ? 0xFFFFFFFFu : (k << 24);   /* VAL_T(k), AP_SAT */
    default: return k;                                     /* integer DataT */
    }
}

typedef struct {
    uint32_t n_packets;     /* = max lane length */
    uint32_t n_slots;       /* number of row packs */
    uint32_t *idx;          /* n_packets * P */
    uint32_t *val;          /* n_packets * P */
    uint32_t *indptr;       /* (n_slots + 1) * P : running lane lengths after each slot */
} hso_block;

typedef struct {
    uint32_t rows, cols;            /* (already rounded) */
    uint32_t P, C, OB, VB;
    uint32_t n_row_parts, n_col_parts;
    int skip_empty_rows, val_kind;
    hso_block *blocks;              /* [j][i][c] */
} hso_cpsr;

static hso_block *blk(const hso_cpsr *m, uint32_t j, uint32_t i, uint32_t c) {
    return &m->blocks[((size_t)j * m->n_col_parts + i) * m->C + c];
}

/* csr2cpsr + util_convert_csr_to_dds + util_pad_marker_end_of_row(_skip/_no_skip) +
 * util_pack_rows, restated as one direct construction.    sw/data_formatter.h:468-544,
 * 256-313, 51-171, 384-446. `val` are 32-bit words already converted to the kernel's VAL_T. */
hso_cpsr *hso_csr2cpsr(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                       const uint32_t *val, uint32_t pack_size, uint32_t out_buf_len,
                       uint32_t vec_buf_len, uint32_t num_channels, int skip_empty_rows, int val_kind) {
    const uint32_t P = pack_size, C = num_channels, S = P * C;
    if (rows % S || cols % P || out_buf_len % S || vec_buf_len % P) return NULL;   /* :475-490 */
    hso_cpsr *m = (hso_cpsr *)calloc(1, sizeof(hso_cpsr));
    m->rows = rows; m->cols = cols; m->P = P; m->C = C; m->OB = out_buf_len; m->VB = vec_buf_len;
    m->n_row_parts = (rows + out_buf_len - 1) / out_buf_len;
    m->n_col_parts = (cols + vec_buf_len - 1) / vec_buf_len;
    m->skip_empty_rows = skip_empty_rows; m->val_kind = val_kind;
    m->blocks = (hso_block *)calloc((size_t)m->n_row_parts * m->n_col_parts * C, sizeof(hso_block));

    for (uint32_t j = 0; j < m->n_row_parts; j++) {
        uint32_t r0 = j * out_buf_len;
        uint32_t nr = (j == m->n_row_parts - 1) ? rows - r0 : out_buf_len;          /* :504-507 */
        uint32_t n_slots = (nr + S - 1) / S;
        /* per (row, col part): [begin,end) into the CSR arrays. Entries of one row that fall in
         * one column partition keep their CSR order (:298-312); column ids need not be sorted. */
        uint32_t *cnt = (uint32_t *)calloc((size_t)nr, sizeof(uint32_t));
        uint32_t *mk = (uint32_t *)calloc((size_t)nr, sizeof(uint32_t));
        for (uint32_t i = 0; i < m->n_col_parts; i++) {
            uint32_t c_lo = i * vec_buf_len, c_hi = c_lo + vec_buf_len;
            for (uint32_t r = 0; r < nr; r++) {
                uint32_t n = 0;
                for (uint32_t e = indptr[r0 + r]; e < indptr[r0 + r + 1]; e++)
                    n += (indices[e] >= c_lo && indices[e] < c_hi);
                cnt[r] = n;
            }
            /* end-of-row marker value per row; 0 = no marker (:51-83 no-skip, :87-171 skip) */
            if (!skip_empty_rows) {
                for (uint32_t r = 0; r < nr; r++) mk[r] = 1;
            } else {
                for (uint32_t r = 0; r < nr; r++) mk[r] = (r < S || cnt[r]) ? 1 : 0;
                for (uint32_t r = 0; r < nr; r++) {
                    if (!mk[r]) continue;
                    for (uint32_t q = r + S; q < nr && !(q < S || cnt[q]); q += S) mk[r]++;
                }
            }
            for (uint32_t c = 0; c < C; c++) {
                hso_block *b = blk(m, j, i, c);
                b->n_slots = n_slots;
                b->indptr = (uint32_t *)calloc((size_t)(n_slots + 1) * P, sizeof(uint32_t));
                for (uint32_t s = 0; s < n_slots; s++)
                    for (uint32_t l = 0; l < P; l++) {
                        uint32_t r = s * S + c * P + l;
                        uint32_t add = (r < nr) ? cnt[r] + (mk[r] ? 1u : 0u) : 0u;
                        b->indptr[(s + 1) * P + l] = b->indptr[s * P + l] + add;
                    }
                uint32_t longest = 0;
                for (uint32_t l = 0; l < P; l++)
                    if (b->indptr[n_slots * P + l] > longest) longest = b->indptr[n_slots * P + l];
                b->n_packets = longest;
                b->idx = (uint32_t *)calloc((size_t)longest * P + 1, sizeof(uint32_t));
                b->val = (uint32_t *)calloc((size_t)longest * P + 1, sizeof(uint32_t));
                for (uint32_t l = 0; l < P; l++) {
                    uint32_t n = 0;
                    for (uint32_t s = 0; s < n_slots; s++) {
                        uint32_t r = s * S + c * P + l;
                        if (r >= nr) continue;
                        for (uint32_t e = indptr[r0 + r]; e < indptr[r0 + r + 1]; e++) {
                            if (indices[e] < c_lo || indices[e] >= c_hi) continue;
                            b->idx[(size_t)n * P + l] = indices[e] - c_lo;               /* :308-309 */
                            b->val[(size_t)n * P + l] = val[e];
                            n++;
                        }
                        if (mk[r]) {
                            b->idx[(size_t)n * P + l] = HSO_MARKER;
                            b->val[(size_t)n * P + l] = marker_word(mk[r], val_kind);
                            n++;
                        }
                    }
                }
            }
        }
        free(cnt); free(mk);
    }
    return m;
}
void hso_cpsr_dims(const hso_cpsr *m, uint32_t out[4]) {
    out[0] = m->rows; out[1] = m->cols; out[2] = m->n_row_parts; out[3] = m->n_col_parts;
}
uint32_t hso_cpsr_block_len(const hso_cpsr *m, uint32_t j, uint32_t i, uint32_t c) { return blk(m, j, i, c)->n_packets; }
uint32_t hso_cpsr_block_slots(const hso_cpsr *m, uint32_t j, uint32_t i, uint32_t c) { return blk(m, j, i, c)->n_slots; }
void hso_cpsr_block_get(const hso_cpsr *m, uint32_t j, uint32_t i, uint32_t c, uint32_t *idx, uint32_t *val, uint32_t *indptr) {
    const hso_block *b = blk(m, j, i, c);
    if (idx) memcpy(idx, b->idx, (size_t)b->n_packets * m->P * 4);
    if (val) memcpy(val, b->val, (size_t)b->n_packets * m->P * 4);
    if (indptr) memcpy(indptr, b->indptr, (size_t)(b->n_slots + 1) * m->P * 4);
}
void hso_cpsr_free(hso_cpsr *m) {
    if (!m) return;
    size_t nb = (size_t)m->n_row_parts * m->n_col_parts * m->C;
    for (size_t k = 0; k < nb; k++) { free(m->blocks[k].idx); free(m->blocks[k].val); free(m->blocks[k].indptr); }
    free(m->blocks); free(m);
}

/* ------------------------------------------------------------------------------------------
 * Per-HBM-channel packet images (what the kernels actually read).
 * sw/host.cpp:163-231 == sw/benchmark.cpp:127-195 == spmv_csim/csim.cpp:229-297.
 * 64-byte packet = 8 x u32 indices then 8 x u32 values (spmv/libfpga/common.h:44-50).
 * m must have been built with pack_size 8 and C = 16 * interleave virtual channels.
 * Returns the number of packets of physical channel pc; writes them to `out` if non-NULL.
 * ---------------------------------------------------------------------------------------- */
size_t hso_channel_image(const hso_cpsr *m, uint32_t pc, uint32_t interleave, uint32_t *out /* 16 words/packet */) {
    const uint32_t IF = interleave, NPC = m->C / IF, np = m->n_row_parts * m->n_col_parts;
    if (m->P != 8 || m->C % IF || pc >= NPC) return 0;
    size_t total_data = 0;
    for (uint32_t ij = 0; ij < np; ij++) {
        uint32_t maxp = 0;
        for (uint32_t f = 0; f < IF; f++) {
            uint32_t n = m->blocks[(size_t)ij * m->C + pc + f * NPC].n_packets;
            if (n > maxp) maxp = n;
        }
        total_data += maxp;
    }
    size_t n_total = (size_t)np * (1 + IF) + total_data * IF;
    if (!out) return n_total;
    memset(out, 0, n_total * 64);
    size_t start = 0, base = (size_t)np * (1 + IF);
    for (uint32_t ij = 0; ij < np; ij++) {
        uint32_t maxp = 0;
        for (uint32_t f = 0; f < IF; f++) {
            const hso_block *b = &m->blocks[(size_t)ij * m->C + pc + f * NPC];
            if (b->n_packets > maxp) maxp = b->n_packets;
        }
        out[((size_t)ij * (1 + IF)) * 16 + 0] = (uint32_t)(start * IF);          /* partition start */
        for (uint32_t f = 0; f < IF; f++) {
            const hso_block *b = &m->blocks[(size_t)ij * m->C + pc + f * NPC];
            uint32_t *hdr = &out[((size_t)ij * (1 + IF) + 1 + f) * 16];
            for (uint32_t l = 0; l < 8; l++) hdr[l] = b->indptr[(size_t)b->n_slots * 8 + l];   /* lane lengths */
            for (uint32_t n = 0; n < b->n_packets; n++) {
                uint32_t *pk = &out[(base + (start + n) * IF + f) * 16];
                memcpy(pk, &b->idx[(size_t)n * 8], 32);
                memcpy(pk + 8, &b->val[(size_t)n * 8], 32);
            }
        }
        start += maxp;
    }
    return n_total;
}

/* ------------------------------------------------------------------------------------------
 * Functional restatement of one kernel invocation (one row partition) over the 16 channel
 * images: CPSR_matrix_loader -> (shuffle) -> vecbuf reader -> (shuffle) -> PE -> packer ->
 * axis_merge -> result_drain, with the two shuffles reduced to their routing function.
 *   loader   spmv/libfpga/spmv_cluster.h:34-107   (fp: spmv-fp/libfpga/spmv_cluster.h:39-129)
 *   x bank   spmv/libfpga/vecbuf_access_unit.h:66-73,124-129  addr=(col/8)%4096, bank=col%8
 *   PE       spmv/libfpga/pe.h:22-90,121-178      (fp: pe-pob.h:28-106, pe-stall.h:39-143)
 *   drain    spmv/libfpga/stream_utils.h:36-75, spmv/spmv_result_drain.cpp:36-113
 * impl: 0 fixed (bit-exact: order independent), 1 float_pob / 2 float_stall (fp32, terms added
 * in stream order per row; the hardware order is timing dependent, so tolerance-level only).
 * x, y: packed raw 32-bit words, natural order. Returns 0, or <0 on malformed images.
 * ---------------------------------------------------------------------------------------- */
int hso_top_wrapper(const uint32_t *const ch[16], const uint32_t *x, uint32_t *y, int impl,
                    uint32_t interleave, uint32_t ob_size, uint32_t vb_size, uint32_t row_part_id,
                    uint32_t part_len, uint32_t num_col_partitions, uint32_t num_partitions,
                    uint32_t num_cols) {
    const uint32_t IF = interleave;
    (void)num_cols;
    uint32_t *acc = (uint32_t *)malloc((size_t)(part_len ? part_len : 1) * 4);
    for (uint32_t pc = 0; pc < 16; pc++) {
        const uint32_t *img = ch[pc];
        memset(acc, 0, (size_t)part_len * 4);                       /* loop_reset_ob, pe.h:131-135 */
        for (uint32_t cp = 0; cp < num_col_partitions; cp++) {
            uint32_t part = row_part_id * num_col_partitions + cp;
            uint32_t start = img[(size_t)part * (1 + IF) * 16];
            uint32_t maxp = 0;
            uint32_t row_idx[64];                                   /* IF <= 8 */
            for (uint32_t f = 0; f < IF; f++)
                for (uint32_t k = 0; k < 8; k++) {
                    uint32_t len = img[((size_t)part * (1 + IF) + 1 + f) * 16 + k];
                    if (len > maxp) maxp = len;
                    row_idx[f * 8 + k] = f * 8 + k;
                }
            size_t base = (size_t)num_partitions * (1 + IF);
            for (uint32_t i = 0; i < maxp * IF; i++) {
                const uint32_t *pk = &img[(base + start + i) * 16];
                uint32_t f = i % IF, n = i / IF;
                const uint32_t *lens = &img[((size_t)part * (1 + IF) + 1 + f) * 16];
                for (uint32_t k = 0; k < 8; k++) {
                    if (n >= lens[k]) continue;
                    uint32_t idx = pk[k], v = pk[8 + k];
                    if (idx == HSO_MARKER) {
                        uint32_t skip = (impl == 0) ? (v >> 24) : v;             /* :82 / fp :104 */
                        row_idx[f * 8 + k] += 8u * skip * IF;
                        continue;
                    }
                    uint32_t r = row_idx[f * 8 + k];
                    if (r >= part_len) continue;                     /* never dumped */
                    if (idx >= vb_size) { free(acc); return -1; }
                    uint32_t xv = x[(size_t)cp * vb_size + idx];
                    if (impl == 0) {
                        acc[r] = q824_add(acc[r], q824_mul(v, xv));
                    } else {
                        float a, b, q;
                        memcpy(&a, &v, 4); memcpy(&b, &xv, 4); memcpy(&q, &acc[r], 4);
                        float p = a * b;
                        q = q + p;
                        memcpy(&acc[r], &q, 4);
                    }
                }
            }
        }
        /* dump + pack + merge + drain: cluster-local row r -> y[part_base + (r/8)*128 + pc*8 + r%8] */
        for (uint32_t r = 0; r < part_len; r++)
            y[(size_t)row_part_id * ob_size + (size_t)(r / 8) * 128 + pc * 8 + (r % 8)] = acc[r];
    }
    free(acc);
    return 0;
}
