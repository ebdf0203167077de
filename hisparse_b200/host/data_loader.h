// spmv::io -- CSR/CSC containers and the .npz dataset loader, with the names and signatures of
// the reference's sw/data_loader.h (CSRMatrix :19-30, create_csr_matrix :35-47,
// load_csr_matrix_from_float_npz :51-70, csr_matrix_convert_from_float :76-84, CSCMatrix :93-104,
// csr2csc :109-144). Own implementation: the reference depends on the external cnpy library; this
// one reads the scipy `.npz` (a zip of .npy members, stored or deflated) with zlib directly.
#ifndef HISPARSE_B200_HOST_DATA_LOADER_H_
#define HISPARSE_B200_HOST_DATA_LOADER_H_

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace spmv {
namespace io {

template <typename data_type> struct CSRMatrix {
    uint32_t num_rows = 0;
    uint32_t num_cols = 0;
    std::vector<data_type> adj_data;
    std::vector<uint32_t> adj_indices;
    std::vector<uint32_t> adj_indptr;
};

template <typename data_type>
CSRMatrix<data_type> create_csr_matrix(uint32_t num_rows, uint32_t num_cols, std::vector<data_type> const &adj_data,
                                       std::vector<uint32_t> const &adj_indices,
                                       std::vector<uint32_t> const &adj_indptr) {
    CSRMatrix<data_type> m;
    m.num_rows = num_rows;
    m.num_cols = num_cols;
    m.adj_data = adj_data;
    m.adj_indices = adj_indices;
    m.adj_indptr = adj_indptr;
    return m;
}

namespace detail {

struct NpyArray {
    std::string descr;              // e.g. "<f4", "<i8", "<u4", "<i4"
    std::vector<size_t> shape;
    std::vector<unsigned char> bytes;
    size_t count() const { size_t n = 1; for (size_t d : shape) n *= d; return n; }
    size_t word() const { return (size_t)std::atoi(descr.c_str() + 2); }
};

inline uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

inline NpyArray parse_npy(const std::vector<unsigned char> &raw) {
    if (raw.size() < 10 || std::memcmp(raw.data(), "\x93NUMPY", 6) != 0) throw std::runtime_error("not an .npy member");
    size_t hlen, off;
    if (raw[6] == 1) { hlen = rd16(&raw[8]); off = 10; } else { hlen = rd32(&raw[8]); off = 12; }
    std::string hdr((const char *)&raw[off], hlen);
    NpyArray a;
    size_t d = hdr.find("'descr'");
    size_t q0 = hdr.find('\'', d + 7), q1 = hdr.find('\'', q0 + 1);
    a.descr = hdr.substr(q0 + 1, q1 - q0 - 1);
    if (hdr.find("'fortran_order': True") != std::string::npos) throw std::runtime_error("fortran order not supported");
    size_t s0 = hdr.find('(', hdr.find("'shape'")), s1 = hdr.find(')', s0);
    std::string sh = hdr.substr(s0 + 1, s1 - s0 - 1);
    for (size_t p = 0; p < sh.size();) {
        while (p < sh.size() && (sh[p] < '0' || sh[p] > '9')) p++;
        if (p >= sh.size()) break;
        a.shape.push_back(std::strtoull(sh.c_str() + p, nullptr, 10));
        while (p < sh.size() && sh[p] >= '0' && sh[p] <= '9') p++;
    }
    a.bytes.assign(raw.begin() + off + hlen, raw.end());
    return a;
}

// minimal zip reader: walks the local file headers (scipy/numpy write sizes up front)
inline std::map<std::string, NpyArray> npz_load(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<unsigned char> z((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::map<std::string, NpyArray> out;
    // central directory: find end-of-central-directory record
    size_t eocd = std::string::npos;
    for (size_t i = z.size() >= 22 ? z.size() - 22 : 0;; i--) {
        if (rd32(&z[i]) == 0x06054b50) { eocd = i; break; }
        if (i == 0) break;
    }
    if (eocd == std::string::npos) throw std::runtime_error("not a zip file: " + path);
    size_t n = rd16(&z[eocd + 10]), cd = rd32(&z[eocd + 16]);
    for (size_t e = 0; e < n; e++) {
        if (rd32(&z[cd]) != 0x02014b50) throw std::runtime_error("corrupt zip central directory");
        uint16_t method = rd16(&z[cd + 10]);
        uint64_t csize = rd32(&z[cd + 20]), usize = rd32(&z[cd + 24]);
        uint16_t nlen = rd16(&z[cd + 28]), xlen = rd16(&z[cd + 30]), clen = rd16(&z[cd + 32]);
        uint64_t lho = rd32(&z[cd + 42]);
        std::string name((const char *)&z[cd + 46], nlen);
        // zip64 extra field (numpy uses it for members written through a stream)
        for (size_t x = cd + 46 + nlen; x + 4 <= cd + 46 + nlen + xlen;) {
            uint16_t id = rd16(&z[x]), sz = rd16(&z[x + 2]);
            if (id == 0x0001) {
                size_t q = x + 4;
                if (usize == 0xFFFFFFFFu) { std::memcpy(&usize, &z[q], 8); q += 8; }
                if (csize == 0xFFFFFFFFu) { std::memcpy(&csize, &z[q], 8); q += 8; }
                if (lho == 0xFFFFFFFFu) { std::memcpy(&lho, &z[q], 8); }
            }
            x += 4 + sz;
        }
        size_t data = lho + 30 + rd16(&z[lho + 26]) + rd16(&z[lho + 28]);
        std::vector<unsigned char> raw(usize);
        if (method == 0) {
            std::memcpy(raw.data(), &z[data], usize);
        } else if (method == 8) {
            z_stream s;
            std::memset(&s, 0, sizeof s);
            if (inflateInit2(&s, -MAX_WBITS) != Z_OK) throw std::runtime_error("zlib init failed");
            s.next_in = &z[data]; s.avail_in = (uInt)csize;
            s.next_out = raw.data(); s.avail_out = (uInt)usize;
            int rc = inflate(&s, Z_FINISH);
            inflateEnd(&s);
            if (rc != Z_STREAM_END) throw std::runtime_error("inflate failed for member " + name);
        } else {
            throw std::runtime_error("unsupported zip compression method");
        }
        if (name.size() > 4 && name.substr(name.size() - 4) == ".npy") name.resize(name.size() - 4);
        out[name] = parse_npy(raw);
        cd += 46 + nlen + xlen + clen;
    }
    return out;
}

template <typename T> std::vector<T> as_vector(const NpyArray &a) {
    std::vector<T> v(a.count());
    const char kind = a.descr.size() > 1 ? a.descr[1] : '?';
    const size_t w = a.word();
    for (size_t i = 0; i < v.size(); i++) {
        const unsigned char *p = &a.bytes[i * w];
        if (kind == 'f' && w == 4) { float x; std::memcpy(&x, p, 4); v[i] = (T)x; }
        else if (kind == 'f' && w == 8) { double x; std::memcpy(&x, p, 8); v[i] = (T)x; }
        else if (w == 4) { int32_t x; std::memcpy(&x, p, 4); v[i] = (T)(kind == 'u' ? (uint32_t)x : x); }
        else if (w == 8) { int64_t x; std::memcpy(&x, p, 8); v[i] = (T)x; }
        else throw std::runtime_error("unsupported dtype " + a.descr);
    }
    return v;
}

}  // namespace detail

// scipy.sparse.save_npz layout: `shape` (2 ints), `data`, `indices`, `indptr` (+ `format`).
// The reference reads shape[0] and shape[2] as u32 (i.e. an int64 pair), data as f32 and the index
// arrays as u32 (sw/data_loader.h:51-70); here the dtypes are taken from the file.
inline CSRMatrix<float> load_csr_matrix_from_float_npz(std::string csr_float_npz_path) {
    auto npz = detail::npz_load(csr_float_npz_path);
    for (const char *k : {"shape", "data", "indices", "indptr"})
        if (!npz.count(k)) throw std::runtime_error(std::string("npz member missing: ") + k);
    auto shape = detail::as_vector<uint64_t>(npz["shape"]);
    CSRMatrix<float> m;
    m.num_rows = (uint32_t)shape.at(0);
    m.num_cols = (uint32_t)shape.at(1);
    m.adj_data = detail::as_vector<float>(npz["data"]);
    m.adj_indices = detail::as_vector<uint32_t>(npz["indices"]);
    m.adj_indptr = detail::as_vector<uint32_t>(npz["indptr"]);
    if (m.adj_indptr.size() != (size_t)m.num_rows + 1 || m.adj_indices.size() != m.adj_data.size())
        throw std::runtime_error("inconsistent CSR arrays in " + csr_float_npz_path);
    return m;
}

template <typename data_type> CSRMatrix<data_type> csr_matrix_convert_from_float(CSRMatrix<float> const &in) {
    CSRMatrix<data_type> out;
    out.num_rows = in.num_rows;
    out.num_cols = in.num_cols;
    out.adj_data.reserve(in.adj_data.size());
    for (float v : in.adj_data) out.adj_data.push_back(data_type(v));
    out.adj_indices = in.adj_indices;
    out.adj_indptr = in.adj_indptr;
    return out;
}

template <typename data_type> struct CSCMatrix {
    uint32_t num_rows = 0;
    uint32_t num_cols = 0;
    std::vector<data_type> adj_data;
    std::vector<uint32_t> adj_indices;
    std::vector<uint32_t> adj_indptr;
};

template <typename data_type> CSCMatrix<data_type> csr2csc(CSRMatrix<data_type> const &csr) {
    CSCMatrix<data_type> csc;
    csc.num_rows = csr.num_rows;
    csc.num_cols = csr.num_cols;
    const size_t nnz = csr.adj_indices.size();
    csc.adj_data.resize(nnz);
    csc.adj_indices.resize(nnz);
    csc.adj_indptr.assign((size_t)csr.num_cols + 1, 0);
    for (uint32_t c : csr.adj_indices) csc.adj_indptr[c + 1]++;
    for (uint32_t c = 0; c < csr.num_cols; c++) csc.adj_indptr[c + 1] += csc.adj_indptr[c];
    std::vector<uint32_t> fill(csc.adj_indptr.begin(), csc.adj_indptr.end() - 1);
    for (uint32_t r = 0; r < csr.num_rows; r++)
        for (uint32_t e = csr.adj_indptr[r]; e < csr.adj_indptr[r + 1]; e++) {
            uint32_t at = fill[csr.adj_indices[e]]++;
            csc.adj_indices[at] = r;
            csc.adj_data[at] = csr.adj_data[e];
        }
    return csc;
}

template <typename data_type> CSCMatrix<data_type> csc_matrix_convert_from_float(CSCMatrix<float> const &in) {
    CSCMatrix<data_type> out;
    out.num_rows = in.num_rows;
    out.num_cols = in.num_cols;
    for (float v : in.adj_data) out.adj_data.push_back(data_type(v));
    out.adj_indices = in.adj_indices;
    out.adj_indptr = in.adj_indptr;
    return out;
}

}  // namespace io
}  // namespace spmv
#endif
