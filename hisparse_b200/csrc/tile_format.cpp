// Host-side builder for the tile-stream format (see tile_format.h). This is the B200
// counterpart of the reference's CSR -> CPSR preprocessing (sw/data_formatter.h:468-544):
// same inputs (a CSR with 32-bit indices and 32-bit value words, a row-partition length, a
// column-partition length), different output layout. All passes are parallel (row slabs, tiles,
// slices), so preprocessing is not the single-threaded bottleneck it is in the reference
// (paper Table 8: 0.02 - 10.6 s).
//
//   stage A  column partitioning: per (row partition, column tile) the non-zeros in CSR order
//            with column ids rebased to the tile           (== util_convert_csr_to_dds, :256-313)
//   stage B  per tile: cut row segments into lane streams of <= kMaxStreamLen, sort the streams
//            by length (descending, stable), group 32 per slice  (replaces util_pack_rows'
//            cyclic row->lane assignment, :407-443, by a length-sorted one)
//   stage C  copy every stream into its slice in the warp-coalesced element order
#include "tile_format.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace hsb {

uint32_t choose_tile_cols(uint32_t cols, uint32_t rows, uint64_t nnz) {
    if (cols == 0) return 8;
    // Measured on B200: a matrix whose whole x fits the 224 KB tile limit is
    // fastest as ONE tile (no row is cut, half the lane streams: C3 15.2 -> 12.3 us). Wider matrices want
    // narrower tiles, because the x staging of a tile sits on every CTA's critical path (C2: 2 tiles of
    // 54 K columns 16.6 us, 3-4 tiles 15.1-15.2 us): at most 44,000 columns when rows still have a few
    // entries per tile (longer lane streams, fewer row updates: C4 65.2 -> 62.5 us against 32,768), at most
    // the full 57,344 for hypersparse matrices with less than one entry per (row, tile): these use the narrow
    // layout, where fewer, wider tiles mean fewer lane streams and fewer per-tile start-ups (C5 shard, narrow
    // layout: 0.894 ms at 32 K, 0.874 at 44 K, 0.815 at 56 K; the wide layout had preferred 32 K).
    uint32_t cap = kMaxTileCols;
    if (cols > kMaxTileCols) {
        const uint64_t tiles44 = (cols + 44000u - 1) / 44000u;
        const double per_row_tile = rows ? (double)nnz / ((double)rows * (double)tiles44) : 0.0;
        cap = per_row_tile >= 1.0 ? 44000u : kMaxTileCols;
    }
    if (const char *e = std::getenv("HSB_TILE_COLS")) {          // tuning aid
        uint32_t v = (uint32_t)std::atoi(e) & ~7u;
        if (v >= 8 && v <= kMaxTileCols) cap = v;
    }
    uint32_t n_tiles = (cols + cap - 1) / cap;
    uint32_t w = (cols + n_tiles - 1) / n_tiles;
    w = (w + 7u) & ~7u;
    return std::min(w, kMaxTileCols);
}

bool choose_narrow(uint64_t nnz, uint64_t n_segments) {
    if (const char *e = std::getenv("HSB_NARROW")) return std::atoi(e) != 0;        // A/B aid
    return n_segments > 0 && (double)nnz < kNarrowBelow * (double)n_segments;
}

namespace {

struct Slab {
    uint32_t part, r0, r1;       // rows [r0, r1) of row partition `part`
};
struct Stream {
    uint32_t row;
    uint32_t len;
    uint64_t src;                // position of its first non-zero in the stage-A arrays
};

template <class F> void parallel_for(size_t n, int n_threads, F f) {
    if (n_threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::vector<std::thread> th;
    size_t nt = std::min<size_t>(n_threads, n);
    std::atomic<size_t> next(0);
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&]() { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto &x : th) x.join();
}

// number of lane streams a segment of n non-zeros is cut into (at most max_len each), and the length of piece q
inline uint32_t n_pieces(uint32_t n, uint32_t max_len) { return (n + max_len - 1) / max_len; }
inline uint32_t piece_len(uint32_t n, uint32_t pieces, uint32_t q) { return n / pieces + (q < n % pieces ? 1u : 0u); }

}  // namespace

bool build_tiled(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                 const uint32_t *vals, uint32_t rows_per_part, uint32_t tile_cols, int n_threads,
                 TiledMatrix *out, std::string *err) {
    auto fail = [&](const char *m) { if (err) *err = m; return false; };
    if (tile_cols == 0 || tile_cols > kMaxTileCols || (tile_cols & 7u)) return fail("tile_cols must be a multiple of 8 in [8, 57344]");
    if (rows && indptr[0] != 0) return fail("indptr[0] must be 0");
    for (uint32_t r = 0; r < rows; r++)
        if (indptr[r + 1] < indptr[r]) return fail("indptr is not monotone");
    const uint64_t nnz = rows ? indptr[rows] : 0;
    if (n_threads <= 0) n_threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    if (nnz < (1u << 16)) n_threads = 1;

    TiledMatrix &M = *out;
    M = TiledMatrix();
    M.rows = rows; M.cols = cols; M.nnz = nnz;
    M.rows_per_part = rows_per_part ? rows_per_part : std::max(rows, 1u);
    M.n_row_parts = rows ? (rows + M.rows_per_part - 1) / M.rows_per_part : 0;
    M.tile_cols = tile_cols;
    M.n_col_tiles = std::max(1u, (cols + tile_cols - 1) / tile_cols);
    const uint32_t T = M.n_col_tiles;
    const size_t NT = (size_t)M.n_row_parts * T;

    // ---- stage A -------------------------------------------------------------------------
    std::vector<Slab> slabs;
    std::vector<uint32_t> part_slab_begin(M.n_row_parts + 1, 0);
    for (uint32_t j = 0; j < M.n_row_parts; j++) {
        uint32_t r0 = j * M.rows_per_part, r1 = (uint32_t)std::min<uint64_t>(rows, (uint64_t)r0 + M.rows_per_part);
        uint64_t e0 = indptr[r0], e1 = indptr[r1];
        int ns = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_threads, (e1 - e0) >> 14));
        uint32_t prev = r0;
        for (int s = 1; s <= ns; s++) {
            uint32_t cut = r1;
            if (s < ns) {
                uint64_t target = e0 + (e1 - e0) * s / ns;
                cut = (uint32_t)(std::lower_bound(indptr + prev, indptr + r1, (uint32_t)target) - indptr);
                cut = std::min(std::max(cut, prev), r1);
            }
            if (cut > prev || s == ns) slabs.push_back(Slab{j, prev, cut});
            prev = cut;
        }
        part_slab_begin[j + 1] = (uint32_t)slabs.size();
    }
    const size_t NS = slabs.size();

    std::vector<uint64_t> cnt(NS * T, 0), seg(NS * T, 0);
    std::atomic<bool> bad_col(false);
    parallel_for(NS, n_threads, [&](size_t s) {
        const Slab &sl = slabs[s];
        std::vector<uint32_t> stamp(T, 0xFFFFFFFFu);
        uint64_t *c = &cnt[s * T], *g = &seg[s * T];
        for (uint32_t r = sl.r0; r < sl.r1; r++)
            for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
                uint32_t col = indices[e];
                if (col >= cols) { bad_col = true; continue; }
                uint32_t t = col / tile_cols;
                c[t]++;
                if (stamp[t] != r) { stamp[t] = r; g[t]++; }
            }
    });
    if (bad_col) return fail("column index out of range");

    std::vector<uint64_t> tile_nnz0(NT + 1, 0), tile_seg0(NT + 1, 0);
    std::vector<uint64_t> pos0(NS * T), seg0(NS * T);
    {
        uint64_t p = 0, g = 0;
        for (uint32_t j = 0; j < M.n_row_parts; j++)
            for (uint32_t t = 0; t < T; t++) {
                tile_nnz0[(size_t)j * T + t] = p;
                tile_seg0[(size_t)j * T + t] = g;
                for (uint32_t s = part_slab_begin[j]; s < part_slab_begin[j + 1]; s++) {
                    pos0[(size_t)s * T + t] = p; seg0[(size_t)s * T + t] = g;
                    p += cnt[(size_t)s * T + t]; g += seg[(size_t)s * T + t];
                }
            }
        tile_nnz0[NT] = p; tile_seg0[NT] = g;
    }
    std::vector<uint32_t> a_vals(nnz), a_seg_row(tile_seg0[NT]), a_seg_len(tile_seg0[NT]);
    std::vector<uint16_t> a_cols(nnz);
    parallel_for(NS, n_threads, [&](size_t s) {
        const Slab &sl = slabs[s];
        std::vector<uint64_t> cur(pos0.begin() + s * T, pos0.begin() + (s + 1) * T);
        std::vector<uint64_t> scur(seg0.begin() + s * T, seg0.begin() + (s + 1) * T);
        std::vector<uint32_t> in_row(T, 0);
        std::vector<uint32_t> touched;
        for (uint32_t r = sl.r0; r < sl.r1; r++) {
            for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
                uint32_t col = indices[e], t = col / tile_cols;
                uint64_t p = cur[t]++;
                a_vals[p] = vals[e];
                a_cols[p] = (uint16_t)(col - t * tile_cols + kColBias);
                if (in_row[t]++ == 0) touched.push_back(t);
            }
            for (uint32_t t : touched) {
                a_seg_row[scur[t]] = r;
                a_seg_len[scur[t]] = in_row[t];
                scur[t]++;
                in_row[t] = 0;
            }
            touched.clear();
        }
    });

    // ---- stage B -------------------------------------------------------------------------
    // layout: wide (4 slots per lane and step, row ids in slice_rows) or narrow (1 slot per lane and unit, a
    // row unit in front of every slice) -- see tile_format.h
    M.narrow = choose_narrow(nnz, tile_seg0[NT]);
    const uint32_t max_len = M.narrow ? kNarrowMaxLen : kMaxStreamLen;
    const uint32_t slot_block = M.narrow ? 1u : (uint32_t)kSlotBlock, row_units = M.narrow ? 1u : 0u;
    const uint32_t step_elems = M.step_elems();
    std::vector<uint64_t> t_streams(NT + 1, 0), t_slices(NT + 1, 0), t_steps(NT + 1, 0);
    auto tile_plan = [&](size_t ti, uint32_t *hist /*[kMaxStreamLen+1]*/) {
        std::fill(hist, hist + kMaxStreamLen + 1, 0u);
        for (uint64_t g = tile_seg0[ti]; g < tile_seg0[ti + 1]; g++) {
            uint32_t n = a_seg_len[g], p = n_pieces(n, max_len);
            hist[n / p] += p - n % p;
            if (n % p) hist[n / p + 1] += n % p;
        }
    };
    parallel_for(NT, n_threads, [&](size_t ti) {
        uint32_t hist[kMaxStreamLen + 1];
        tile_plan(ti, hist);
        uint64_t ns = 0, steps = 0, slices = 0;
        // walk lengths in descending order, 32 streams per slice; a slice is as long as its first stream
        uint64_t idx = 0;
        for (uint32_t len = kMaxStreamLen; len >= 1; len--) {
            uint64_t h = hist[len];
            if (!h) continue;
            // slices that START inside this length class
            uint64_t first_start = (idx + kLanes - 1) / kLanes * kLanes;       // next slice boundary >= idx
            if (first_start < idx + h) {
                uint64_t n_start = (idx + h - first_start + kLanes - 1) / kLanes;
                slices += n_start;
                steps += n_start * ((len + slot_block - 1) / slot_block + row_units);
            }
            idx += h;
            ns += h;
        }
        t_streams[ti + 1] = ns; t_slices[ti + 1] = slices; t_steps[ti + 1] = steps;
    });
    for (size_t ti = 0; ti < NT; ti++) {
        t_streams[ti + 1] += t_streams[ti]; t_slices[ti + 1] += t_slices[ti]; t_steps[ti + 1] += t_steps[ti];
    }
    if (t_slices[NT] >= (1ull << 31) || t_steps[NT] >= (1ull << 32) || NT >= (1ull << 24))
        return fail("matrix too large for 32-bit slice / step ids");
    M.n_streams = t_streams[NT];
    const size_t NSL = (size_t)t_slices[NT];
    M.slices.assign(NSL, SliceDesc{0, 0});
    if (!M.narrow) M.slice_rows.assign(NSL * kLanes, rows);
    M.vals.assign((size_t)t_steps[NT] * step_elems, 0u);
    M.cols16.assign((size_t)t_steps[NT] * step_elems, kPadCol);
    M.tiles.resize(NT);
    M.part_slice_begin.assign(M.n_row_parts + 1, 0);
    std::vector<Stream> streams((size_t)t_streams[NT]);

    parallel_for(NT, n_threads, [&](size_t ti) {
        uint32_t hist[kMaxStreamLen + 1];
        tile_plan(ti, hist);
        // descending counting sort: first output index of every length class
        uint64_t start[kMaxStreamLen + 2];
        uint64_t run = t_streams[ti];
        for (uint32_t len = kMaxStreamLen; len >= 1; len--) { start[len] = run; run += hist[len]; }
        uint64_t src = tile_nnz0[ti];
        for (uint64_t g = tile_seg0[ti]; g < tile_seg0[ti + 1]; g++) {
            uint32_t n = a_seg_len[g], p = n_pieces(n, max_len);
            for (uint32_t q = 0; q < p; q++) {
                uint32_t l = piece_len(n, p, q);
                streams[start[l]++] = Stream{a_seg_row[g], l, src};
                src += l;
            }
        }
        TileDesc &td = M.tiles[ti];
        std::memset(&td, 0, sizeof td);
        uint32_t t = (uint32_t)(ti % T);
        td.col_base = t * tile_cols;
        uint32_t width = std::min(tile_cols, cols > td.col_base ? cols - td.col_base : 0u);
        td.col_count = (width + 7u) & ~7u;
        td.row_part = (uint32_t)(ti / T);
        td.slice_begin = (uint32_t)t_slices[ti];
        td.slice_end = (uint32_t)t_slices[ti + 1];
        td.step_begin = (uint32_t)t_steps[ti];
        uint64_t off = t_steps[ti];
        for (uint64_t s = t_slices[ti], i = t_streams[ti]; s < t_slices[ti + 1]; s++, i += kLanes) {
            uint32_t steps = (streams[i].len + slot_block - 1) / slot_block;
            M.slices[s].off = (uint32_t)off;
            M.slices[s].tile_steps = ((uint32_t)ti << 8) | (steps + row_units);
            off += steps + row_units;
            for (uint32_t c = 0; c < steps; c++) td.cnt_ge[c]++;
        }
    });
    for (uint32_t j = 0; j <= M.n_row_parts; j++) M.part_slice_begin[j] = (uint32_t)t_slices[(size_t)j * T];

    // ---- stage C -------------------------------------------------------------------------
    // The order of the non-zeros inside a lane stream is free (fixed point: the sum is order
    // independent; float: within tolerance), so it is chosen here such that the 32 lanes of a slice
    // gather from 32 different shared-memory banks (bank = column % 32) at every slot whenever
    // possible. This is the reference's shuffle unit #1 (spmv/libfpga/shuffle.h:24-99: an arbiter
    // that grants one request per vector-buffer bank per cycle and re-sends the losers) done once,
    // offline, instead of dynamically in hardware.
    const size_t kBlock = 256;                               // slices per task
    parallel_for((NSL + kBlock - 1) / kBlock, n_threads, [&](size_t blk) {
        std::vector<uint32_t> v((size_t)kLanes * kMaxStreamLen);
        std::vector<uint16_t> cidx((size_t)kLanes * kMaxStreamLen);
        for (size_t s = blk * kBlock; s < std::min(NSL, (blk + 1) * kBlock); s++) {
            const size_t ti = M.slices[s].tile_steps >> 8;
            const uint64_t i0 = t_streams[ti] + (s - t_slices[ti]) * kLanes;
            const size_t base = (size_t)M.slices[s].off * step_elems;
            uint32_t rem[kLanes], maxlen = 0;
            for (int lane = 0; lane < kLanes; lane++) {
                rem[lane] = 0;
                uint64_t i = i0 + lane;
                if (M.narrow) M.vals[base + lane] = i < t_streams[ti + 1] ? streams[i].row : rows;     // the row unit
                if (i >= t_streams[ti + 1]) continue;
                const Stream &st = streams[i];
                if (!M.narrow) M.slice_rows[s * kLanes + lane] = st.row;
                rem[lane] = st.len;
                maxlen = std::max(maxlen, st.len);
                for (uint32_t k = 0; k < st.len; k++) {
                    v[(size_t)lane * kMaxStreamLen + k] = a_vals[st.src + k];
                    cidx[(size_t)lane * kMaxStreamLen + k] = a_cols[st.src + k];
                }
            }
            for (uint32_t k = 0; k < maxlen; k++) {
                uint32_t busy = 0;                              // banks granted at this slot
                for (int q = 0; q < kLanes; q++) {
                    const int lane = (q + (int)k) & (kLanes - 1);    // rotating priority, like the arbiter
                    uint32_t n = rem[lane];
                    if (!n) continue;
                    uint32_t *lv = &v[(size_t)lane * kMaxStreamLen];
                    uint16_t *lc = &cidx[(size_t)lane * kMaxStreamLen];
                    uint32_t pick = 0;
                    const uint32_t window = std::min(n, 16u);
                    for (uint32_t c = 0; c < window; c++)
                        if (!((busy >> (lc[c] & 31u)) & 1u)) { pick = c; break; }
                    busy |= 1u << (lc[pick] & 31u);
                    const size_t e = M.narrow ? base + (size_t)(1 + k) * kUnitElems + lane : slice_elem(base, lane, k);
                    M.vals[e] = lv[pick];
                    M.cols16[e] = lc[pick];
                    lv[pick] = lv[n - 1];                       // order inside a stream is free
                    lc[pick] = lc[n - 1];
                    rem[lane] = n - 1;
                }
            }
        }
    });
    return true;
}

// Cost of finishing a slice (row-id load, warp vote, up to 32 row updates) in units of one step (768 B of
// matrix stream) for the equal-cost cuts of the planner. Per-CTA traces fit 1.5-1.7 steps, but what the cuts
// have to balance is the END of a length-sorted tile, where slices are one or two steps long and a warp is
// bound by the latency of a slice rather than by its bytes. Measured optimum (B200, final kernel): C2 (11.6
// steps per slice on average) 15.0 us at 2.5, 14.05 us at 6-8, 14.6 us at 15; C4 (3.1 steps per slice) 64.1 us
// at 2.5, 67.4 us at 6. The rule that fits both: 0.6 x the average steps per slice of the launch, within
// [2.5, 8], when at least a quarter of the slices are one or two steps long; 2.5 otherwise. HSB_SLICE_COST
// overrides it, HSB_DEBUG_PLAN prints the choice.
namespace {
thread_local double g_slice_cost = 2.5;
}
double slice_cost() { return g_slice_cost; }
static void set_slice_cost_for(const TiledMatrix &m, uint32_t tile_begin, uint32_t tile_end) {
    static const double forced = [] {
        const char *e = std::getenv("HSB_SLICE_COST");
        return e ? std::atof(e) : 0.0;
    }();
    if (forced > 0.0) { g_slice_cost = forced; return; }
    // narrow layout: a slice already pays for its row unit in units; measured on a C5 shard (B200, same box):
    // 0.778 ms at 1, 0.807 at 1.7, 0.841 at 2.5, 0.937 at 4, 1.06 at 6
    if (m.narrow) { g_slice_cost = 1.0; return; }
    uint64_t steps = 0, slices = 0, short_slices = 0;
    for (uint32_t t = tile_begin; t < tile_end; t++) {
        const TileDesc &td = m.tiles[t];
        if (td.slice_end == td.slice_begin) continue;
        const SliceDesc &last = m.slices[td.slice_end - 1];
        steps += last.off + (last.tile_steps & 0xFFu) - td.step_begin;
        slices += td.slice_end - td.slice_begin;
        short_slices += (td.slice_end - td.slice_begin) - td.cnt_ge[2];        // slices of one or two steps
    }
    const double avg = slices ? (double)steps / (double)slices : 1.0;
    // Only a mix of long and short slices needs the higher cost; when (nearly) all slices are long -- the
    // pruned transformer layers: every slice 32 steps -- the cuts inside a CTA should stay close to equal steps.
    const bool mixed = slices && (double)short_slices >= 0.25 * (double)slices;
    g_slice_cost = mixed ? std::min(8.0, std::max(2.5, 0.6 * avg)) : 2.5;
    static const bool debug = std::getenv("HSB_DEBUG_PLAN") != nullptr;
    if (debug)
        std::fprintf(stderr, "[hsb plan] tiles %u..%u: %llu slices, %.2f steps per slice, %.0f %% of one or two steps -> slice cost %.2f\n",
                     tile_begin, tile_end, (unsigned long long)slices, avg, slices ? 100.0 * short_slices / slices : 0.0, g_slice_cost);
}
namespace {
// tile-relative step position at which the cost prefix (steps + kSliceCost * slices started) of
// tile `td` reaches w
uint32_t step_at_cost(const TiledMatrix &m, const TileDesc &td, double w) {
    const double B = slice_cost();
    // cost prefix before slice i: S(i) + B*i ; inside slice i after k steps: S(i) + B*i + B + k
    uint32_t lo = 0, hi = td.slice_end - td.slice_begin;       // find last slice whose start cost <= w
    auto start_cost = [&](uint32_t i) {
        return (double)(m.slices[td.slice_begin + i].off - td.step_begin) + B * i;
    };
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) / 2;
        if (start_cost(mid) <= w) lo = mid; else hi = mid;
    }
    const SliceDesc &sd = m.slices[td.slice_begin + lo];
    uint32_t steps = sd.tile_steps & 0xFFu, s_before = sd.off - td.step_begin;
    double inside = w - start_cost(lo) - B;
    uint32_t k = inside <= 0 ? 0u : (uint32_t)std::min<double>(steps, inside + 0.5);
    if (m.narrow) k = 2 * k >= steps ? steps : 0u;             // narrow layout: shares begin and end at slice boundaries
    return s_before + k;
}
uint32_t tile_steps_total(const TiledMatrix &m, const TileDesc &td) {
    if (td.slice_end == td.slice_begin) return 0;
    const SliceDesc &last = m.slices[td.slice_end - 1];
    return last.off + (last.tile_steps & 0xFFu) - td.step_begin;
}
}  // namespace

namespace {
// tile-relative slice containing step t (t < total steps of the tile)
uint32_t slice_of_step_host(const TiledMatrix &m, const TileDesc &td, uint32_t t) {
    uint32_t lo = 0, hi = td.slice_end - td.slice_begin;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) / 2;
        if (m.slices[td.slice_begin + mid].off - td.step_begin <= t) lo = mid; else hi = mid;
    }
    return lo;
}
// cost of everything before tile-relative step t
double cost_at_step(const TiledMatrix &m, const TileDesc &td, uint32_t t) {
    if (td.slice_end == td.slice_begin) return 0;
    if (t >= tile_steps_total(m, td)) return tile_steps_total(m, td) + slice_cost() * (td.slice_end - td.slice_begin);
    uint32_t i = slice_of_step_host(m, td, t);
    uint32_t s0 = m.slices[td.slice_begin + i].off - td.step_begin;
    return t + slice_cost() * (i + (t > s0 ? 1 : 0));
}
Segment make_segment(const TiledMatrix &m, uint32_t tile, uint32_t t_lo, uint32_t t_hi) {
    const TileDesc &td = m.tiles[tile];
    Segment g;
    std::memset(&g, 0, sizeof g);
    g.tile = tile; g.t_lo = t_lo; g.t_hi = t_hi;
    g.col_base = td.col_base; g.col_count = td.col_count; g.slice_begin = td.slice_begin;
    g.n_slices = td.slice_end - td.slice_begin; g.step_begin = td.step_begin;
    std::memcpy(g.cnt_ge, td.cnt_ge, sizeof g.cnt_ge);
    const double c0 = cost_at_step(m, td, t_lo), c1 = cost_at_step(m, td, t_hi);
    const uint32_t total = tile_steps_total(m, td);
    if (t_hi > t_lo && t_lo < total) {
        g.comb_first = slice_of_step_host(m, td, t_lo);
        const uint32_t span = slice_of_step_host(m, td, t_hi - 1) - g.comb_first + 1;
        static const bool combine = [] { const char *e = std::getenv("HSB_COMBINE"); return !e || std::atoi(e) != 0; }();   // A/B aid
        g.comb_n = (!m.narrow && combine && span <= kCombineSlots) ? span : 0u;
    }
    for (int w = 0; w <= kWarpsPerCta; w++) {
        uint32_t t = w == 0 ? t_lo : (w == kWarpsPerCta ? t_hi : step_at_cost(m, td, c0 + (c1 - c0) * w / (double)kWarpsPerCta));
        t = std::min(std::max(t, w ? g.warp_t[w - 1] : t_lo), t_hi);
        g.warp_t[w] = t;
        if (w < kWarpsPerCta) g.warp_slice[w] = t < total ? slice_of_step_host(m, td, t) : g.n_slices;
    }
    return g;
}
}  // namespace

void plan_launch(const TiledMatrix &m, uint32_t tile_begin, uint32_t tile_end, uint32_t ctas,
                 std::vector<uint32_t> *cta_seg, std::vector<Segment> *segs) {
    cta_seg->assign(ctas + 1, (uint32_t)segs->size());
    set_slice_cost_for(m, tile_begin, tile_end);
    std::vector<uint32_t> live;
    std::vector<double> cost;
    double total = 0;
    for (uint32_t t = tile_begin; t < tile_end; t++) {
        const TileDesc &td = m.tiles[t];
        if (td.slice_end == td.slice_begin) continue;
        live.push_back(t);
        cost.push_back((double)tile_steps_total(m, td) + slice_cost() * (td.slice_end - td.slice_begin));
        total += cost.back();
    }
    if (live.empty() || ctas == 0) return;
    if (live.size() <= ctas) {
        // whole CTAs per tile, proportional to cost (largest remainder, at least one each)
        std::vector<uint32_t> n(live.size(), 1);
        uint32_t used = (uint32_t)live.size();
        std::vector<double> want(live.size());
        for (size_t i = 0; i < live.size(); i++) want[i] = cost[i] / total * ctas;
        while (used < ctas) {
            size_t best = 0;
            double gap = -1e300;
            for (size_t i = 0; i < live.size(); i++)
                if (want[i] - n[i] > gap) { gap = want[i] - n[i]; best = i; }
            n[best]++;
            used++;
        }
        uint32_t b = 0;
        for (size_t i = 0; i < live.size(); i++) {
            const TileDesc &td = m.tiles[live[i]];
            const uint32_t steps = tile_steps_total(m, td);
            uint32_t prev = 0;
            for (uint32_t k = 1; k <= n[i]; k++) {
                uint32_t cut = k == n[i] ? steps : std::max(prev, std::min(steps, step_at_cost(m, td, cost[i] * k / n[i])));
                (*cta_seg)[b] = (uint32_t)segs->size();
                if (cut > prev) segs->push_back(make_segment(m, live[i], prev, cut));
                b++;
                prev = cut;
            }
        }
        for (; b <= ctas; b++) (*cta_seg)[b] = (uint32_t)segs->size();
        return;
    }
    // more tiles than CTAs: equal-cost split of the tile sequence; a CTA runs whole tiles and at most a
    // partial tile at either end. With many tiles per CTA the sequence is first interleaved -- position
    // b * per + k holds tile k * ctas + b -- so that at any moment the CTAs work on NEIGHBOURING tiles
    // (the same row partition, adjacent column ranges) instead of 'ctas' far-apart regions: the row
    // accumulators the concurrent CTAs update then form a window that stays in L2 (matrices with a
    // diagonal band: C5), and so do the x tiles being staged.
    static const bool interleave = [] { const char *e = std::getenv("HSB_TILE_INTERLEAVE"); return !e || std::atoi(e) != 0; }();
    if (interleave && live.size() >= 2 * (size_t)ctas) {
        const size_t n = live.size(), per = (n + ctas - 1) / ctas;
        std::vector<uint32_t> l2;
        std::vector<double> c2;
        l2.reserve(n); c2.reserve(n);
        for (size_t b = 0; b < ctas; b++)
            for (size_t k = 0; k < per; k++) {
                const size_t i = k * ctas + b;
                if (i < n) { l2.push_back(live[i]); c2.push_back(cost[i]); }
            }
        live.swap(l2);
        cost.swap(c2);
    }
    size_t i = 0;
    uint32_t prev = 0;                 // next unassigned step of tile live[i]
    double done = 0;                   // cost of everything before tile live[i]
    for (uint32_t b = 0; b < ctas; b++) {
        (*cta_seg)[b] = (uint32_t)segs->size();
        const double target = total * (b + 1) / ctas;
        while (i < live.size()) {
            const TileDesc &td = m.tiles[live[i]];
            const uint32_t steps = tile_steps_total(m, td);
            if (b + 1 == ctas || done + cost[i] <= target) {            // the rest of this tile
                if (steps > prev) segs->push_back(make_segment(m, live[i], prev, steps));
                done += cost[i];
                i++;
                prev = 0;
                continue;
            }
            uint32_t cut = std::max(prev, std::min(steps, step_at_cost(m, td, target - done)));
            if (cut > prev) segs->push_back(make_segment(m, live[i], prev, cut));
            prev = cut;
            break;
        }
    }
    (*cta_seg)[ctas] = (uint32_t)segs->size();
}

}  // namespace hsb
