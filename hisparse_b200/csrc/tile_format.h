// B200 tile-stream matrix format ("TS-CPSR") -- the HBM layout the sm_100a SpMV kernels read.
//
// It plays the role of the reference's CPSR channel images (sw/data_formatter.h:194-238,
// sw/host.cpp:163-231) and keeps CPSR's central idea -- every processing lane streams whole rows
// (sw/data_formatter.h:407-443: row r goes to channel (r/8)%C, lane r%8) -- re-shaped for
// 32-wide warps, 128-byte coalescing and a load-balanced persistent grid:
//
//   * rows are cut into row partitions (the reference's LOGICAL_OB_SIZE cut,
//     sw/data_formatter.h:494), columns into tiles of <= 57344 columns (the role of the
//     reference's LOGICAL_VB_SIZE = 32768-word vector buffer, spmv/libfpga/common.h:165,179, sized
//     for B200's 227 KB of shared memory per CTA) so that the x tile of a work unit fits in one
//     CTA's shared memory and a local column id fits 16 bits;
//   * inside a (row partition, column tile) every non-empty row segment becomes one or more
//     LANE STREAMS of at most kMaxStreamLen non-zeros (long rows are split so that no lane
//     serialises a hub row -- the reference instead pads every lane to the longest,
//     sw/host.cpp:184-199);
//   * lane streams are sorted by length and packed 32 at a time into SLICES; a slice is padded
//     to its longest stream (rounded up to 4), which after sorting costs a few percent, and is
//     stored so that each warp-wide 128-bit value load / 64-bit column load is one contiguous
//     512 B / 256 B run: vals[len/4][32 lanes][4], cols[len/4][32 lanes][4];
//   * slice_rows[s][lane] is the matrix row a lane's partial sum is added to (the row the
//     reference recovers by counting end-of-row markers, spmv/libfpga/spmv_cluster.h:78-83);
//     no in-band markers are needed.
//
// 6 bytes per stored non-zero slot + 4 bytes per lane stream + 8 bytes per slice.
//
// NARROW layout (hypersparse matrices: on average fewer than kNarrowBelow non-zeros per (row, tile) segment --
// the C5 shards, where nearly every segment is a single entry). Rounding every stream up to 4 slots would
// double the bytes, and a slice would be over before the load of its row ids has returned. So:
//   * the unit of the stream is 32 elements (one per lane): 128 B of value words + 64 B of column ids;
//   * lane streams hold at most kNarrowMaxLen non-zeros; a slice of 32 streams, padded to its longest stream L,
//     is ONE ROW UNIT (the value word of lane l is the matrix row of stream l, `rows` for an unused lane; column
//     ids 0) followed by L step units (slot k of every lane). The row ids travel IN the stream, so the kernel's
//     prefetch ring covers their latency like that of the values;
//   * slice_rows is empty; SliceDesc.off / TileDesc.step_begin / cnt_ge count UNITS: tile-relative slice i has
//     L(i) = #{c : cnt_ge[c] > i} step units and starts at unit i + sum_c min(i, cnt_ge[c]); SliceDesc's low byte
//     holds 1 + L;
//   * warp shares and CTA segments are cut at slice boundaries only (slices are at most 33 units long).
// 6 bytes per stored slot, padding only up to the longest of 32 length-sorted streams: C5 shard 8.6 B per non-zero
// against 16.3 in the wide layout.
#ifndef HISPARSE_B200_TILE_FORMAT_H_
#define HISPARSE_B200_TILE_FORMAT_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace hsb {

constexpr int kLanes = 32;
#ifndef HSB_WARPS_PER_CTA
#define HSB_WARPS_PER_CTA 32
#endif
#ifndef HSB_CTAS_PER_SM
#define HSB_CTAS_PER_SM 1
#endif
constexpr int kCtasPerSm = HSB_CTAS_PER_SM;               // resident CTAs per SM the grid is sized for (tuning builds)
constexpr int kWarpsPerCta = HSB_WARPS_PER_CTA;           // warps of the kernel's one CTA per SM (the planner cuts shares for them)
constexpr int kSlotBlock = 4;                            // non-zeros per lane per load step
constexpr int kStepElems = kLanes * kSlotBlock;          // 128 elements per slice step
constexpr uint32_t kMaxStreamLen = 128;                  // non-zeros per lane stream (keeps 32-bit partial sums exact)
constexpr uint32_t kMaxTileCols = 57344;                 // 224 KB of x in shared memory (of 227 KB per CTA), 16-bit ids
constexpr uint16_t kColBias = 8;                         // stored column id = tile-local column + 8 ...
constexpr uint16_t kPadCol = 0;                          // ... so that id 0 (padding slots) can point at a constant 0 word
constexpr int kUnitElems = kLanes;                       // narrow layout: elements per unit (one slot per lane)
constexpr uint32_t kNarrowMaxLen = 32;                   // narrow layout: non-zeros per lane stream (cnt_ge has 32 entries)
constexpr double kNarrowBelow = 2.5;                     // average non-zeros per (row, tile) segment below which the narrow layout is used

struct SliceDesc {
    uint32_t off;           // element offset of the slice / kStepElems
    uint32_t tile_steps;    // (tile index << 8) | steps, steps = padded stream length / 4  (1..32)
};

struct TileDesc {
    uint32_t col_base;      // first column of the tile (multiple of 8)
    uint32_t col_count;     // columns in the tile, rounded up to a multiple of 8
    uint32_t slice_begin;   // [slice_begin, slice_end) in the global slice list
    uint32_t slice_end;
    uint32_t row_part;      // row partition the tile belongs to
    uint32_t step_begin;    // element offset of the tile's first slice / kStepElems
    uint32_t pad_[2];
    // Slices of a tile are sorted by length, so the whole per-slice geometry is 32 numbers:
    // cnt_ge[c] = slices of this tile with more than c steps. Slice i (tile-relative) has
    // steps(i) = #{c : cnt_ge[c] > i} and starts at step S(i) = sum_c min(i, cnt_ge[c]).
    uint32_t cnt_ge[32];
};
constexpr int kMaxSteps = 32;

struct TiledMatrix {
    uint32_t rows = 0, cols = 0;            // as given (rows may include padding rows)
    uint64_t nnz = 0;
    uint32_t rows_per_part = 0, n_row_parts = 0, n_col_tiles = 0, tile_cols = 0;
    uint64_t n_streams = 0;                 // lane streams (row segments after splitting)
    bool narrow = false;                    // narrow layout (see the header comment): units of 32 elements, row ids in the stream
    std::vector<uint32_t> vals;             // n_elems words
    std::vector<uint16_t> cols16;           // n_elems local column ids
    std::vector<SliceDesc> slices;          // n_slices (host side only: planning and inspection)
    std::vector<uint32_t> slice_rows;       // n_slices * 32 ; `rows` (one past the end) marks an unused lane
    std::vector<TileDesc> tiles;            // n_row_parts * n_col_tiles, row-partition major
    std::vector<uint32_t> part_slice_begin; // n_row_parts + 1
    uint32_t step_elems() const { return narrow ? (uint32_t)kUnitElems : (uint32_t)kStepElems; }   // elements per step / unit
    size_t n_slices() const { return slices.size(); }
    size_t n_elems() const { return vals.size(); }
    size_t format_bytes() const {
        // what the kernel reads from HBM (the SliceDesc list stays on the host)
        return vals.size() * 4 + cols16.size() * 2 + slice_rows.size() * 4 + tiles.size() * sizeof(TileDesc);
    }
};

// element index of slot k (0..len-1) of lane l inside a slice that starts at element `base`
inline size_t slice_elem(size_t base, int lane, uint32_t k) {
    return base + (size_t)(k / kSlotBlock) * kStepElems + (size_t)lane * kSlotBlock + (k % kSlotBlock);
}

// layout choice: narrow when the matrix has on average fewer than kNarrowBelow non-zeros per (row, tile) segment
// (HSB_NARROW=0 / 1 forces the wide / narrow layout: A/B aid)
bool choose_narrow(uint64_t nnz, uint64_t n_segments);

// Choose a tile width (multiple of 8, widths equalised): one tile when x fits shared memory, otherwise tiles
// of at most 44,000 columns, or the full 57,344 for hypersparse matrices (less than one entry per row and tile).
uint32_t choose_tile_cols(uint32_t cols, uint32_t rows, uint64_t nnz);

// CSR (32-bit value words, passed through untouched) -> tile streams. rows_per_part == 0 means
// one row partition. Returns false and fills *err on malformed input.
bool build_tiled(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                 const uint32_t *vals, uint32_t rows_per_part, uint32_t tile_cols, int n_threads,
                 TiledMatrix *out, std::string *err);

// A unit of CTA work: steps [t_lo, t_hi) (tile-relative) of one tile. A CTA stages the tile's x
// once per segment.
struct Segment {
    uint32_t tile, t_lo, t_hi;
    // copy of the tile's geometry, so that a CTA needs ONE dependent load before it can start
    // streaming: column range of the x tile, slice list, first step, and the slice-length table
    uint32_t col_base, col_count, slice_begin, n_slices, step_begin;
    uint32_t cnt_ge[32];
    // equal-cost shares of [t_lo, t_hi) for the CTA's 32 warps: warp w streams steps
    // [warp_t[w], warp_t[w+1]) and starts inside tile-relative slice warp_slice[w]
    uint32_t warp_t[33];
    uint32_t warp_slice[32];
    // In-CTA combining of row updates (wide layout): when the segment touches at most kCombineSlots slices, a slice
    // is usually split over several warps (the pruned transformer layers: 32-step slices, one or two per CTA), and
    // every warp would send its own 32 partial sums to the same few rows. The warps then add their partial sums
    // into a shared-memory table indexed by (slice - comb_first, lane) and the table is flushed once per segment.
    uint32_t comb_first;     // tile-relative index of the first slice the segment touches
    uint32_t comb_n;         // slices it touches if <= kCombineSlots, else 0 (no combining)
    uint32_t pad_[1];
};
constexpr uint32_t kCombineSlots = 48;
static_assert(sizeof(Segment) % 16 == 0, "Segment is loaded with 128-bit loads");
// Work plan for one launch over tiles [tile_begin, tile_end) on `ctas` CTAs: CTA b runs
// segs[cta_seg[b] .. cta_seg[b+1]). Cuts are placed at equal cost (steps + one unit per slice,
// the latter paying for the slice's 32 row updates) and may fall inside a slice; when there are
// no more tiles than CTAs no CTA works on two tiles (a second x staging would double its time).
// cost of finishing a slice in units of one step, as set by the last plan_launch of this thread
// (0.6 x the average steps per slice, within [2.5, 8]; HSB_SLICE_COST overrides)
double slice_cost();
void plan_launch(const TiledMatrix &m, uint32_t tile_begin, uint32_t tile_end, uint32_t ctas,
                 std::vector<uint32_t> *cta_seg, std::vector<Segment> *segs);

}  // namespace hsb
#endif
