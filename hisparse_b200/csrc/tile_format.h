// B200 tile-stream matrix format ("TS-CPSR") -- the HBM layout the sm_100a SpMV kernels read.
//
// It plays the role of the reference's CPSR channel images (sw/data_formatter.h:194-238,
// sw/host.cpp:163-231) but is laid out for 32-wide warps and 128-byte coalescing instead of
// 8-lane 64-byte HBM packets:
//
//   * rows are cut into row partitions (the reference's LOGICAL_OB_SIZE cut,
//     sw/data_formatter.h:494), columns into tiles of <= 32768 columns (the reference's
//     LOGICAL_VB_SIZE = 32768-word vector buffer, spmv/libfpga/common.h:165,179) so that the
//     x tile of a work unit fits in one CTA's shared memory and a local column id fits 15 bits;
//   * inside a (row partition, column tile) the non-zeros are kept in CSR order as one stream;
//     the last non-zero of every (row, tile) segment carries an end-of-segment flag in bit 15
//     of its 16-bit column word -- the in-band analogue of the reference's end-of-row marker
//     (IDX_MARKER entries, sw/data_formatter.h:51-171) at zero extra bytes;
//   * the stream is cut into chunks of 256 non-zeros = 32 lanes x 8 consecutive non-zeros;
//     values are stored so that each of a warp's two 128-bit loads is one contiguous 512 B;
//   * seg_row[s] is the matrix row of the s-th segment (the row the reference recovers by
//     counting markers, spmv/libfpga/spmv_cluster.h:78-83).
//
// 6 bytes per non-zero + 4 bytes per non-empty (row, tile) segment + 8 bytes per 256 non-zeros.
#ifndef HISPARSE_B200_TILE_FORMAT_H_
#define HISPARSE_B200_TILE_FORMAT_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace hsb {

constexpr int kLanes = 32;
constexpr int kNnzPerLane = 8;
constexpr int kChunkNnz = kLanes * kNnzPerLane;          // 256
constexpr uint32_t kMaxTileCols = 32768;                 // 15-bit local column id, 128 KB of x
constexpr uint16_t kSegEndFlag = 0x8000;
constexpr uint32_t kChunkContinues = 0x80000000u;        // ChunkDesc.tile bit: real non-zeros follow the last flag

struct ChunkDesc {
    uint32_t seg_base;      // global index (into seg_row) of the segment open at the start of the chunk
    uint32_t tile;          // tile index | kChunkContinues
};

struct TileDesc {
    uint32_t col_base;      // first column of the tile (multiple of 8)
    uint32_t col_count;     // columns in the tile, rounded up to a multiple of 8
    uint32_t chunk_begin;   // [chunk_begin, chunk_end) in the global chunk stream
    uint32_t chunk_end;
    uint32_t row_part;      // row partition the tile belongs to
    uint32_t pad_[3];
};

struct TiledMatrix {
    uint32_t rows = 0, cols = 0;            // as given (rows may include padding rows)
    uint64_t nnz = 0;
    uint32_t rows_per_part = 0, n_row_parts = 0, n_col_tiles = 0, tile_cols = 0;
    std::vector<uint32_t> vals;             // n_chunks * 256 words, warp-transposed inside a chunk
    std::vector<uint16_t> cidx;             // n_chunks * 256 : local column | end-of-segment flag
    std::vector<ChunkDesc> chunks;          // n_chunks
    std::vector<uint32_t> seg_row;          // n_segments
    std::vector<TileDesc> tiles;            // n_row_parts * n_col_tiles, row-partition major
    std::vector<uint32_t> part_chunk_begin; // n_row_parts + 1
    size_t n_chunks() const { return chunks.size(); }
    size_t format_bytes() const {
        return vals.size() * 4 + cidx.size() * 2 + chunks.size() * sizeof(ChunkDesc) + seg_row.size() * 4 +
               tiles.size() * sizeof(TileDesc);
    }
};

// position of the k-th (0..7) non-zero of lane l inside a chunk's value block
inline size_t val_slot(int lane, int k) { return (size_t)((k >> 2) * kLanes + lane) * 4 + (k & 3); }

// Choose a tile width: the fewest tiles of <= kMaxTileCols columns, widths equalised, multiple of 8.
uint32_t choose_tile_cols(uint32_t cols);

// CSR (32-bit value words, passed through untouched) -> tile streams. rows_per_part == 0 means
// one row partition. Returns false and fills *err on malformed input.
bool build_tiled(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                 const uint32_t *vals, uint32_t rows_per_part, uint32_t tile_cols, int n_threads,
                 TiledMatrix *out, std::string *err);

}  // namespace hsb
#endif
