// TEST AID, not part of libhisparse_b200.so: the contract between the formatter, the planner and the kernel,
// checked without a GPU (tests/test_host_logic.py). Built by hisparse_b200/host/Makefile into
// host/bin/libhsb_plancheck.so from this file + csrc/tile_format.cpp.
#include "../../include/hisparse_b200.h"

#include <algorithm>
#include <string>
#include <vector>

#include "../csrc/tile_format.h"

// Host-side walk of a whole-matrix launch, step for step the way spmv_tiles_kernel does it: every CTA takes
// its segments from the plan, every warp its share [warp_t[w], warp_t[w+1]) starting inside slice
// warp_slice[w], slice lengths come from the tile's cnt_ge table, lanes accumulate and hand their sums to
// slice_rows at every slice end (and at the end of a share that stops inside a slice). Fixed-point
// arithmetic (exact 64-bit sums of the rounded, saturated products, one clamp at the end). It exists so that
// the contract between the planner / formatter and the kernel can be checked without a GPU.
static int walk_plan_fixed(const hsb::TiledMatrix &M, uint32_t ctas, const uint32_t *x_words, uint32_t *y_words) {
    if (!x_words || !y_words || ctas == 0) return HSB_EINVAL;
    std::vector<uint32_t> cta_seg;
    std::vector<hsb::Segment> segs;
    hsb::plan_launch(M, 0, (uint32_t)M.tiles.size(), ctas, &cta_seg, &segs);
    std::vector<unsigned long long> acc((size_t)M.rows + 1, 0ull);
    auto steps_before = [](const uint32_t *cnt, uint32_t i) { uint32_t s = 0; for (int c = 0; c < 32; c++) s += std::min(i, cnt[c]); return s; };
    auto steps_of = [](const uint32_t *cnt, uint32_t i) { uint32_t s = 0; for (int c = 0; c < 32; c++) s += cnt[c] > i ? 1u : 0u; return s; };
    for (uint32_t b = 0; b < ctas; b++)
        for (uint32_t g = cta_seg[b]; g < cta_seg[b + 1]; g++) {
            const hsb::Segment &sg = segs[g];
            for (int w = 0; w < hsb::kWarpsPerCta; w++) {
                const uint32_t ta = sg.warp_t[w], tb = sg.warp_t[w + 1];
                if (tb < ta || ta < sg.t_lo || tb > sg.t_hi) return HSB_EINVAL;
                uint32_t remaining = tb - ta;
                if (!remaining) continue;
                uint32_t sl = sg.warp_slice[w];
                if (sl >= sg.n_slices) return HSB_EINVAL;
                if (M.narrow) {
                    // narrow layout (stream_units_narrow in the kernel): units of 32 elements; a share starts at the
                    // row unit of slice sl and ends at a slice boundary; L(sl) step units follow every row unit
                    if (sl + steps_before(sg.cnt_ge, sl) != ta) return HSB_EINVAL;          // not at a slice boundary
                    const size_t ubase = (size_t)(sg.step_begin + ta) * hsb::kUnitElems;
                    uint32_t nleft = 0, rows_of[hsb::kLanes] = {};
                    unsigned long long lacc[hsb::kLanes] = {};
                    for (uint32_t k = 0; k < remaining; k++) {
                        const size_t ub = ubase + (size_t)k * hsb::kUnitElems;
                        if (ub + hsb::kLanes > M.vals.size()) return HSB_EINVAL;
                        if (nleft == 0) {                                                   // a row unit opens slice sl
                            if (sl >= sg.n_slices) return HSB_EINVAL;
                            for (int l = 0; l < hsb::kLanes; l++) {
                                if (M.cols16[ub + l] != hsb::kPadCol || M.vals[ub + l] > M.rows) return HSB_EINVAL;
                                rows_of[l] = M.vals[ub + l];
                            }
                            nleft = steps_of(sg.cnt_ge, sl);
                            if (nleft == 0) return HSB_EINVAL;
                            continue;
                        }
                        for (int l = 0; l < hsb::kLanes; l++) {
                            const uint32_t id = M.cols16[ub + l];
                            uint32_t xv = 0;
                            if (id >= hsb::kColBias) {
                                const uint32_t col = sg.col_base + id - hsb::kColBias;
                                if (id - hsb::kColBias >= sg.col_count || col >= M.cols) return HSB_EINVAL;
                                xv = x_words[col];
                            } else if (id != hsb::kPadCol || M.vals[ub + l] != 0) {
                                return HSB_EINVAL;
                            }
                            unsigned long long q = ((unsigned long long)M.vals[ub + l] * xv + 0x800000ull) >> 24;
                            lacc[l] += std::min<unsigned long long>(q, 0xFFFFFFFFull);
                        }
                        if (--nleft == 0) {
                            for (int m = 0; m < hsb::kLanes; m++) { acc[rows_of[m]] += lacc[m]; lacc[m] = 0; }
                            sl++;
                        }
                    }
                    if (nleft != 0) return HSB_EINVAL;                                       // the share ended inside a slice
                    continue;
                }
                if (steps_before(sg.cnt_ge, sl) > ta || steps_before(sg.cnt_ge, sl + 1) <= ta) return HSB_EINVAL;   // wrong first slice
                uint32_t left = steps_before(sg.cnt_ge, sl + 1) - ta;
                const size_t base = (size_t)(sg.step_begin + ta) * hsb::kStepElems;
                unsigned long long lane_acc[hsb::kLanes] = {};
                auto flush = [&]() {
                    // (the kernel's shared-memory combining table covers slices [comb_first, comb_first + comb_n))
                    if (sg.comb_n && (sl < sg.comb_first || sl - sg.comb_first >= sg.comb_n)) return false;
                    for (int l = 0; l < hsb::kLanes; l++) {
                        const uint32_t row = M.slice_rows[(size_t)(sg.slice_begin + sl) * hsb::kLanes + l];
                        if (row > M.rows) return false;
                        acc[row] += lane_acc[l];
                        lane_acc[l] = 0;
                    }
                    return true;
                };
                for (uint32_t k = 0; k < remaining; k++) {
                    for (int l = 0; l < hsb::kLanes; l++)
                        for (int j = 0; j < hsb::kSlotBlock; j++) {
                            const size_t e = base + (size_t)k * hsb::kStepElems + (size_t)l * hsb::kSlotBlock + j;
                            if (e >= M.vals.size()) return HSB_EINVAL;
                            const uint32_t id = M.cols16[e];
                            uint32_t xv = 0;
                            if (id >= hsb::kColBias) {
                                const uint32_t col = sg.col_base + id - hsb::kColBias;
                                if (id - hsb::kColBias >= sg.col_count || col >= M.cols) return HSB_EINVAL;
                                xv = x_words[col];
                            } else if (id != hsb::kPadCol || M.vals[e] != 0) {
                                return HSB_EINVAL;            // ids 1..7 are never stored; padding carries value 0
                            }
                            unsigned long long q = ((unsigned long long)M.vals[e] * xv + 0x800000ull) >> 24;
                            lane_acc[l] += std::min<unsigned long long>(q, 0xFFFFFFFFull);
                        }
                    if (--left == 0) {
                        if (!flush()) return HSB_EINVAL;
                        sl++;
                        left = steps_of(sg.cnt_ge, sl);
                    }
                }
                if (left != steps_of(sg.cnt_ge, sl) && sl < sg.n_slices && !flush()) return HSB_EINVAL;
            }
        }
    for (uint32_t r = 0; r < M.rows; r++) y_words[r] = acc[r] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)acc[r];
    return HSB_OK;
}


extern "C" int hsbt_emulate_fixed(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                                  const uint32_t *vals, uint32_t rows_per_partition, uint32_t tile_cols, uint32_t ctas,
                                  const uint32_t *x_words, uint32_t *y_words) {
    hsb::TiledMatrix M;
    std::string err;
    if (!hsb::build_tiled(rows, cols, indptr, indices, vals, rows_per_partition,
                          tile_cols ? tile_cols : hsb::choose_tile_cols(cols, rows, rows ? indptr[rows] : 0), 0, &M, &err))
        return HSB_EINVAL;
    return walk_plan_fixed(M, ctas, x_words, y_words);
}
