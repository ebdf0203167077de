// Link-compatible stand-in for the reference's C-simulation entry point `top_wrapper`
// (spmv_csim/csim.cpp:22-46): the SAME name and the SAME 23-argument signature over the reference's own
// packet types, so that a harness written against csim.cpp -- csim.cpp's own spmv_test_harness
// (:203-381) included -- runs on the B200 by swapping one object at link time.
//
// Include it AFTER the reference's common.h (spmv/libfpga/common.h or spmv-fp/libfpga/common.h): it uses
// that header's SPMV_MAT_PKT_T / PACKED_VAL_T and its FP_POB / FP_STALL switches (spmv_csim/Makefile:27-38).
// A translation unit that defines HSB_TOP_WRAPPER_DEFINE before the include also gets the definition, which
// forwards to hsb_top_wrapper (include/hisparse_b200.h); errors are reported the reference's way: message on
// stderr, exit(EXIT_FAILURE) (xrt/includes/xcl2/xcl2.hpp:40-46). oracle/Makefile shows the link recipe
// (the reference's own definition in csim.o is weakened with objcopy, this one takes its place).
#ifndef HISPARSE_B200_HOST_TOP_WRAPPER_H_
#define HISPARSE_B200_HOST_TOP_WRAPPER_H_

#include "../../include/hisparse_b200.h"

void top_wrapper(const SPMV_MAT_PKT_T *matrix_hbm_0, const SPMV_MAT_PKT_T *matrix_hbm_1,
                 const SPMV_MAT_PKT_T *matrix_hbm_2, const SPMV_MAT_PKT_T *matrix_hbm_3,
                 const SPMV_MAT_PKT_T *matrix_hbm_4, const SPMV_MAT_PKT_T *matrix_hbm_5,
                 const SPMV_MAT_PKT_T *matrix_hbm_6, const SPMV_MAT_PKT_T *matrix_hbm_7,
                 const SPMV_MAT_PKT_T *matrix_hbm_8, const SPMV_MAT_PKT_T *matrix_hbm_9,
                 const SPMV_MAT_PKT_T *matrix_hbm_10, const SPMV_MAT_PKT_T *matrix_hbm_11,
                 const SPMV_MAT_PKT_T *matrix_hbm_12, const SPMV_MAT_PKT_T *matrix_hbm_13,
                 const SPMV_MAT_PKT_T *matrix_hbm_14, const SPMV_MAT_PKT_T *matrix_hbm_15,
                 const PACKED_VAL_T *packed_dense_vector, PACKED_VAL_T *packed_dense_result,
                 const unsigned row_part_id, const unsigned part_len, const unsigned num_col_partitions,
                 const unsigned num_partitions, const unsigned num_cols);

#ifdef HSB_TOP_WRAPPER_DEFINE
#include <cstdio>
#include <cstdlib>

static_assert(sizeof(SPMV_MAT_PKT_T) == 64, "a matrix packet is 8 column ids + 8 value words (common.h:44-50)");
static_assert(sizeof(PACKED_VAL_T) == 32, "a vector packet is 8 value words");

void top_wrapper(const SPMV_MAT_PKT_T *matrix_hbm_0, const SPMV_MAT_PKT_T *matrix_hbm_1,
                 const SPMV_MAT_PKT_T *matrix_hbm_2, const SPMV_MAT_PKT_T *matrix_hbm_3,
                 const SPMV_MAT_PKT_T *matrix_hbm_4, const SPMV_MAT_PKT_T *matrix_hbm_5,
                 const SPMV_MAT_PKT_T *matrix_hbm_6, const SPMV_MAT_PKT_T *matrix_hbm_7,
                 const SPMV_MAT_PKT_T *matrix_hbm_8, const SPMV_MAT_PKT_T *matrix_hbm_9,
                 const SPMV_MAT_PKT_T *matrix_hbm_10, const SPMV_MAT_PKT_T *matrix_hbm_11,
                 const SPMV_MAT_PKT_T *matrix_hbm_12, const SPMV_MAT_PKT_T *matrix_hbm_13,
                 const SPMV_MAT_PKT_T *matrix_hbm_14, const SPMV_MAT_PKT_T *matrix_hbm_15,
                 const PACKED_VAL_T *packed_dense_vector, PACKED_VAL_T *packed_dense_result,
                 const unsigned row_part_id, const unsigned part_len, const unsigned num_col_partitions,
                 const unsigned num_partitions, const unsigned num_cols) {
#if defined(FP_POB)
    const int impl = HSB_IMPL_FLOAT_POB;
#elif defined(FP_STALL)
    const int impl = HSB_IMPL_FLOAT_STALL;
#else
    const int impl = HSB_IMPL_FIXED;
#endif
    const void *const ch[HSB_NUM_HBM_CHANNELS] = {
        matrix_hbm_0, matrix_hbm_1, matrix_hbm_2,  matrix_hbm_3,  matrix_hbm_4,  matrix_hbm_5,  matrix_hbm_6,  matrix_hbm_7,
        matrix_hbm_8, matrix_hbm_9, matrix_hbm_10, matrix_hbm_11, matrix_hbm_12, matrix_hbm_13, matrix_hbm_14, matrix_hbm_15};
    const int rc = hsb_top_wrapper(impl, ch, packed_dense_vector, packed_dense_result, row_part_id, part_len,
                                   num_col_partitions, num_partitions, num_cols);
    if (rc != HSB_OK) {
        std::fprintf(stderr, "%s:%d top_wrapper on the GPU failed (%d): %s\n", __FILE__, __LINE__, rc, hsb_last_error());
        std::exit(EXIT_FAILURE);
    }
}
#endif  // HSB_TOP_WRAPPER_DEFINE
#endif
