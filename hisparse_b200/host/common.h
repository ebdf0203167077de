// Compile-time configuration of the three implementations, mirroring the reference's
// spmv/libfpga/common.h:28-50,162-179 and spmv-fp/libfpga/common.h:17-62,175-199 on the host side
// (no HLS types). Select with -DFP_POB / -DFP_STALL exactly like sw/Makefile:2-12.
#ifndef HISPARSE_B200_HOST_COMMON_H_
#define HISPARSE_B200_HOST_COMMON_H_

#include <cstdint>
#include "fixed_point.h"
#include "../../include/hisparse_b200.h"

#define IDX_MARKER 0xffffffff

const unsigned PACK_SIZE = 8;
typedef unsigned IDX_T;

#if defined(FP_POB)
typedef float VAL_T;
const unsigned OB_BANK_SIZE = 1024;
const unsigned INTERLEAVE_FACTOR = 1;
const int HSB_IMPL = HSB_IMPL_FLOAT_POB;
#elif defined(FP_STALL)
typedef float VAL_T;
const unsigned OB_BANK_SIZE = 1024 * 8;
const unsigned INTERLEAVE_FACTOR = 8;
const int HSB_IMPL = HSB_IMPL_FLOAT_STALL;
#else
typedef spmv::ufixed_q8_24 VAL_T;
const unsigned OB_BANK_SIZE = 1024 * 8;
const unsigned INTERLEAVE_FACTOR = 1;
const int HSB_IMPL = HSB_IMPL_FIXED;
#endif
const unsigned VB_BANK_SIZE = 1024 * 4;

typedef struct { IDX_T data[PACK_SIZE]; } PACKED_IDX_T;
typedef struct { VAL_T data[PACK_SIZE]; } PACKED_VAL_T;
typedef struct {
    PACKED_IDX_T indices;
    PACKED_VAL_T vals;
} SPMV_MAT_PKT_T;
static_assert(sizeof(SPMV_MAT_PKT_T) == 64, "one HBM packet is 64 bytes");

const unsigned SK0_CLUSTER = 4;
const unsigned SK1_CLUSTER = 6;
const unsigned SK2_CLUSTER = 6;
const unsigned NUM_HBM_CHANNELS = SK0_CLUSTER + SK1_CLUSTER + SK2_CLUSTER;
const unsigned OB_PER_CLUSTER = OB_BANK_SIZE * PACK_SIZE;
const unsigned VB_PER_CLUSTER = VB_BANK_SIZE * PACK_SIZE;
const unsigned LOGICAL_OB_SIZE = NUM_HBM_CHANNELS * OB_PER_CLUSTER;
const unsigned LOGICAL_VB_SIZE = VB_PER_CLUSTER;

#endif
