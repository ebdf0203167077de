// The object that replaces the reference's own `top_wrapper` when spmv_csim/csim.cpp is linked against
// libhisparse_b200.so (oracle/Makefile, targets csim_gpu_*). "common.h" is the reference's, found through the
// include path exactly as csim.cpp finds it.
#include "common.h"
#define HSB_TOP_WRAPPER_DEFINE
#include "top_wrapper.h"
