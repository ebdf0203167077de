// TEST INFRASTRUCTURE ONLY (oracle). The reference includes <ap_axi_sdata.h>
// (spmv/libfpga/common.h:6) but uses nothing from it in C simulation.
#ifndef HISPARSE_ORACLE_SHIM_AP_AXI_SDATA_H_
#define HISPARSE_ORACLE_SHIM_AP_AXI_SDATA_H_
#endif
