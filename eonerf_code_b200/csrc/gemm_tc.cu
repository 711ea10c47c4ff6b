// placeholder until the tcgen05 kernels land
#include "gemm.cuh"
namespace eonerf {
int gemm_nt_tc(const GemmNT&, cudaStream_t) { set_error("tensor-core GEMM not built"); return EONERF_EINVAL; }
int gemm_tn_tc(const GemmTN&, cudaStream_t) { set_error("tensor-core GEMM not built"); return EONERF_EINVAL; }
}
