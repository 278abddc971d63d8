// tcgen05 (REGT_PREC_TF32X3 / REGT_PREC_BF16) kernels of the cell -- placeholder until the
// tensor-core path lands; fails loudly rather than falling back.
#include "common.cuh"
namespace regt {
int cell_forward_tc(const regt_args*, const Layout&, cudaStream_t) {
  set_error("tensor-core precision modes are not built in this version; use REGT_PREC_FP32");
  return -10;
}
int cell_backward_tc(const regt_args*, const Layout&, cudaStream_t) {
  set_error("tensor-core precision modes are not built in this version; use REGT_PREC_FP32");
  return -10;
}
}  // namespace regt
