// tcgen05 streaming decode kernel instantiations: __nv_bfloat16 activations, 4-bit weights (see tcs_impl.cuh / tcs_kernel.cu)
#include "tcs_impl.cuh"

namespace cb200 {
CB200_TCS_ENTRY(launch_tcs_bf16_w4) { return launch_tcs_typed<__nv_bfloat16, 4>(mt, mw, mx, p, geo, ctas, smem, wsp, wsc, stream); }
}  // namespace cb200
