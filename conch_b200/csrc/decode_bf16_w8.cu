// decode kernel instantiations: __nv_bfloat16 activations, 8-bit weights (see decode_impl.cuh / decode_kernel.cu)
#include "decode_impl.cuh"

namespace cb200 {
CB200_DECODE_ENTRY(launch_decode_bf16_w8) { return launch_typed<__nv_bfloat16, 8>(u, fused, p, geo, blocks, smem, wsp, wsc, stream); }
}  // namespace cb200
