// decode kernel instantiations: __half activations, 4-bit weights (see decode_impl.cuh / decode_kernel.cu)
#include "decode_impl.cuh"

namespace cb200 {
CB200_DECODE_ENTRY(launch_decode_f16_w4) { return launch_typed<__half, 4>(u, fused, p, geo, blocks, smem, wsp, wsc, stream); }
}  // namespace cb200
