// streaming decode kernel instantiations: __nv_bfloat16 activations, 8-bit weights as two nibble planes (see stream_impl.cuh / stream_kernel.cu)
#include "stream_impl.cuh"

namespace cb200 {
CB200_STREAM_ENTRY(launch_stream_bf16_w8) { return launch_stream_typed<__nv_bfloat16,8>(mw, p, geo, grid, smem, pdl, stream); }
}  // namespace cb200
