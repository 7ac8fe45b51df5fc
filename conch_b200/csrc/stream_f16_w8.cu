// streaming decode kernel instantiations: __half activations, 8-bit weights (see stream_impl.cuh / stream_kernel.cu)
#include "stream_impl.cuh"

namespace cb200 {
CB200_STREAM_ENTRY(launch_stream_f16_w8) { return launch_stream_typed<__half,8>(mw, p, geo, grid, smem, pdl, stream); }
}  // namespace cb200
