// trace_f32.cu -- fp32 geometry mode (placeholder until the fp32 kernel lands; the API refuses the mode).
#include "rtb_device.cuh"

namespace rtb {
cudaError_t launch_trace_f32(const TraceParams &, int, cudaStream_t) { return cudaErrorNotSupported; }
} // namespace rtb
