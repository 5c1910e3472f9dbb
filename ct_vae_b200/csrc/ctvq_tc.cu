// tcgen05/TMEM distance GEMM path (placeholder until the kernel lands): reports "unsupported" so the
// dispatcher uses the SIMT kernel.
#include "ctvq_common.cuh"

namespace ctvq {
bool tc_supported(const QuantParams&) { return false; }
int launch_forward_tc(const QuantParams&, cudaStream_t) { return CTVQ_E_UNSUPPORTED; }
}  // namespace ctvq
