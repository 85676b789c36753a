// filter_fast.cu -- fast tier (placeholder until the sm_100a kernels land)
#include "common.cuh"
namespace oxi {
bool fast_supported(const Geom &, const StrategySet &) { return false; }
cudaError_t launch_fast(const LaunchCtx &, const Geom &, const uint8_t *, size_t, const StrategySet &) {
    return cudaErrorNotSupported;
}
} // namespace oxi
