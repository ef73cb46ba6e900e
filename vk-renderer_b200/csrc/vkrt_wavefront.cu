// placeholder, replaced below
#include "vkrt_device.cuh"
#include "vkrt_internal.h"
namespace vkrt {
cudaError_t wave_alloc(WaveBuffers &, size_t) { return cudaErrorNotSupported; }
void wave_free(WaveBuffers &) {}
cudaError_t launch_path_wavefront(const DevScene &, const RenderParams &, WaveBuffers &, bool, bool, int, cudaStream_t, uint32_t *) { return cudaErrorNotSupported; }
}
