// Multi-GPU frame exchange over peer memory (NVLink / NVSwitch), SURVEY.md 8(e).
//
// The reference is single-GPU (Source/VulkanState.h:52-55: "Single threaded ... Single GPU"); the one dispatch of
// Source/GraphicsDevice.cpp:1266 is split here by screen tiles (and sample ranges) over several GPUs.  Instead of
// pack -> gather -> unpack, every GPU's last kernel of a frame (k_wf_reduce, the megakernel, k_whitted) stores its
// owned pixels STRAIGHT into a frame target that lives in the gathering GPU's memory: RenderParams.accum is a peer
// pointer, the stores travel over NVLink while the kernel is still computing -- reduce and transfer are one kernel.
//
// Exchange block (one cudaMalloc on the gathering GPU, mapped by the others through CUDA IPC or plain peer access):
//     [0, 8192)       done[r]   one uint64 per rank, 128 bytes apart: frames rank r has completely written
//     [8192, 8320)    consumed  frames the gathering GPU has collected (their targets may be overwritten)
//     [8320, 8324)    error     set by a wait that timed out
//     [16384, ...)    2 * S frame targets (float4 per pixel): frame f of sample group s -> target[(f & 1) * S + s]
// Synchronisation is value based and entirely on the device: k_xsignal publishes a counter with system scope after a
// rank's frame (kernel boundaries order it behind the frame's stores), k_xwait spins on counters with acquire loads.
// No host synchronisation, no collective call in the data path; frames stay two deep in flight.
#include "vkrt_internal.h"

namespace vkrt {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// waits until flags[i * stride] >= need for every i < n; one thread per flag.  A wait that lasts longer than
// `timeout_ns` raises the block's error word and returns, so a dead peer cannot hang the GPU forever.
__global__ void k_xwait(const unsigned long long *flags, uint32_t n, uint32_t stride, unsigned long long need, uint32_t *error,
                        unsigned long long timeout_ns)
{
    const uint32_t i = threadIdx.x;
    if (i >= n) return;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flags + (size_t)i * stride) < need) {
        __nanosleep(100);
        if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) { atomicExch(error, 1u); return; }
    }
}
__global__ void k_xsignal(unsigned long long *flag, unsigned long long value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(value) : "memory");
}

// gathering GPU: accumulator = (add ? accumulator : 0) + target[0] + target[1] + ... in sample-group order
// (one group: a plain copy -- the same float operations as a single GPU's k_wf_reduce / progressive add)
__global__ void __launch_bounds__(256) k_xcollect(float4 *__restrict__ accum, const float4 *__restrict__ targets, size_t n_px, uint32_t S, int add)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    float4 a = __ldcg(targets + i);
    for (uint32_t s = 1; s < S; ++s) {
        const float4 b = __ldcg(targets + (size_t)s * n_px + i);
        a.x = a.x + b.x; a.y = a.y + b.y; a.z = a.z + b.z; a.w = a.w + b.w;
    }
    if (add) { const float4 o = accum[i]; a.x = o.x + a.x; a.y = o.y + a.y; a.z = o.z + a.z; a.w = o.w + a.w; }
    accum[i] = a;
}

cudaError_t launch_xwait(const unsigned long long *flags, uint32_t n, uint32_t stride, unsigned long long need, uint32_t *error, cudaStream_t st)
{
    k_xwait<<<1, 64, 0, st>>>(flags, n, stride, need, error, 60ull * 1000000000ull);
    return cudaGetLastError();
}
cudaError_t launch_xsignal(unsigned long long *flag, unsigned long long value, cudaStream_t st)
{
    k_xsignal<<<1, 1, 0, st>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_xcollect(float4 *accum, const float4 *targets, size_t n_px, uint32_t S, int add, cudaStream_t st)
{
    k_xcollect<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(accum, targets, n_px, S, add);
    return cudaGetLastError();
}

} // namespace vkrt
