// Host-side declarations shared by the translation units of libvkrt_cuda (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/vkrt.h"

namespace vkrt {

struct DevScene;   // vkrt_device.cuh

// per-draw parameters: the "push constants" of the CUDA path (passed by value as kernel arguments,
// like vkCmdPushConstants at Source/GraphicsDevice.cpp:1264 -- no per-frame H2D copy is needed)
struct RenderParams {
    vkrt_frame_data fd;
    uint32_t width, height;
    uint32_t s_begin, s_end;      // sample range of this frame rendered by this context
    uint32_t max_depth;           // DEPTH (path) / BOUNCES (whitted)
    uint32_t fkey;                // frame key of the integer RNG
    uint32_t tile_rank, tile_count, tiles_x, tiles_y;
    uint32_t n_work;              // owned tiles * 1024 pixel slots
    uint32_t accumulate;
    float4 *accum;
    uint32_t *hit_ids;            // may be null
    unsigned long long *counters; // closest, shadow, nodes, leaves, paths
    uint32_t *work_head;
};

enum { TILE = 32, TILE_PX = TILE * TILE };
enum { CNT_CLOSEST = 0, CNT_SHADOW = 1, CNT_NODES = 2, CNT_LEAVES = 3, CNT_PATHS = 4, CNT_SKIPPED = 5, CNT_SHARED = 6, CNT_N = 8 };

struct LaunchCfg { int sm_count; };

// vkrt_render.cu
cudaError_t launch_path_mega(const DevScene &sc, const RenderParams &rp, bool bvh, bool stats, int sm_count,
                             cudaStream_t st);
cudaError_t launch_whitted(const DevScene &sc, const RenderParams &rp, bool bvh, bool stats, cudaStream_t st);
// where the resolve writes: dense / pitched linear memory (ptr, pitch in bytes) or a surface object (surf != 0)
struct ResolveTarget { uchar4 *ptr; size_t pitch; cudaSurfaceObject_t surf; };
cudaError_t launch_resolve(const RenderParams &rp, uint32_t integrator, const ResolveTarget &tg, cudaStream_t st);
cudaError_t launch_pack(const RenderParams &rp, float4 *packed, cudaStream_t st);
cudaError_t launch_unpack(float4 *accum, const float4 *packed, uint32_t width, uint32_t height, uint32_t tile_rank,
                          uint32_t tile_count, int add, cudaStream_t st);
cudaError_t launch_clear_accum(float4 *accum, size_t n, cudaStream_t st);
cudaError_t launch_present(const uchar4 *b0, const uchar4 *b1, uint32_t tw, uint32_t th, uchar4 *out, uint32_t W, uint32_t H,
                           cudaStream_t st);

// vkrt_wavefront.cu
struct WaveBuffers {
    size_t capacity;              // path records per wave
    float4 *rec;                  // 64-byte path records: {origin.xyz t_hit | dir.xyz hit id | acc.xyz pixel | mask.xyz sample<<8|depth}
    uint32_t *queue[2];           // active path indices (ping-pong over depth iterations)
    uint32_t *queue_mat[2];       // material bins of the shade stage: [0] dielectric, [1] diffuse
    uint32_t *counts;             // queue counters and work-fetch heads
    float4 *sample_rad;           // finished radiance per (pixel slot, sample in wave); reduced in sample order
    float4 *shadow;               // light-sample shadow rays {L.xyz, t_light}, n_lights per path
    float4 *term;                 // their unoccluded contributions (evaluated before the ray is queued)
    float4 *shrec;                // fused pipeline (one light): 64-byte shadow records {P.xyz t | L.xyz ended | acc_if_visible.xyz pixel | -}
    uint8_t *occ;                 // their any-hit results
    uint32_t *queue_shadow;       // (path << 4 | light) items that still need the sphere any-hit query
    uint32_t shadow_lights;
    // dense fused pipeline (scenes with <= 1 light): ping-pong ray / state arrays, nearest-hit results, shadow records
    float4 *d_ray[2], *d_st[2], *d_shr;
    float2 *d_hit;
    // VKRT_FLAG_LAUNCH_TIMING: CUDA-event pairs around every kernel launch of the last frame (roofline timing); the pool
    // grows with the number of launches, nothing is dropped
    std::vector<cudaEvent_t> ev;
    std::vector<uint8_t> ev_tag;  // kernel of pair k = (ev[2k], ev[2k+1]): 0 generate 1 extend 2 classify 3 shadow 4 shade 5 reduce
    uint32_t n_ev;                // events recorded by the last frame
};
cudaError_t wave_alloc(WaveBuffers &wb, size_t capacity);
void wave_free(WaveBuffers &wb);
// up to four lanes (buffer set + stream) whose waves overlap on the GPU
struct WaveEngine {
    WaveBuffers lane[4];
    cudaStream_t stream[4];
    cudaEvent_t ev_fork, ev_reduce[4];
    cudaEvent_t ev_stage[4];      // staggered lanes: "the wave on this lane has reached the stagger depth"
    uint32_t stage_lane; bool have_stage;
    float4 *frame_sum;            // running per-slot sum across the waves of a frame
    float2 *prim[2];              // per-slot primary hit {t, id} of the frame being launched / the frame before (frames overlap)
    uint32_t prim_flip;
    cudaEvent_t ev_prim;          // the first wave's generate kernel has stored the primary hits
    uint32_t n_lanes;
    uint32_t wave_seq;            // waves launched so far: lanes alternate across frame boundaries too
    uint32_t prev_reduce_lane; bool have_prev_reduce;
};
cudaError_t wave_engine_init(WaveEngine &eng, size_t lane_capacity, uint32_t n_lanes);
void wave_engine_free(WaveEngine &eng);
// ev_consumed: recorded by the caller on `st` once everything that still reads the accumulator / AOVs of the
// previous frame has been enqueued; only the kernels that overwrite those wait for it, so the bulk of a frame
// overlaps the tail (and the read-back / gather) of the frame before -- two frames in flight, like the
// reference's FRAMES_IN_FLIGHT (Source/Main.cpp:110).  *tail = the stream the frame's last kernel went to.
cudaError_t launch_path_wavefront(const DevScene &sc, const RenderParams &rp, WaveEngine &eng, bool bvh, bool stats, bool timing,
                                  int sm_count, cudaStream_t st, cudaEvent_t ev_consumed, cudaEvent_t ev_begin,
                                  cudaStream_t *tail, uint32_t *n_launches);

// vkrt_bvh.cu
struct BvhBuild {
    float4 *nodes = nullptr;      // binary nodes: 4 float4 (two child records) per inner node
    float4 *nodes4 = nullptr;     // 4-wide traversal nodes: 8 float4 (four child records) per binary node id
    float4 *tnodes = nullptr;     // exact nodes of the traversal tree (binned SAH, same leaves) the coded nodes are made from; null: `nodes`
    int tdepth = 0;               // its depth
    float sah_ms = 0.f;           // device time of its build (part of build_ms)
    uint4 *qnodes = nullptr;      // 32-byte traversal nodes: two 16-byte child records on the 16-bit grid below
    float qgrid[6] = {0, 0, 0, 0, 0, 0};   // per axis: scale s[3], offset b2[3]; coordinate of code q = (2^23 + q) * s + b2
    uint32_t n_nodes = 0;
    int depth = 0;                // levels of inner nodes on the longest root-to-leaf chain (<= 64, see k_tree_depth)
    float build_ms = 0.f;
    uint32_t launches = 0;
};
cudaError_t build_lbvh(const float4 *d_spheres, uint32_t n, BvhBuild &out, cudaStream_t st);
struct TriBvhBuild {
    float4 *nodes = nullptr;      // binary 64-byte nodes over the padded triangle boxes (rule T)
    uint32_t n_nodes = 0;
    int depth = 0;
    float build_ms = 0.f;
    uint32_t launches = 0;
};
cudaError_t build_tri_lbvh(const float4 *d_tris, uint32_t n, TriBvhBuild &out, cudaStream_t st);

// vkrt_exchange.cu: device-side synchronisation and collection of the multi-GPU frame exchange
enum { X_MAX_RANKS = 64, X_FLAG_STRIDE = 16 /* uint64s = 128 bytes */, X_OFF_CONSUMED = 8192, X_OFF_ERROR = 8320, X_OFF_TARGETS = 16384 };
cudaError_t launch_xwait(const unsigned long long *flags, uint32_t n, uint32_t stride, unsigned long long need, uint32_t *error, cudaStream_t st);
cudaError_t launch_xsignal(unsigned long long *flag, unsigned long long value, cudaStream_t st);
cudaError_t launch_xcollect(float4 *accum, const float4 *targets, size_t n_px, uint32_t S, int add, cudaStream_t st);

// vkrt_micro.cu
cudaError_t measure_fp32_peak(float *tflops);
cudaError_t measure_l2_bandwidth(float *gbs);

} // namespace vkrt
