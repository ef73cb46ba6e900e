// Render kernels of libvkrt_cuda: the persistent-thread path megakernel, the whitted kernel,
// the resolve (tone-map) kernel and the shard pack/unpack kernels.
//
// This replaces vkCmdDispatch(RES/16, RES/16, 1) of Source/GraphicsDevice.cpp:1266, i.e. the
// execution of Assets/Tracer.comp::main (:557-593) / Assets/Raytracer.comp::main (:357-399).
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

namespace vkrt {

// work slot -> pixel.  Slots are laid out tile by tile (32x32 px, this context owns tiles
// rank, rank+count, ...); inside a tile 32 consecutive slots form an 8x4 pixel block, so the
// first fetch of a warp is a coherent packet of primary rays.
VKRT_DEV bool slot_to_pixel(const RenderParams &rp, uint32_t w, uint32_t &px, uint32_t &py)
{
    const uint32_t lt = w >> 10, in = w & 1023u;
    const uint32_t gt = rp.tile_rank + lt * rp.tile_count;
    const uint32_t tx = gt % rp.tiles_x, ty = gt / rp.tiles_x;
    const uint32_t blk = in >> 5, l = in & 31u;
    px = tx * TILE + (blk & 3u) * 8u + (l & 7u);
    py = ty * TILE + (blk >> 2) * 4u + (l >> 3);
    return px < rp.width && py < rp.height;
}

VKRT_DEV void flush_stats(const Stats &st, unsigned long long *counters, bool stats)
{
    const unsigned full = 0xffffffffu;
    const uint32_t c = __reduce_add_sync(full, st.closest), s = __reduce_add_sync(full, st.shadow),
                   p = __reduce_add_sync(full, st.paths), k = __reduce_add_sync(full, st.skipped),
                   h = __reduce_add_sync(full, st.shared);
    uint32_t n = 0, l = 0;
    if (stats) { n = __reduce_add_sync(full, st.nodes); l = __reduce_add_sync(full, st.leaves); }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counters + CNT_CLOSEST, (unsigned long long)c);
        atomicAdd(counters + CNT_SHADOW, (unsigned long long)s);
        atomicAdd(counters + CNT_PATHS, (unsigned long long)p);
        if (k) atomicAdd(counters + CNT_SKIPPED, (unsigned long long)k);
        if (h) atomicAdd(counters + CNT_SHARED, (unsigned long long)h);
        if (stats) { atomicAdd(counters + CNT_NODES, (unsigned long long)n); atomicAdd(counters + CNT_LEAVES, (unsigned long long)l); }
    }
}

// ------------------------------------------------------------------------------------------------
// Path megakernel.  Persistent warps; every LANE owns one pixel at a time and walks its samples
// in order (so the per-pixel sum is formed in sample order, bit-identical to the oracle).  A lane
// whose path ends regenerates the next sample of its pixel at once, and a lane whose pixel is done
// fetches a new pixel slot: the lanes that need work are found with __ballot_sync, one lane does a
// single atomicAdd for all of them and the base is broadcast with __shfl_sync.
// ------------------------------------------------------------------------------------------------
#ifndef VKRT_MEGA_MINBLOCKS
#define VKRT_MEGA_MINBLOCKS 8      // 64 registers: 1/6/8/10/12 measured on the default scene (10.27 / 9.67 / 9.47 / 9.94 / 10.29 ms)
#endif
template <bool BVH, bool STATS>
__global__ void __launch_bounds__(128, VKRT_MEGA_MINBLOCKS) k_path_mega(const __grid_constant__ DevScene sc,
                                                    const __grid_constant__ RenderParams rp)
{
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    Stats st; stats_zero(st);
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);

    int pix = -1;
    bool exhausted = false;
    uint32_t s = 0, skey = 0;
    V3 po = v3(0.f), pd = v3(0.f), sum = v3(0.f);
    PathState ps; path_begin(ps, po, pd);

    for (;;) {
        const bool need = pix < 0 && !exhausted;
        const unsigned m = __ballot_sync(full, need);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(rp.work_head, (uint32_t)__popc(m));
            base = __shfl_sync(full, base, leader);
            if (need) {
                const uint32_t w = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
                if (w >= rp.n_work) exhausted = true;
                else {
                    uint32_t px, py;
                    if (slot_to_pixel(rp, w, px, py)) {
                        pix = (int)(py * rp.width + px);
                        primary_ray(rp.fd, rp.width, rp.height, px, py, po, pd);
                        s = rp.s_begin; sum = v3(0.0f);
                        skey = sample_key(rp.fkey, (uint32_t)pix, s);
                        path_begin(ps, po, pd);
                        ++st.paths;
                    }
                }
            }
        }
        if (__all_sync(full, pix < 0 && exhausted)) break;
        if (pix >= 0) {
            uint32_t pid;
            const bool want_id = rp.hit_ids != nullptr && ps.depth == 0 && s == rp.s_begin;
            const bool alive = path_bounce<BVH, STATS>(sc, cam_pos, rp.max_depth, skey, ps, st, want_id ? &pid : nullptr);
            if (want_id) rp.hit_ids[pix] = pid;
            if (!alive) {
                sum = sum + ps.acc;                                       // Tracer.comp:580
                ++s;
                if (s == rp.s_end) {
                    const float n = (float)(rp.s_end - rp.s_begin);
                    float4 a = make_float4(sum.x, sum.y, sum.z, n);
                    if (rp.accumulate) { const float4 o = rp.accum[pix]; a.x = o.x + a.x; a.y = o.y + a.y; a.z = o.z + a.z; a.w = o.w + a.w; }
                    rp.accum[pix] = a;
                    pix = -1;
                } else {
                    skey = sample_key(rp.fkey, (uint32_t)pix, s);
                    path_begin(ps, po, pd);
                    ++st.paths;
                }
            }
        }
    }
    flush_stats(st, rp.counters, STATS);
}

cudaError_t launch_path_mega(const DevScene &sc, const RenderParams &rp, bool bvh, bool stats, int sm_count,
                             cudaStream_t stream)
{
    void (*k)(const DevScene, const RenderParams) =
        bvh ? (stats ? k_path_mega<true, true> : k_path_mega<true, false>)
            : (stats ? k_path_mega<false, true> : k_path_mega<false, false>);
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const unsigned need_warps = (rp.n_work + 31u) / 32u;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned max_useful = (need_warps + 3u) / 4u;
    if (grid > max_useful) grid = max_useful;
    if (grid == 0) grid = 1;
    k<<<grid, 128, 0, stream>>>(sc, rp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Whitted kernel (Raytracer.comp): deterministic, <= 6 rays per pixel; one thread per pixel slot.
// ------------------------------------------------------------------------------------------------
template <bool BVH, bool STATS>
__global__ void __launch_bounds__(256) k_whitted(const __grid_constant__ DevScene sc,
                                                  const __grid_constant__ RenderParams rp)
{
    Stats st; stats_zero(st);
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px, py;
    if (w < rp.n_work && slot_to_pixel(rp, w, px, py)) {
        V3 o, d;
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        uint32_t pid;
        const V3 c = whitted_pixel<BVH, STATS>(sc, rp.fd, o, d, rp.max_depth, st, &pid);
        const uint32_t pix = py * rp.width + px;
        rp.accum[pix] = make_float4(c.x, c.y, c.z, 1.0f);
        if (rp.hit_ids) rp.hit_ids[pix] = pid;
    }
    flush_stats(st, rp.counters, STATS);
}

cudaError_t launch_whitted(const DevScene &sc, const RenderParams &rp, bool bvh, bool stats, cudaStream_t stream)
{
    void (*k)(const DevScene, const RenderParams) =
        bvh ? (stats ? k_whitted<true, true> : k_whitted<true, false>)
            : (stats ? k_whitted<false, true> : k_whitted<false, false>);
    const unsigned grid = (rp.n_work + 255u) / 256u;
    k<<<grid ? grid : 1, 256, 0, stream>>>(sc, rp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Resolve: Tracer.comp:585-592 (mean, Reinhard, gamma 1/2.2, +rand()/64 dither, unorm8 store) or
// Raytracer.comp:398 (plain unorm8 store).
// ------------------------------------------------------------------------------------------------
// TARGET 0: the library's dense image; 1: linear memory with a row pitch (a bound / imported buffer or
// VK_IMAGE_TILING_LINEAR image); 2: a CUDA array behind a surface object (an imported VK_IMAGE_TILING_OPTIMAL image,
// ref: Source/GraphicsDevice.cpp:672-673).  Texel (x, y) = imageStore(ivec2(gid.xy)) of Tracer.comp:592.
template <int TARGET>
__global__ void __launch_bounds__(256) k_resolve(const __grid_constant__ RenderParams rp, uint32_t integrator,
                                                  const ResolveTarget tg)
{
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= rp.width * rp.height) return;
    const float4 a = rp.accum[pix];
    uchar4 o;
    if (integrator == VKRT_INTEGRATOR_WHITTED) {
        o = make_uchar4(unorm8(a.x), unorm8(a.y), unorm8(a.z), 255);
    } else {
        V3 c = v3(a.x, a.y, a.z) / a.w;
        c = c / (c + v3(1.0f));
        c = v3(pow_(c.x, 1.0f / 2.2f), pow_(c.y, 1.0f / 2.2f), pow_(c.z, 1.0f / 2.2f));
        const float dither = u01(sample_key(rp.fkey, pix, VKRT_DITHER_SAMPLE), 0) / 64.0f;
        o = make_uchar4(unorm8(c.x + dither), unorm8(c.y + dither), unorm8(c.z + dither), 255);
    }
    if (TARGET == 0) tg.ptr[pix] = o;
    else {
        const uint32_t y = pix / rp.width, x = pix - y * rp.width;
        if (TARGET == 1) *reinterpret_cast<uchar4 *>(reinterpret_cast<char *>(tg.ptr) + (size_t)y * tg.pitch + (size_t)x * 4u) = o;
        else surf2Dwrite(o, tg.surf, (int)(x * 4u), (int)y);
    }
}

cudaError_t launch_resolve(const RenderParams &rp, uint32_t integrator, const ResolveTarget &tg, cudaStream_t stream)
{
    const unsigned n = rp.width * rp.height;
    const unsigned grid = (n + 255u) / 256u;
    if (tg.surf) k_resolve<2><<<grid, 256, 0, stream>>>(rp, integrator, tg);
    else if (tg.pitch != (size_t)rp.width * 4u) k_resolve<1><<<grid, 256, 0, stream>>>(rp, integrator, tg);
    else k_resolve<0><<<grid, 256, 0, stream>>>(rp, integrator, tg);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Shard exchange helpers: owned tiles <-> compact buffer (n_owned_tiles * 1024 float4).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const __grid_constant__ RenderParams rp, float4 *__restrict__ packed)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= rp.n_work) return;
    uint32_t px, py;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot_to_pixel(rp, w, px, py)) v = rp.accum[py * rp.width + px];
    packed[w] = v;
}
cudaError_t launch_pack(const RenderParams &rp, float4 *packed, cudaStream_t stream)
{
    k_pack<<<(rp.n_work + 255u) / 256u, 256, 0, stream>>>(rp, packed);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_unpack(float4 *__restrict__ accum, const float4 *__restrict__ packed,
                                                 RenderParams rp, int add)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= rp.n_work) return;
    uint32_t px, py;
    if (!slot_to_pixel(rp, w, px, py)) return;
    float4 v = packed[w];
    const uint32_t pix = py * rp.width + px;
    if (add) { const float4 o = accum[pix]; v.x = o.x + v.x; v.y = o.y + v.y; v.z = o.z + v.z; v.w = o.w + v.w; }
    accum[pix] = v;
}
cudaError_t launch_unpack(float4 *accum, const float4 *packed, uint32_t width, uint32_t height, uint32_t tile_rank,
                          uint32_t tile_count, int add, cudaStream_t stream)
{
    RenderParams rp{};
    rp.width = width; rp.height = height;
    rp.tiles_x = (width + TILE - 1) / TILE; rp.tiles_y = (height + TILE - 1) / TILE;
    rp.tile_rank = tile_rank; rp.tile_count = tile_count;
    const uint32_t n_tiles = rp.tiles_x * rp.tiles_y;
    const uint32_t owned = tile_rank < n_tiles ? (n_tiles - tile_rank + tile_count - 1) / tile_count : 0;
    rp.n_work = owned * TILE_PX;
    if (rp.n_work == 0) return cudaSuccess;
    k_unpack<<<(rp.n_work + 255u) / 256u, 256, 0, stream>>>(accum, packed, rp, add);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Present filter: Assets/Fullscreen.frag:14-31 on the two traced images (bilinear, clamp-to-border black
// sampler of Source/GraphicsDevice.cpp:770-794), one thread per framebuffer pixel.
// ------------------------------------------------------------------------------------------------
VKRT_DEV V3 sample_bilinear(const uchar4 *__restrict__ img, uint32_t w, uint32_t h, float u, float v)
{
    const float s = u * (float)w - 0.5f, t = v * (float)h - 0.5f;
    const float fs0 = floorf(s), ft0 = floorf(t);
    const float fx = s - fs0, fy = t - ft0;
    const int i0 = (int)fs0, j0 = (int)ft0;
    V3 c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + (k & 1), j = j0 + (k >> 1);
        if (i < 0 || j < 0 || i >= (int)w || j >= (int)h) c[k] = v3(0.0f);
        else { const uchar4 p = __ldg(img + (size_t)j * w + (size_t)i); c[k] = v3((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f); }
    }
    return mix3(mix3(c[0], c[1], fx), mix3(c[2], c[3], fx), fy);
}

__global__ void __launch_bounds__(256) k_present(const uchar4 *__restrict__ b0, const uchar4 *__restrict__ b1, uint32_t tw, uint32_t th,
                                                  uchar4 *__restrict__ out, uint32_t W, uint32_t H)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const uint32_t x = i % W, y = i / W;
    const float inv_size = 1.0f / 2048.0f;
    const float u = ((float)x + 0.5f) / (float)W, v = 1.0f - ((float)y + 0.5f) / (float)H;
    const V3 c0 = sample_bilinear(b0, tw, th, u, v);
    const V3 c1 = sample_bilinear(b1, tw, th, u, v);
    const V3 N = sample_bilinear(b0, tw, th, u + 0.0f * inv_size, v + 1.0f * inv_size);
    const V3 S = sample_bilinear(b0, tw, th, u + 0.0f * inv_size, v + -1.0f * inv_size);
    const V3 E = sample_bilinear(b0, tw, th, u + 1.0f * inv_size, v + 0.0f * inv_size);
    const V3 Wc = sample_bilinear(b0, tw, th, u + -1.0f * inv_size, v + 0.0f * inv_size);
    const V3 dlt = c0 - c1;
    const float vT = dot3(dlt, dlt);
    const V3 fin = vT > 0.0005f ? (((N + S) + E) + Wc) / 4.0f : c0;
    out[i] = make_uchar4(unorm8(fin.x), unorm8(fin.y), unorm8(fin.z), 255);
}

cudaError_t launch_present(const uchar4 *b0, const uchar4 *b1, uint32_t tw, uint32_t th, uchar4 *out, uint32_t W, uint32_t H,
                           cudaStream_t stream)
{
    k_present<<<(W * H + 255u) / 256u, 256, 0, stream>>>(b0, b1, tw, th, out, W, H);
    return cudaGetLastError();
}

cudaError_t launch_clear_accum(float4 *accum, size_t n, cudaStream_t stream)
{
    return cudaMemsetAsync(accum, 0, n * sizeof(float4), stream);
}

} // namespace vkrt
