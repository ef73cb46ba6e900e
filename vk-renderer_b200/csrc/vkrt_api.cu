// C ABI of libvkrt_cuda (include/vkrt.h): context, scene upload, per-frame draw, outputs.
//
// Mirrors the ownership model of the reference's GraphicsDevice (Source/GraphicsDevice.cpp:40-43:
// the library owns every GPU object; the caller owns FrameData, copied inside Draw at :1258).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <new>
#include <string>
#include <vector>

#include "vkrt_device.cuh"
#include "vkrt_internal.h"

using namespace vkrt;

static_assert(sizeof(vkrt_camera_data) == 64, "CameraData must be 64 B (ref: Include/Camera.h:5-12)");
static_assert(sizeof(vkrt_frame_data) == 96, "FrameData must be 96 B (ref: Include/GraphicsDevice.h:20-29)");
static_assert(offsetof(vkrt_frame_data, aspect_ratio) == 0 && offsetof(vkrt_frame_data, seed) == 4 &&
              offsetof(vkrt_frame_data, light_pos) == 16 && offsetof(vkrt_frame_data, camera) == 32,
              "FrameData offsets (Tracer.comp.spv Offset decorations 0/4/16/32)");
static_assert(offsetof(vkrt_camera_data, dir) == 16 && offsetof(vkrt_camera_data, right) == 32 &&
              offsetof(vkrt_camera_data, up) == 48, "Camera offsets 0/16/32/48");
static_assert(sizeof(vkrt_triangle) == 48, "Triangle must be 48 B (ref: Include/GraphicsDevice.h:13-18)");
static_assert(sizeof(vkrt_material) == 48 && sizeof(vkrt_sphere) == 16 && sizeof(vkrt_plane) == 16, "scene records");
static_assert(sizeof(DevScene) <= 1024 && sizeof(RenderParams) <= 512, "kernel parameter budget");

#ifndef VKRT_WAVE_LANES
#define VKRT_WAVE_LANES 2      // wavefront: waves of a frame alternate between this many buffer sets / streams
#endif

static thread_local std::string g_create_error;

// One traced-image slot whose pixels live outside the library's own dense image (include/vkrt.h "interop"): linear
// memory the caller bound or an imported Vulkan allocation mapped as a buffer (kind 1), or a CUDA array behind a
// surface object = an imported VK_IMAGE_TILING_OPTIMAL image / the diagnostics array (kind 2).
struct ExtSlot {
    int kind = 0;
    cudaExternalMemory_t mem = nullptr;       // the imported allocation (owns the fd after a successful import)
    cudaMipmappedArray_t mip = nullptr;       // OPTIMAL import: the image as a 1-level mipmapped array
    cudaArray_t arr = nullptr;                // level 0 of `mip`, or the library-owned diagnostics array (own_arr)
    bool own_arr = false;
    cudaSurfaceObject_t surf = 0;
    uchar4 *ptr = nullptr; size_t pitch = 0;  // kind 1
    cudaExternalSemaphore_t sem[2] = {nullptr, nullptr};
    uint64_t frames = 0;                      // frames (per-draw resolves) written into this slot since its first semaphore was imported
};

struct vkrt_ctx {
    vkrt_create_info info{};
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // scene (host mirrors + device buffers)
    std::vector<vkrt_material> mats;
    std::vector<vkrt_sphere> spheres;
    std::vector<uint32_t> sphere_mat;
    std::vector<vkrt_plane> planes;
    std::vector<uint32_t> plane_mat;
    std::vector<vkrt_triangle> tris;
    std::vector<uint32_t> tri_mats;       // per-triangle material ids (empty: every triangle uses tri_mat)
    uint32_t tri_mat = 0;
    bool scene_dirty = true;
    float4 *d_spheres = nullptr, *d_mats = nullptr, *d_tris = nullptr;
    uint32_t *d_sphere_mat = nullptr, *d_tri_mats = nullptr;
    BvhBuild bvh;
    TriBvhBuild tbvh;                      // the triangles' own tree (built by vkrt_build_bvh when the scene has triangles)
    bool use_bvh = false;
    DevScene dev{};

    // frame state
    uint32_t spp = 4, max_depth = 4;
    uint64_t seed = 0;
    uint32_t frame_index = 0;
    uint32_t last_fkey = 0;
    vkrt_frame_data last_fd{};
    uint32_t frames_drawn = 0;
    uint32_t n_tiles = 0, owned_tiles = 0;
    bool accum_valid = false;

    float4 *d_accum = nullptr;
    uint32_t *d_hit_ids = nullptr;
    std::vector<uchar4 *> d_rgba;
    std::vector<ExtSlot> ext;      // one per rgba8 target
    uint32_t cur_target = 0;
    unsigned long long *d_counters = nullptr;
    uint32_t *d_work_head = nullptr;
    float4 *d_packed = nullptr;
    uchar4 *d_present = nullptr; size_t present_bytes = 0;
    uint64_t frames = 0;
    WaveEngine wave{};
    bool wave_ready = false;

    cudaEvent_t ev_begin = nullptr, ev_trace0 = nullptr, ev_trace1 = nullptr, ev_end = nullptr, ev_consumed = nullptr;
    bool pending_join = false;     // the last frame ended on a lane stream that `stream` has not waited for yet
    bool timing_valid = false;
    uint32_t last_launches = 0;

    // frame exchange (include/vkrt.h, csrc/vkrt_exchange.cu): the block in the gathering GPU's memory as this context sees it
    struct Exchange {
        char *base = nullptr; size_t bytes = 0;
        bool owner = false, ipc = false;
        uint32_t T = 1, S = 1, rank = 0;          // rank = sample_rank * T + tile_rank
        uint64_t frame = 0;                        // frames drawn since the exchange was attached
    } xch;
    // in-library multi-GPU: this context is the gathering one, the children render the other tile shards
    std::vector<vkrt_ctx *> children;
};

namespace {

vkrt_error fail(vkrt_ctx *c, vkrt_error code, const std::string &msg)
{
    // same prefix as the reference's diagnostics (Source/GraphicsDevice.cpp:25)
    const std::string full = "[app] - err :: " + msg;
    if (c) c->err = full; else g_create_error = full;
    return code;
}
vkrt_error cuda_fail(vkrt_ctx *c, cudaError_t e, const char *what)
{
    return fail(c, VKRT_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(c, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail((c), _e, #call); } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

vkrt_material make_mat(float r, float g, float b, float e, float rough, float metal, uint32_t type)
{
    vkrt_material m{};
    m.albedo[0] = r; m.albedo[1] = g; m.albedo[2] = b;
    m.emissive[0] = m.emissive[1] = m.emissive[2] = e;
    m.roughness = rough; m.metalness = metal; m.type = type;
    return m;
}

vkrt_error upload_scene(vkrt_ctx *c)
{
    if (!c->scene_dirty) return VKRT_SUCCESS;
    cudaDeviceSynchronize();          // frames in flight may still read the buffers replaced below
    for (uint32_t m : c->sphere_mat) if (m >= c->mats.size()) return fail(c, VKRT_BAD_ARG, "sphere material id out of range");
    for (uint32_t m : c->plane_mat) if (m >= c->mats.size()) return fail(c, VKRT_BAD_ARG, "plane material id out of range");
    if (!c->tris.empty() && c->tri_mat >= c->mats.size()) return fail(c, VKRT_BAD_ARG, "triangle material id out of range");
    if (!c->tri_mats.empty() && c->tri_mats.size() != c->tris.size()) return fail(c, VKRT_BAD_ARG, "per-triangle materials: one id per triangle");
    for (uint32_t m : c->tri_mats) if (m >= c->mats.size()) return fail(c, VKRT_BAD_ARG, "triangle material id out of range");
    for (const vkrt_material &m : c->mats) if (m.type > 1u) return fail(c, VKRT_BAD_ARG, "unknown material type");

    // the emissive-sphere list of Tracer.comp:458-462, in sphere order
    std::vector<uint32_t> lights;
    for (size_t i = 0; i < c->spheres.size(); ++i) {
        const vkrt_material &m = c->mats[c->sphere_mat[i]];
        if (!(m.emissive[0] == 0.0f && m.emissive[1] == 0.0f && m.emissive[2] == 0.0f)) lights.push_back((uint32_t)i);
    }
    if (lights.size() > MAX_LIGHTS) return fail(c, VKRT_BAD_ARG, "more than 14 emissive spheres");

    auto up = [&](auto **dptr, const void *src, size_t bytes) -> cudaError_t {
        if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
        if (bytes == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void **)dptr, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    CU(c, up(&c->d_spheres, c->spheres.data(), c->spheres.size() * sizeof(vkrt_sphere)));
    CU(c, up(&c->d_sphere_mat, c->sphere_mat.data(), c->sphere_mat.size() * sizeof(uint32_t)));
    CU(c, up(&c->d_mats, c->mats.data(), c->mats.size() * sizeof(vkrt_material)));
    CU(c, up(&c->d_tris, c->tris.data(), c->tris.size() * sizeof(vkrt_triangle)));
    CU(c, up(&c->d_tri_mats, c->tri_mats.data(), c->tri_mats.size() * sizeof(uint32_t)));
    CU(c, cudaStreamSynchronize(c->stream));

    DevScene &d = c->dev;
    std::memset(&d, 0, sizeof(d));
    d.spheres = c->d_spheres; d.sphere_mat = c->d_sphere_mat; d.mats = c->d_mats; d.tris = c->d_tris;
    d.bvh = c->use_bvh ? c->bvh.nodes : nullptr;
    d.bvh4 = c->use_bvh ? c->bvh.nodes4 : nullptr;
    d.qbvh = c->use_bvh ? c->bvh.qnodes : nullptr;
    for (int k = 0; k < 3; ++k) { d.qs[k] = c->bvh.qgrid[k]; d.qb2[k] = c->bvh.qgrid[3 + k]; }
    d.n_nodes = c->use_bvh ? c->bvh.n_nodes : 0;
    d.tbvh = (c->use_bvh && c->tbvh.n_nodes) ? c->tbvh.nodes : nullptr;
    d.n_tnodes = d.tbvh ? c->tbvh.n_nodes : 0;
    d.tri_mats = c->d_tri_mats;
    d.n_spheres = (uint32_t)c->spheres.size(); d.n_tris = (uint32_t)c->tris.size(); d.tri_mat = c->tri_mat;
    d.n_planes = (uint32_t)c->planes.size(); d.n_mats = (uint32_t)c->mats.size();
    for (size_t i = 0; i < c->planes.size(); ++i) {
        d.planes[i] = make_float4(c->planes[i].nx, c->planes[i].ny, c->planes[i].nz, c->planes[i].len);
        d.plane_mat[i] = c->plane_mat[i];
    }
    d.n_lights = (uint32_t)lights.size();
    for (size_t i = 0; i < lights.size(); ++i) d.lights[i] = lights[i];
    c->scene_dirty = false;
    return VKRT_SUCCESS;
}

// Frames of the wavefront variant end on a lane stream (two frames in flight).  Everything that consumes a frame
// through the context's stream first makes that stream wait for the frame's completion event.
cudaError_t join(vkrt_ctx *c)
{
    if (!c->pending_join) return cudaSuccess;
    c->pending_join = false;
    return cudaStreamWaitEvent(c->stream, c->ev_end, 0);
}

void fill_params(vkrt_ctx *c, RenderParams &rp)
{
    std::memset(&rp, 0, sizeof(rp));
    rp.fd = c->last_fd;
    rp.width = c->info.width; rp.height = c->info.height;
    rp.tiles_x = (rp.width + TILE - 1) / TILE; rp.tiles_y = (rp.height + TILE - 1) / TILE;
    rp.tile_rank = c->info.tile_shard_rank; rp.tile_count = c->info.tile_shard_count;
    rp.n_work = c->owned_tiles * TILE_PX;
    rp.max_depth = c->max_depth;
    rp.fkey = c->last_fkey;
    const uint32_t sc = c->info.sample_shard_count, sr = c->info.sample_shard_rank;
    rp.s_begin = (uint32_t)((uint64_t)sr * c->spp / sc);
    rp.s_end = (uint32_t)((uint64_t)(sr + 1) * c->spp / sc);
    rp.accum = c->d_accum; rp.hit_ids = c->d_hit_ids; rp.counters = c->d_counters; rp.work_head = c->d_work_head;
}

// detaches slot `e` from whatever backs it (the caller has made the device idle)
void ext_release_target(ExtSlot &e)
{
    if (e.surf) cudaDestroySurfaceObject(e.surf);
    if (e.own_arr && e.arr) cudaFreeArray(e.arr);
    if (e.mip) cudaFreeMipmappedArray(e.mip);
    if (e.mem && e.ptr) cudaFree(e.ptr);          // a mapped buffer of an imported allocation is freed with cudaFree
    if (e.mem) cudaDestroyExternalMemory(e.mem);
    e.kind = 0; e.surf = 0; e.arr = nullptr; e.own_arr = false; e.mip = nullptr; e.ptr = nullptr; e.pitch = 0; e.mem = nullptr;
}
void ext_release_semaphores(ExtSlot &e)
{
    for (auto &sm : e.sem) { if (sm) cudaDestroyExternalSemaphore(sm); sm = nullptr; }
    e.frames = 0;
}
// The library never consumes the caller's descriptor: it imports a duplicate.  A successful import hands the
// duplicate to the CUDA driver; after a failed one it is closed here unless the driver already did.
int dup_fd(int fd) { return fcntl(fd, F_DUPFD_CLOEXEC, 0); }
void close_if_open(int fd) { if (fd >= 0 && fcntl(fd, F_GETFD) != -1) close(fd); }

vkrt_error make_idle(vkrt_ctx *c)
{
    CU(c, join(c));
    CU(c, cudaStreamSynchronize(c->stream));
    return VKRT_SUCCESS;
}

// The resolve of the frame into the current target (Tracer.comp:585-592), bracketed by the slot's semaphores: what
// the reference's two vkCmdPipelineBarriers around the dispatch do (Source/GraphicsDevice.cpp:1234-1252, :1268-1284).
// `frame` = this is the resolve of a new frame (vkrt_draw); an explicit re-resolve of the same frame (vkrt_resolve,
// e.g. after a shard gather) rewrites the image the engine already owns for this frame and neither waits nor signals:
// the engine never signals ACQUIRE twice for one frame, and a binary RELEASE must not be signalled twice.
vkrt_error resolve_current(vkrt_ctx *c, const RenderParams &rp, cudaStream_t st, bool frame)
{
    ExtSlot &e = c->ext[c->cur_target];
    ResolveTarget tg{c->d_rgba[c->cur_target], (size_t)c->info.width * 4u, 0};
    if (e.kind == 1) { tg.ptr = e.ptr; tg.pitch = e.pitch; }
    else if (e.kind == 2) { tg.ptr = nullptr; tg.surf = e.surf; }
    const bool sync = frame && (e.sem[VKRT_SEMAPHORE_ACQUIRE] || e.sem[VKRT_SEMAPHORE_RELEASE]);
    if (sync) ++e.frames;
    if (sync && e.sem[VKRT_SEMAPHORE_ACQUIRE] && e.frames >= 2) {
        cudaExternalSemaphoreWaitParams wp{};
        wp.params.fence.value = e.frames - 1;            // timeline semaphores; a binary semaphore ignores the value
        CU(c, cudaWaitExternalSemaphoresAsync(&e.sem[VKRT_SEMAPHORE_ACQUIRE], &wp, 1, st));
    }
    CU(c, launch_resolve(rp, c->info.integrator, tg, st));
    if (sync && e.sem[VKRT_SEMAPHORE_RELEASE]) {
        cudaExternalSemaphoreSignalParams sp{};
        sp.params.fence.value = e.frames;
        CU(c, cudaSignalExternalSemaphoresAsync(&e.sem[VKRT_SEMAPHORE_RELEASE], &sp, 1, st));
    }
    return VKRT_SUCCESS;
}

unsigned long long *x_done(vkrt_ctx *c, uint32_t r) { return reinterpret_cast<unsigned long long *>(c->xch.base) + (size_t)r * X_FLAG_STRIDE; }
unsigned long long *x_consumed(vkrt_ctx *c) { return reinterpret_cast<unsigned long long *>(c->xch.base + X_OFF_CONSUMED); }
uint32_t *x_error(vkrt_ctx *c) { return reinterpret_cast<uint32_t *>(c->xch.base + X_OFF_ERROR); }
float4 *x_target(vkrt_ctx *c, uint64_t frame, uint32_t sample_group)
{
    const size_t n_px = (size_t)c->info.width * c->info.height;
    return reinterpret_cast<float4 *>(c->xch.base + X_OFF_TARGETS) + ((size_t)(frame & 1u) * c->xch.S + sample_group) * n_px;
}
const uint32_t X_MAGIC = 0x58524b56u;     // "VKRX"

// a context on one device (the shard fields of `info` say which part of the frame it renders)
vkrt_error create_single(const vkrt_create_info *info, vkrt_ctx **out_ctx);

// applies a call to the other devices' contexts of an in-library multi-GPU context
template <class F>
vkrt_error for_children(vkrt_ctx *c, F f)
{
    for (vkrt_ctx *ch : c->children) {
        const vkrt_error r = f(ch);
        if (r != VKRT_SUCCESS) return fail(c, r, "device " + std::to_string(ch->info.device_id) + ": " + ch->err);
    }
    return VKRT_SUCCESS;
}
#define FWD(c, call) do { if (!(c)->children.empty()) { const vkrt_error _r = for_children((c), [&](vkrt_ctx *ch) { return call; }); if (_r != VKRT_SUCCESS) return _r; } } while (0)

} // namespace

extern "C" {

VKRT_API const char *vkrt_version(void) { return "libvkrt_cuda 0.1 (sm_100a)"; }

VKRT_API const char *vkrt_last_error_string(vkrt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

VKRT_API vkrt_error vkrt_create(const vkrt_create_info *info, vkrt_ctx **out_ctx)
{
    if (!info || !out_ctx) return fail(nullptr, VKRT_BAD_ARG, "null argument");
    *out_ctx = nullptr;
    if (info->struct_size != sizeof(vkrt_create_info)) return fail(nullptr, VKRT_BAD_ARG, "vkrt_create_info.struct_size mismatch");
    if (info->n_devices <= 1) {
        vkrt_create_info one = *info;
        if (info->n_devices == 1) one.device_id = info->device_ids[0];
        one.n_devices = 0;
        return create_single(&one, out_ctx);
    }
    // in-library multi-GPU: one context per device, interleaved tile shards, frame exchange into device_ids[0]
    const uint32_t n = info->n_devices;
    if (n > 8) return fail(nullptr, VKRT_BAD_ARG, "n_devices > 8");
    if (info->tile_shard_count > 1 || info->sample_shard_count > 1 || info->tile_shard_rank || info->sample_shard_rank)
        return fail(nullptr, VKRT_BAD_ARG, "n_devices > 1: the library shards the frame by itself, leave the shard fields 0");
    if (info->flags & VKRT_FLAG_NO_RESOLVE) return fail(nullptr, VKRT_BAD_ARG, "n_devices > 1 with VKRT_FLAG_NO_RESOLVE");
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (info->device_ids[i] == info->device_ids[j]) return fail(nullptr, VKRT_BAD_ARG, "device_ids must be distinct");
    vkrt_ctx *parent = nullptr;
    vkrt_create_info one = *info;
    one.n_devices = 0; one.device_id = info->device_ids[0]; one.tile_shard_rank = 0; one.tile_shard_count = n;
    vkrt_error r = create_single(&one, &parent);
    if (r != VKRT_SUCCESS) return r;
    vkrt_exchange_handle h;
    r = vkrt_exchange_create(parent, &h);
    if (r != VKRT_SUCCESS) { g_create_error = parent->err; vkrt_destroy(parent); return r; }
    for (uint32_t i = 1; i < n; ++i) {
        vkrt_ctx *ch = nullptr;
        one.device_id = info->device_ids[i]; one.tile_shard_rank = i; one.stream = nullptr;
        one.flags = (info->flags | VKRT_FLAG_NO_RESOLVE) & ~(uint32_t)VKRT_FLAG_PROGRESSIVE;
        r = create_single(&one, &ch);
        if (r == VKRT_SUCCESS) {
            parent->children.push_back(ch);
            r = vkrt_exchange_attach(ch, parent);
            if (r != VKRT_SUCCESS) g_create_error = ch->err;
        }
        if (r != VKRT_SUCCESS) { vkrt_destroy(parent); return r; }
    }
    *out_ctx = parent;
    return VKRT_SUCCESS;
}

} // extern "C"

namespace {
vkrt_error create_single(const vkrt_create_info *info, vkrt_ctx **out_ctx)
{
    if (info->width == 0 || info->height == 0 || info->width > 65536 || info->height > 65536)
        return fail(nullptr, VKRT_BAD_ARG, "bad resolution");
    if ((uint64_t)info->width * info->height > ((uint64_t)1 << 30))
        return fail(nullptr, VKRT_BAD_ARG, "more than 2^30 pixels (pixel and path indices are 32-bit)");
    if (info->integrator > VKRT_INTEGRATOR_PATH || info->variant > VKRT_VARIANT_WAVEFRONT)
        return fail(nullptr, VKRT_BAD_ARG, "bad integrator / variant");
    // the same ranges as vkrt_set_sampling (0 = the shader's default): the wavefront keeps one counter set per depth
    // (C_SETS) and packs the depth into 8 bits of the path record, the sample index into 24
    if (info->spp > (1u << 24) || info->max_depth > 255)
        return fail(nullptr, VKRT_BAD_ARG, "spp / max_depth out of range (spp <= 2^24, max_depth <= 255)");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, VKRT_NO_SUITABLE_GPU, "no CUDA device (libvkrt_cuda has no CPU fallback)");
    }
    if (info->device_id < 0 || info->device_id >= n_dev) return fail(nullptr, VKRT_NO_SUITABLE_GPU, "device_id out of range");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, info->device_id) != cudaSuccess || prop.major < 10)
        return fail(nullptr, VKRT_NO_SUITABLE_GPU, "device is not sm_100-class (this library ships sm_100a code only)");

    vkrt_ctx *c = new (std::nothrow) vkrt_ctx();
    if (!c) return fail(nullptr, VKRT_UNKNOWN, "out of host memory");
    c->info = *info;
    if (c->info.tile_shard_count == 0) c->info.tile_shard_count = 1;
    if (c->info.sample_shard_count == 0) c->info.sample_shard_count = 1;
    if (c->info.frames_in_flight == 0) c->info.frames_in_flight = 2;
    if (c->info.tile_shard_rank >= c->info.tile_shard_count || c->info.sample_shard_rank >= c->info.sample_shard_count ||
        c->info.frames_in_flight > 8) {
        delete c;
        return fail(nullptr, VKRT_BAD_ARG, "bad shard rank / frames_in_flight");
    }
    c->spp = info->spp ? info->spp : 4;                 // SAMPLES (Tracer.comp:180)
    c->max_depth = info->max_depth ? info->max_depth : (info->integrator == VKRT_INTEGRATOR_PATH ? 4 : 2);
    c->sm_count = prop.multiProcessorCount;
    DeviceGuard g(info->device_id);
#define CC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { vkrt_error r = cuda_fail(nullptr, _e, #call); vkrt_destroy(c); return r; } } while (0)
    if (info->stream) c->stream = (cudaStream_t)info->stream;
    else { CC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    const size_t n_px = (size_t)info->width * info->height;
    const uint32_t tiles_x = (info->width + TILE - 1) / TILE, tiles_y = (info->height + TILE - 1) / TILE;
    c->n_tiles = tiles_x * tiles_y;
    c->owned_tiles = c->info.tile_shard_rank < c->n_tiles
                         ? (c->n_tiles - c->info.tile_shard_rank + c->info.tile_shard_count - 1) / c->info.tile_shard_count : 0;
    CC(cudaMalloc(&c->d_accum, n_px * sizeof(float4)));
    CC(cudaMemsetAsync(c->d_accum, 0, n_px * sizeof(float4), c->stream));
    if (info->flags & VKRT_FLAG_HIT_IDS) {
        CC(cudaMalloc(&c->d_hit_ids, n_px * sizeof(uint32_t)));
        CC(cudaMemsetAsync(c->d_hit_ids, 0, n_px * sizeof(uint32_t), c->stream));
    }
    c->d_rgba.assign(c->info.frames_in_flight, nullptr);
    c->ext.assign(c->info.frames_in_flight, ExtSlot{});
    c->cur_target = c->info.frames_in_flight - 1;       // the first frame goes to image slot 0 like state.currentFrame
    for (auto &p : c->d_rgba) { CC(cudaMalloc(&p, n_px * sizeof(uchar4))); CC(cudaMemsetAsync(p, 0, n_px * sizeof(uchar4), c->stream)); }
    CC(cudaMalloc(&c->d_counters, CNT_N * sizeof(unsigned long long)));
    CC(cudaMemsetAsync(c->d_counters, 0, CNT_N * sizeof(unsigned long long), c->stream));
    CC(cudaMalloc(&c->d_work_head, 64));
    CC(cudaMemsetAsync(c->d_work_head, 0, 64, c->stream));
    CC(cudaEventCreate(&c->ev_begin)); CC(cudaEventCreate(&c->ev_trace0));
    CC(cudaEventCreate(&c->ev_trace1)); CC(cudaEventCreate(&c->ev_end));
    CC(cudaEventCreateWithFlags(&c->ev_consumed, cudaEventDisableTiming));
    CC(cudaStreamSynchronize(c->stream));
#undef CC
    *out_ctx = c;
    return VKRT_SUCCESS;
}
} // namespace

extern "C" {

VKRT_API vkrt_error vkrt_destroy(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    if (!c->children.empty()) vkrt_wait_idle(c);       // the devices wait for each other's frame counters
    for (vkrt_ctx *ch : c->children) vkrt_destroy(ch);
    c->children.clear();
    DeviceGuard g(c->info.device_id);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->xch.base) vkrt_exchange_close(c);
    cudaFree(c->d_spheres); cudaFree(c->d_sphere_mat); cudaFree(c->d_mats); cudaFree(c->d_tris); cudaFree(c->d_tri_mats);
    cudaFree(c->tbvh.nodes);
    cudaFree(c->bvh.nodes); cudaFree(c->bvh.nodes4); cudaFree(c->bvh.qnodes); cudaFree(c->bvh.tnodes); cudaFree(c->d_accum); cudaFree(c->d_hit_ids);
    for (auto p : c->d_rgba) cudaFree(p);
    for (auto &e : c->ext) { ext_release_target(e); ext_release_semaphores(e); }
    cudaFree(c->d_counters); cudaFree(c->d_work_head); cudaFree(c->d_packed); cudaFree(c->d_present);
    if (c->wave_ready) wave_engine_free(c->wave);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_trace0) cudaEventDestroy(c->ev_trace0);
    if (c->ev_trace1) cudaEventDestroy(c->ev_trace1);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->ev_consumed) cudaEventDestroy(c->ev_consumed);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_set_sampling(vkrt_ctx *c, uint32_t spp, uint32_t max_depth)
{
    if (!c) return VKRT_BAD_ARG;
    if (spp == 0 || max_depth == 0 || spp > (1u << 24) || max_depth > 255) return fail(c, VKRT_BAD_ARG, "spp / max_depth out of range");
    c->spp = spp; c->max_depth = max_depth;
    FWD(c, vkrt_set_sampling(ch, spp, max_depth));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_seed(vkrt_ctx *c, uint64_t seed)
{
    if (!c) return VKRT_BAD_ARG;
    c->seed = seed;
    FWD(c, vkrt_set_seed(ch, seed));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_frame_index(vkrt_ctx *c, uint32_t i)
{
    if (!c) return VKRT_BAD_ARG;
    c->frame_index = i;
    FWD(c, vkrt_set_frame_index(ch, i));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_reset_accum(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    CU(c, launch_clear_accum(c->d_accum, (size_t)c->info.width * c->info.height, c->stream));
    c->accum_valid = false;
    return VKRT_SUCCESS;
}

// ---- scene --------------------------------------------------------------------------------------
VKRT_API vkrt_error vkrt_set_triangles(vkrt_ctx *c, const vkrt_triangle *t, uint32_t n)
{
    if (!c || (n && !t)) return VKRT_BAD_ARG;
    c->tris.assign(t, t + n); c->scene_dirty = true;
    c->tri_mats.clear();                       // a new triangle list: back to the one shared material
    c->tbvh.n_nodes = 0;                       // ... and its tree is gone (vkrt_build_bvh makes a new one)
    FWD(c, vkrt_set_triangles(ch, t, n));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_triangle_material(vkrt_ctx *c, uint32_t mat_id)
{
    if (!c) return VKRT_BAD_ARG;
    c->tri_mat = mat_id; c->scene_dirty = true;
    FWD(c, vkrt_set_triangle_material(ch, mat_id));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_triangle_materials(vkrt_ctx *c, const uint32_t *mat_ids, uint32_t n)
{
    if (!c || (n && !mat_ids)) return VKRT_BAD_ARG;
    if (n != 0 && n != c->tris.size()) return fail(c, VKRT_BAD_ARG, "per-triangle materials: one id per triangle (set the triangles first)");
    c->tri_mats.assign(mat_ids, mat_ids + n); c->scene_dirty = true;
    FWD(c, vkrt_set_triangle_materials(ch, mat_ids, n));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_materials(vkrt_ctx *c, const vkrt_material *m, uint32_t n)
{
    if (!c || (n && !m)) return VKRT_BAD_ARG;
    for (uint32_t i = 0; i < n; ++i) if (m[i].type > 1u) return fail(c, VKRT_BAD_ARG, "unknown material type");
    c->mats.assign(m, m + n); c->scene_dirty = true;
    FWD(c, vkrt_set_materials(ch, m, n));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_spheres(vkrt_ctx *c, const vkrt_sphere *s, const uint32_t *mat_id, uint32_t n)
{
    if (!c || (n && (!s || !mat_id))) return VKRT_BAD_ARG;
    c->spheres.assign(s, s + n); c->sphere_mat.assign(mat_id, mat_id + n);
    c->use_bvh = false;                       // a new sphere list invalidates the tree
    c->scene_dirty = true;
    FWD(c, vkrt_set_spheres(ch, s, mat_id, n));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_set_planes(vkrt_ctx *c, const vkrt_plane *p, const uint32_t *mat_id, uint32_t n)
{
    if (!c || (n && (!p || !mat_id))) return VKRT_BAD_ARG;
    if (n > MAX_PLANES) return fail(c, VKRT_BAD_ARG, "more than 16 planes");
    c->planes.assign(p, p + n); c->plane_mat.assign(mat_id, mat_id + n); c->scene_dirty = true;
    FWD(c, vkrt_set_planes(ch, p, mat_id, n));
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_use_default_scene(vkrt_ctx *c, uint32_t which)
{
    if (!c) return VKRT_BAD_ARG;
    // the one triangle the host uploads, Source/GraphicsDevice.cpp:798-803 (== Raytracer.comp:116)
    vkrt_triangle tri{};
    tri.v0 = {10.0f, 10.0f, 0.0f, 0.0f}; tri.v1 = {0.0f, 20.0f, 0.0f, 0.0f}; tri.v2 = {-10.0f, 10.0f, 0.0f, 0.0f};
    if (which == VKRT_SCENE_TRACER) {
        const vkrt_material mats[8] = {                                   // Tracer.comp:186-194
            make_mat(1.0f, 1.0f, 1.0f, 0.0f, 0.3f, 0.7f, VKRT_MAT_DIFFUSE),      // 0 matte_white
            make_mat(0.75f, 0.25f, 0.25f, 0.0f, 0.4f, 0.0f, VKRT_MAT_DIFFUSE),   // 1 matte_red
            make_mat(0.25f, 0.75f, 0.25f, 0.0f, 0.4f, 0.0f, VKRT_MAT_DIFFUSE),   // 2 matte_green
            make_mat(0.25f, 0.25f, 0.75f, 0.0f, 0.4f, 0.0f, VKRT_MAT_DIFFUSE),   // 3 matte_blue
            make_mat(0.25f, 0.25f, 0.75f, 0.0f, 0.3f, 0.6f, VKRT_MAT_DIFFUSE),   // 4 plastic
            make_mat(1.0f, 0.5f, 0.5f, 0.0f, 0.0f, 1.0f, VKRT_MAT_DIFFUSE),      // 5 mirror
            make_mat(1.0f, 1.0f, 1.0f, 0.0f, 0.42f, 0.0f, VKRT_MAT_DIELECTRIC),  // 6 glass
            make_mat(1.0f, 1.0f, 1.0f, 128.0f, 0.6f, 0.0f, VKRT_MAT_DIFFUSE)};   // 7 light
        const vkrt_sphere sp[4] = {{42.0f, 16.0f, 12.0f, 16.0f}, {0.0f, 96.0f, 0.0f, 12.0f},   // Tracer.comp:196-202
                                   {-32.0f, 24.0f, 24.0f, 24.0f}, {-24.0f, 11.0f, -48.0f, 11.0f}};
        const uint32_t spm[4] = {6, 7, 5, 4};
        const vkrt_plane pl[5] = {{0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, -1.0f, 0.0f, 128.0f}, {1.0f, 0.0f, 0.0f, 64.0f}, // :204-211
                                  {0.0f, 0.0f, -1.0f, 64.0f}, {-1.0f, 0.0f, 0.0f, 64.0f}};
        const uint32_t plm[5] = {0, 0, 1, 2, 3};
        vkrt_set_materials(c, mats, 8); vkrt_set_spheres(c, sp, spm, 4); vkrt_set_planes(c, pl, plm, 5);
        vkrt_set_triangles(c, &tri, 1); vkrt_set_triangle_material(c, 5);     // every triangle is `mirror` (:386)
        return VKRT_SUCCESS;
    }
    if (which == VKRT_SCENE_RAYTRACER) {
        auto R = [](bool refl, float r, float g, float b) {
            return make_mat(r, g, b, 0.0f, refl ? 0.0f : 0.4f, refl ? 1.0f : 0.0f, VKRT_MAT_DIFFUSE);
        };
        const vkrt_material mats[6] = {R(true, 1, 1, 1), R(false, 1, 0, 0), R(true, 0, 1, 0),      // Raytracer.comp:119-127
                                       R(false, 0, 0, 1), R(false, 1, 1, 0), R(false, 1, 0, 1)};
        const vkrt_sphere sp[2] = {{-14.0f, 12.0f, 32.0f, 5.0f}, {32.0f, 24.0f, 25.0f, 12.0f}};   // :98-102
        const uint32_t spm[2] = {5, 4};
        const vkrt_plane pl[5] = {{0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, -1.0f, 0.0f, 128.0f}, {0.0f, 0.0f, -1.0f, 64.0f}, // :104-112
                                  {1.0f, 0.0f, 0.0f, 64.0f}, {-1.0f, 0.0f, 0.0f, 64.0f}};
        const uint32_t plm[5] = {0, 0, 2, 1, 3};
        vkrt_set_materials(c, mats, 6); vkrt_set_spheres(c, sp, spm, 2); vkrt_set_planes(c, pl, plm, 5);
        vkrt_set_triangles(c, &tri, 1); vkrt_set_triangle_material(c, 1);     // Raytracer.comp:116
        return VKRT_SUCCESS;
    }
    return fail(c, VKRT_BAD_ARG, "unknown default scene");
}

VKRT_API vkrt_error vkrt_build_bvh(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    c->scene_dirty = true;
    c->use_bvh = false;
    vkrt_error r = upload_scene(c);
    if (r != VKRT_SUCCESS) return r;
    CU(c, build_lbvh(c->d_spheres, (uint32_t)c->spheres.size(), c->bvh, c->stream));
    // the traversal stacks (BVH_STACK entries of thread-local memory, unchecked in the kernels) hold at most one entry
    // per tree level; the LBVH cannot be deeper than 64 levels, and this is where the bound is enforced for both trees
    if (c->bvh.depth + 2 > (int)BVH_STACK)
        return fail(c, VKRT_BAD_ARG, "the LBVH is " + std::to_string(c->bvh.depth) + " levels deep: deeper than the traversal stack");
    // (the traversal tree splits index ranges in the middle from level 48 on: at most 48 + 32 levels)
    if (c->bvh.tdepth + 2 > (int)BVH_STACK)
        return fail(c, VKRT_BAD_ARG, "the traversal tree is " + std::to_string(c->bvh.tdepth) + " levels deep: deeper than the traversal stack");
    // triangles get a tree of their own (rule T); a scene without triangles keeps none
    CU(c, build_tri_lbvh(c->d_tris, (uint32_t)c->tris.size(), c->tbvh, c->stream));
    if (c->tbvh.depth + 2 > (int)BVH_STACK)
        return fail(c, VKRT_BAD_ARG, "the triangle LBVH is " + std::to_string(c->tbvh.depth) + " levels deep: deeper than the traversal stack");
    c->use_bvh = true;
    c->scene_dirty = true;
    FWD(c, vkrt_build_bvh(ch));                // the build is deterministic: every device gets the identical tree
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_clear_bvh(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    c->use_bvh = false; c->scene_dirty = true;
    FWD(c, vkrt_clear_bvh(ch));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_get_bvh_info(vkrt_ctx *c, vkrt_bvh_info *out)
{
    if (!c || !out) return VKRT_BAD_ARG;
    out->n_spheres = (uint32_t)c->spheres.size();
    out->n_nodes = c->use_bvh ? c->bvh.n_nodes : 0;
    out->node_bytes = 64;
    out->build_ms = c->bvh.build_ms;
    out->build_launches = c->bvh.launches;
    out->depth = c->use_bvh ? (uint32_t)c->bvh.depth : 0;
    out->n_triangles = (uint32_t)c->tris.size();
    out->n_tri_nodes = c->use_bvh ? c->tbvh.n_nodes : 0;
    out->tri_depth = c->use_bvh ? (uint32_t)c->tbvh.depth : 0;
    out->tri_build_ms = c->use_bvh ? c->tbvh.build_ms : 0.f;
    out->traversal_depth = c->use_bvh ? (uint32_t)c->bvh.tdepth : 0;
    out->traversal_is_sah = (c->use_bvh && c->bvh.tnodes) ? 1u : 0u;
    out->traversal_build_ms = c->use_bvh ? c->bvh.sah_ms : 0.f;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_read_bvh_nodes(vkrt_ctx *c, float *host, size_t bytes)
{
    if (!c || !host) return VKRT_BAD_ARG;
    if (!c->use_bvh) return fail(c, VKRT_BAD_ARG, "no BVH built");
    if (bytes < (size_t)c->bvh.n_nodes * 64) return fail(c, VKRT_BAD_ARG, "buffer too small");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(host, c->bvh.nodes, (size_t)c->bvh.n_nodes * 64, cudaMemcpyDeviceToHost));
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_read_bvh_traversal_nodes(vkrt_ctx *c, float *host, size_t bytes)
{
    if (!c || !host) return VKRT_BAD_ARG;
    if (!c->use_bvh) return fail(c, VKRT_BAD_ARG, "no BVH built");
    if (bytes < (size_t)c->bvh.n_nodes * 64) return fail(c, VKRT_BAD_ARG, "buffer too small");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(host, c->bvh.tnodes ? c->bvh.tnodes : c->bvh.nodes, (size_t)c->bvh.n_nodes * 64, cudaMemcpyDeviceToHost));
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_read_bvh_qnodes(vkrt_ctx *c, uint32_t *host, size_t bytes, float grid[6])
{
    if (!c || !host || !grid) return VKRT_BAD_ARG;
    if (!c->use_bvh) return fail(c, VKRT_BAD_ARG, "no BVH built");
    if (bytes < (size_t)c->bvh.n_nodes * 32) return fail(c, VKRT_BAD_ARG, "buffer too small");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(host, c->bvh.qnodes, (size_t)c->bvh.n_nodes * 32, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; ++k) grid[k] = c->bvh.qgrid[k];
    return VKRT_SUCCESS;
}

// ---- per-frame ----------------------------------------------------------------------------------
VKRT_API vkrt_error vkrt_draw(vkrt_ctx *c, const vkrt_frame_data *frame)
{
    if (!c || !frame) return VKRT_BAD_ARG;
    FWD(c, vkrt_draw(ch, frame));                 // the other devices' shards of this frame first: they only enqueue
    DeviceGuard g(c->info.device_id);
    vkrt_error r = upload_scene(c);
    if (r != VKRT_SUCCESS) return r;

    c->last_fd = *frame;                                            // "FrameData frame_data_real = frame_data" (:1258)
    uint32_t seed_bits; std::memcpy(&seed_bits, &frame->seed, 4);
    c->last_fkey = frame_key(c->seed, seed_bits, c->frame_index);
    RenderParams rp;
    fill_params(c, rp);
    const bool progressive = (c->info.flags & VKRT_FLAG_PROGRESSIVE) != 0 && c->info.integrator == VKRT_INTEGRATOR_PATH;
    rp.accumulate = (progressive && c->accum_valid) ? 1u : 0u;
    const bool stats = (c->info.flags & VKRT_FLAG_STATS) != 0;
    uint32_t launches = 0;
    // frame exchange: this frame's pixels go straight into the frame target in the gathering GPU's memory.  The target
    // was last read by the collect of frame f - 2; the kernels that overwrite it are ordered behind ev_consumed below.
    const bool exchange = c->xch.base != nullptr;
    const uint64_t xf = c->xch.frame;
    if (exchange) {
        CU(c, join(c));
        if (!c->xch.owner && xf >= 2) { CU(c, launch_xwait(x_consumed(c), 1, 0, xf - 1, x_error(c), c->stream)); ++launches; }
        rp.accum = x_target(c, xf, c->info.sample_shard_rank);
        rp.accumulate = 0;                          // VKRT_FLAG_PROGRESSIVE is applied by the gathering context's collect
    }

    // everything enqueued on the stream so far may still read the previous frame's accumulator / AOVs / image, and
    // so may that frame's own resolve on its lane stream: the stream joins the previous frame first (that does not
    // hold back the new frame's lanes, which only wait for ev_consumed where they overwrite those buffers)
    CU(c, join(c));
    CU(c, cudaEventRecord(c->ev_consumed, c->stream));
    cudaStream_t tail = c->stream;
    const bool wavefront = c->info.integrator == VKRT_INTEGRATOR_PATH && c->info.variant == VKRT_VARIANT_WAVEFRONT;
    if (!wavefront) {
        CU(c, join(c));
        CU(c, cudaEventRecord(c->ev_begin, c->stream));
        CU(c, cudaMemsetAsync(c->d_work_head, 0, 64, c->stream));
    }
    if (rp.n_work > 0 && rp.s_end > rp.s_begin) {
        if (c->info.integrator == VKRT_INTEGRATOR_WHITTED) {
            CU(c, launch_whitted(c->dev, rp, c->use_bvh, stats, c->stream)); ++launches;
        } else if (wavefront) {
            {
                // two lanes (buffer set + stream) of up to 16 samples of every owned pixel each, at most 48 Mi path records
                // per lane (~10 GB of the 180 GB).  A frame of <= 16 spp is ONE wave: the lanes then alternate between
                // consecutive frames (two frames in flight, like the reference's FRAMES_IN_FLIGHT), every launch is as
                // large as the frame allows and one frame's thin deep bounces run next to the next frame's fat early ones
                // (measured on cfg4: 28.5 -> 28.3 ms/frame, and 4.44 -> 3.96 ms for the 1/8 tile shard of an 8-GPU run)
                const size_t slots = (size_t)c->owned_tiles * TILE_PX;
                uint32_t lanes = VKRT_WAVE_LANES;
                if (const char *ev = std::getenv("VKRT_TUNE_LANES")) {          // tuning aid: 1..4 lanes (the image does not depend on it)
                    const int v = std::atoi(ev);
                    if (v >= 1 && v <= 4) lanes = (uint32_t)v;
                }
                size_t per_wave = c->spp < 16u ? c->spp : 16u;
                if (const char *ev = std::getenv("VKRT_TUNE_WAVE_SPP")) {       // tuning aid: samples per wave (the image does not depend on it)
                    const int v = std::atoi(ev);
                    if (v >= 1 && (uint32_t)v <= c->spp) per_wave = (size_t)v;
                }
                while (per_wave > 1 && slots * per_wave > ((size_t)48 << 20)) --per_wave;
                // VKRT_FLAG_SERIAL_WAVES: waves of the same size, one buffer set, one stream
                const uint32_t eng_lanes = (c->info.flags & VKRT_FLAG_SERIAL_WAVES) ? 1u : lanes;
                const size_t cap = slots * per_wave;
                if (cap >= ((size_t)1 << 28))
                    return fail(c, VKRT_BAD_ARG, "the wavefront variant holds at most 2^28 path records per wave: this context owns too many "
                                                 "pixels (shard the frame by tiles, or use the megakernel variant)");
                // sized from the sampling in force: vkrt_set_sampling may have changed it since the last draw
                if (c->wave_ready && (c->wave.n_lanes != eng_lanes || c->wave.lane[0].capacity != cap)) {
                    CU(c, cudaDeviceSynchronize());
                    wave_engine_free(c->wave);
                    c->wave_ready = false;
                    c->pending_join = false;
                }
                if (!c->wave_ready) {
                    CU(c, wave_engine_init(c->wave, cap, eng_lanes));
                    c->wave_ready = true;
                }
            }
            uint32_t nl = 0;
            CU(c, launch_path_wavefront(c->dev, rp, c->wave, c->use_bvh, stats, (c->info.flags & VKRT_FLAG_LAUNCH_TIMING) != 0,
                                        c->sm_count, c->stream, c->ev_consumed,
                                        c->ev_begin, &tail, &nl));
            launches += nl;
        } else {
            CU(c, launch_path_mega(c->dev, rp, c->use_bvh, stats, c->sm_count, c->stream)); ++launches;
        }
    } else if (wavefront) {
        CU(c, join(c));
        CU(c, cudaEventRecord(c->ev_begin, c->stream));
    }
    CU(c, cudaEventRecord(c->ev_trace1, tail));
    if (exchange) {
        // publish "frame xf of this rank is complete" next to the pixels; the gathering context then waits (on the
        // device) for every rank, adds the sample groups in rank order and frees the target for frame xf + 2
        CU(c, launch_xsignal(x_done(c, c->xch.rank), xf + 1, tail)); ++launches;
        if (c->xch.owner) {
            CU(c, launch_xwait(x_done(c, 0), c->xch.T * c->xch.S, X_FLAG_STRIDE, xf + 1, x_error(c), tail)); ++launches;
            CU(c, launch_xcollect(c->d_accum, x_target(c, xf, 0), (size_t)c->info.width * c->info.height, c->xch.S,
                                  (progressive && c->accum_valid) ? 1 : 0, tail)); ++launches;
            CU(c, launch_xsignal(x_consumed(c), xf + 1, tail)); ++launches;
        }
        rp.accum = c->d_accum;
        ++c->xch.frame;
    }
    // frame k targets traced_images[k % FRAMES_IN_FLIGHT] (state.currentFrame, :1234 / :1341); cur_target stays on the
    // most recent frame's slot afterwards, which is what the read-backs return
    c->cur_target = (c->cur_target + 1) % (uint32_t)c->d_rgba.size();
    if (!(c->info.flags & VKRT_FLAG_NO_RESOLVE) && !(exchange && !c->xch.owner)) {     // only the gathering rank has the whole frame
        r = resolve_current(c, rp, tail, true); ++launches;
        if (r != VKRT_SUCCESS) return r;
    }
    CU(c, cudaEventRecord(c->ev_end, tail));
    c->pending_join = (tail != c->stream);
    c->timing_valid = true;
    c->last_launches = launches;
    c->accum_valid = true;
    ++c->frame_index;
    ++c->frames;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_resolve(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    RenderParams rp;
    fill_params(c, rp);
    return resolve_current(c, rp, c->stream, false);
}

VKRT_API vkrt_error vkrt_wait_idle(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    CU(c, cudaStreamSynchronize(c->stream));
    FWD(c, vkrt_wait_idle(ch));
    if (c->xch.base) {
        uint32_t err = 0;
        CU(c, cudaMemcpy(&err, x_error(c), sizeof err, cudaMemcpyDeviceToHost));
        if (err) return fail(c, VKRT_NCCL_ERROR, "frame exchange: a rank did not deliver its frame within the time limit");
    }
    return VKRT_SUCCESS;
}

/* Makes the context's stream wait (on the device, no host sync) for every frame in flight. */
VKRT_API vkrt_error vkrt_flush(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    return VKRT_SUCCESS;
}

// ---- outputs --------------------------------------------------------------------------------------
VKRT_API vkrt_error vkrt_get_rgba8(vkrt_ctx *c, void **dev_ptr, size_t *pitch)
{
    if (!c || !dev_ptr) return VKRT_BAD_ARG;
    { DeviceGuard g(c->info.device_id); CU(c, join(c)); }
    const ExtSlot &e = c->ext[c->cur_target];
    if (e.kind == 2) return fail(c, VKRT_BAD_ARG, "the current target is a CUDA array (imported OPTIMAL image): it has no linear pointer");
    *dev_ptr = e.kind == 1 ? e.ptr : c->d_rgba[c->cur_target];
    if (pitch) *pitch = e.kind == 1 ? e.pitch : (size_t)c->info.width * 4;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_get_accum(vkrt_ctx *c, float **dev_ptr)
{
    if (!c || !dev_ptr) return VKRT_BAD_ARG;
    { DeviceGuard g(c->info.device_id); CU(c, join(c)); }
    *dev_ptr = (float *)c->d_accum;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_get_hit_ids(vkrt_ctx *c, uint32_t **dev_ptr)
{
    if (!c || !dev_ptr) return VKRT_BAD_ARG;
    if (!c->d_hit_ids) return fail(c, VKRT_BAD_ARG, "context created without VKRT_FLAG_HIT_IDS");
    { DeviceGuard g(c->info.device_id); CU(c, join(c)); }
    *dev_ptr = c->d_hit_ids;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_get_stream(vkrt_ctx *c, void **stream)
{
    if (!c || !stream) return VKRT_BAD_ARG;
    *stream = (void *)c->stream;
    return VKRT_SUCCESS;
}

static vkrt_error read_back(vkrt_ctx *c, void *host, const void *dev, size_t have, size_t bytes)
{
    if (!host || !dev) return fail(c, VKRT_BAD_ARG, "null buffer");
    if (bytes < have) return fail(c, VKRT_BAD_ARG, "host buffer too small");
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    CU(c, cudaMemcpyAsync(host, dev, have, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_read_rgba8_async(vkrt_ctx *c, void *host, size_t bytes)
{
    if (!c || !host) return VKRT_BAD_ARG;
    const size_t row = (size_t)c->info.width * 4, have = row * c->info.height;
    if (bytes < have) return fail(c, VKRT_BAD_ARG, "host buffer too small");
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    const ExtSlot &e = c->ext[c->cur_target];
    if (e.kind == 2) CU(c, cudaMemcpy2DFromArrayAsync(host, row, e.arr, 0, 0, row, c->info.height, cudaMemcpyDeviceToHost, c->stream));
    else if (e.kind == 1) CU(c, cudaMemcpy2DAsync(host, row, e.ptr, e.pitch, row, c->info.height, cudaMemcpyDeviceToHost, c->stream));
    else CU(c, cudaMemcpyAsync(host, c->d_rgba[c->cur_target], have, cudaMemcpyDeviceToHost, c->stream));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_read_rgba8(vkrt_ctx *c, void *host, size_t bytes)
{
    if (!c) return VKRT_BAD_ARG;
    if (!host) return fail(c, VKRT_BAD_ARG, "null buffer");
    const vkrt_error r = vkrt_read_rgba8_async(c, host, bytes);
    if (r != VKRT_SUCCESS) return r;
    DeviceGuard g(c->info.device_id);
    CU(c, cudaStreamSynchronize(c->stream));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_read_accum(vkrt_ctx *c, float *host, size_t bytes)
{
    if (!c) return VKRT_BAD_ARG;
    return read_back(c, host, c->d_accum, (size_t)c->info.width * c->info.height * 16, bytes);
}
VKRT_API vkrt_error vkrt_read_hit_ids(vkrt_ctx *c, uint32_t *host, size_t bytes)
{
    if (!c) return VKRT_BAD_ARG;
    if (!c->d_hit_ids) return fail(c, VKRT_BAD_ARG, "context created without VKRT_FLAG_HIT_IDS");
    return read_back(c, host, c->d_hit_ids, (size_t)c->info.width * c->info.height * 4, bytes);
}

VKRT_API vkrt_error vkrt_get_counters(vkrt_ctx *c, vkrt_counters *out)
{
    if (!c || !out) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    unsigned long long h[CNT_N];
    CU(c, join(c));
    CU(c, cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    out->closest_rays = h[CNT_CLOSEST]; out->shadow_rays = h[CNT_SHADOW]; out->node_visits = h[CNT_NODES];
    out->leaf_tests = h[CNT_LEAVES]; out->paths = h[CNT_PATHS]; out->frames = c->frames;
    out->shared_primary_rays = h[CNT_SHARED]; out->zero_term_shadow_rays = h[CNT_SKIPPED];
    for (vkrt_ctx *ch : c->children) {           // one frame, several devices: the rays add up, the frames do not
        vkrt_counters o;
        const vkrt_error r = vkrt_get_counters(ch, &o);
        if (r != VKRT_SUCCESS) return fail(c, r, ch->err);
        out->closest_rays += o.closest_rays; out->shadow_rays += o.shadow_rays; out->node_visits += o.node_visits;
        out->leaf_tests += o.leaf_tests; out->paths += o.paths;
        out->shared_primary_rays += o.shared_primary_rays; out->zero_term_shadow_rays += o.zero_term_shadow_rays;
    }
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_reset_counters(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    CU(c, cudaMemsetAsync(c->d_counters, 0, CNT_N * sizeof(unsigned long long), c->stream));
    c->frames = 0;
    FWD(c, vkrt_reset_counters(ch));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_last_frame_timing(vkrt_ctx *c, float *trace_ms, float *total_ms, uint32_t *n_launches)
{
    if (!c) return VKRT_BAD_ARG;
    if (!c->timing_valid) return fail(c, VKRT_BAD_ARG, "no frame drawn yet");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaEventSynchronize(c->ev_end));
    float a = 0.f, b = 0.f;
    CU(c, cudaEventElapsedTime(&a, c->ev_begin, c->ev_trace1));
    CU(c, cudaEventElapsedTime(&b, c->ev_begin, c->ev_end));
    if (trace_ms) *trace_ms = a;
    if (total_ms) *total_ms = b;
    if (n_launches) *n_launches = c->last_launches;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_last_frame_traversal_timing(vkrt_ctx *c, float *traversal_ms, uint32_t *n_launches, float *all_kernels_ms)
{
    if (!c) return VKRT_BAD_ARG;
    if (!c->timing_valid) return fail(c, VKRT_BAD_ARG, "no frame drawn yet");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaEventSynchronize(c->ev_end));
    float sum = 0.f, all = 0.f; uint32_t n = 0;
    if (c->info.integrator == VKRT_INTEGRATOR_PATH && c->info.variant == VKRT_VARIANT_WAVEFRONT && c->wave_ready) {
        if (!(c->info.flags & VKRT_FLAG_LAUNCH_TIMING))
            return fail(c, VKRT_BAD_ARG, "per-launch timing needs a context created with VKRT_FLAG_LAUNCH_TIMING");
        for (uint32_t l = 0; l < c->wave.n_lanes; ++l)
            for (uint32_t i = 0; i + 1 < c->wave.lane[l].n_ev; i += 2) {
                const uint8_t tag = c->wave.lane[l].ev_tag[i / 2];
                float ms = 0.f;
                CU(c, cudaEventElapsedTime(&ms, c->wave.lane[l].ev[i], c->wave.lane[l].ev[i + 1]));
                all += ms;
                if (tag == 1 || tag == 3) { sum += ms; ++n; }      // extend / shadow launches
            }
    } else {
        CU(c, cudaEventElapsedTime(&sum, c->ev_begin, c->ev_trace1));
        CU(c, cudaEventElapsedTime(&all, c->ev_begin, c->ev_end));
        n = 1;
    }
    if (traversal_ms) *traversal_ms = sum;
    if (n_launches) *n_launches = n;
    if (all_kernels_ms) *all_kernels_ms = all;
    return VKRT_SUCCESS;
}

/* Diagnostics: writes "lane kernel start_ms end_ms" for every launch of the most recent wavefront frame
 * (times relative to the frame's first launch) -- a poor man's timeline, there is no nsys in the image. */
VKRT_API vkrt_error vkrt_debug_dump_timeline(vkrt_ctx *c, const char *path)
{
    if (!c || !path) return VKRT_BAD_ARG;
    if (!c->timing_valid || !c->wave_ready) return fail(c, VKRT_BAD_ARG, "no wavefront frame drawn yet");
    if (!(c->info.flags & VKRT_FLAG_LAUNCH_TIMING)) return fail(c, VKRT_BAD_ARG, "the timeline needs a context created with VKRT_FLAG_LAUNCH_TIMING");
    DeviceGuard g(c->info.device_id);
    CU(c, cudaEventSynchronize(c->ev_end));
    FILE *f = std::fopen(path, "w");
    if (!f) return fail(c, VKRT_BAD_ARG, "cannot open timeline file");
    static const char *names[6] = {"generate", "extend", "classify", "shadow", "shade", "reduce"};
    for (uint32_t l = 0; l < c->wave.n_lanes; ++l)
        for (uint32_t i = 0; i + 1 < c->wave.lane[l].n_ev; i += 2) {
            float a = 0.f, b = 0.f;
            if (cudaEventElapsedTime(&a, c->ev_begin, c->wave.lane[l].ev[i]) != cudaSuccess ||
                cudaEventElapsedTime(&b, c->ev_begin, c->wave.lane[l].ev[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
            std::fprintf(f, "%u %s %.4f %.4f\n", l, names[c->wave.lane[l].ev_tag[i / 2] % 6], a, b);
        }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, c->ev_begin, c->ev_end);
    std::fprintf(f, "frame total %.4f\n", tot);
    std::fclose(f);
    return VKRT_SUCCESS;
}

/* Assets/Fullscreen.frag on the traced images: binding 0 = image slot 0, binding 1 = image slot 1 -- fixed, as the
 * reference writes its graphics descriptor set once (Source/GraphicsDevice.cpp:964-983) while the compute pass
 * alternates its target, so "current" and "previous" swap roles every other frame. */
VKRT_API vkrt_error vkrt_present(vkrt_ctx *c, void *host_rgba8, uint32_t out_w, uint32_t out_h)
{
    if (!c || !host_rgba8 || out_w == 0 || out_h == 0) return VKRT_BAD_ARG;
    if (c->d_rgba.size() < 2) return fail(c, VKRT_BAD_ARG, "the present filter needs frames_in_flight >= 2");
    if (c->ext[0].kind != 0 || c->ext[1].kind != 0)
        return fail(c, VKRT_BAD_ARG, "image slots 0/1 are external targets: the engine's own Fullscreen pass presents them");
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    const size_t bytes = (size_t)out_w * out_h * 4;
    if (c->present_bytes < bytes) {
        cudaFree(c->d_present); c->d_present = nullptr; c->present_bytes = 0;
        CU(c, cudaMalloc(&c->d_present, bytes));
        c->present_bytes = bytes;
    }
    CU(c, launch_present(c->d_rgba[0], c->d_rgba[1], c->info.width, c->info.height, c->d_present, out_w, out_h, c->stream));
    CU(c, cudaMemcpyAsync(host_rgba8, c->d_present, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return VKRT_SUCCESS;
}

// ---- Vulkan <-> CUDA interop (include/vkrt.h) -------------------------------------------------------
VKRT_API vkrt_error vkrt_bind_rgba8_target(vkrt_ctx *c, uint32_t slot, void *dev_ptr, size_t row_pitch)
{
    if (!c) return VKRT_BAD_ARG;
    if (slot >= c->ext.size()) return fail(c, VKRT_BAD_ARG, "slot >= frames_in_flight");
    const size_t row = (size_t)c->info.width * 4;
    if (row_pitch == 0) row_pitch = row;
    if (dev_ptr && (row_pitch < row || (row_pitch & 3u) || ((uintptr_t)dev_ptr & 3u)))
        return fail(c, VKRT_BAD_ARG, "row_pitch < width * 4, or pointer / pitch not 4-byte aligned");
    DeviceGuard g(c->info.device_id);
    if (dev_ptr) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, dev_ptr) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
            cudaGetLastError();
            return fail(c, VKRT_BAD_ARG, "dev_ptr is not device memory");
        }
    }
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    ExtSlot &e = c->ext[slot];
    ext_release_target(e);
    if (dev_ptr) { e.kind = 1; e.ptr = (uchar4 *)dev_ptr; e.pitch = row_pitch; }
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_debug_bind_array_target(vkrt_ctx *c, uint32_t slot)
{
    if (!c) return VKRT_BAD_ARG;
    if (slot >= c->ext.size()) return fail(c, VKRT_BAD_ARG, "slot >= frames_in_flight");
    DeviceGuard g(c->info.device_id);
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    ExtSlot &e = c->ext[slot];
    ext_release_target(e);
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CU(c, cudaMallocArray(&e.arr, &fmt, c->info.width, c->info.height, cudaArraySurfaceLoadStore));
    e.own_arr = true;
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = e.arr;
    const cudaError_t ce = cudaCreateSurfaceObject(&e.surf, &rd);
    if (ce != cudaSuccess) { ext_release_target(e); return cuda_fail(c, ce, "cudaCreateSurfaceObject"); }
    e.kind = 2;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_import_vk_image(vkrt_ctx *c, uint32_t slot, const vkrt_external_image *im)
{
    if (!c || !im) return VKRT_BAD_ARG;
    if (im->struct_size != sizeof(vkrt_external_image)) return fail(c, VKRT_BAD_ARG, "vkrt_external_image.struct_size mismatch");
    if (slot >= c->ext.size()) return fail(c, VKRT_BAD_ARG, "slot >= frames_in_flight");
    if (im->fd < 0) return fail(c, VKRT_BAD_ARG, "bad file descriptor");
    if (im->tiling > VKRT_TILING_OPTIMAL) return fail(c, VKRT_BAD_ARG, "unknown tiling");
    const size_t row = (size_t)c->info.width * 4;
    const size_t pitch = im->row_pitch ? im->row_pitch : row;
    if (im->tiling == VKRT_TILING_LINEAR && (pitch < row || (pitch & 3u) || (im->offset & 3u)))
        return fail(c, VKRT_BAD_ARG, "row_pitch < width * 4, or offset / pitch not 4-byte aligned");
    const uint64_t need = im->tiling == VKRT_TILING_LINEAR ? (uint64_t)pitch * c->info.height : (uint64_t)row * c->info.height;
    if (im->allocation_size == 0 || im->offset > im->allocation_size || im->allocation_size - im->offset < need)
        return fail(c, VKRT_BAD_ARG, "the image does not fit in the allocation (allocation_size / offset)");
    DeviceGuard g(c->info.device_id);
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;

    ExtSlot n;                                      // built aside: a failed import leaves the slot as it was
    const int own = dup_fd(im->fd);
    if (own < 0) return fail(c, VKRT_BAD_ARG, "bad file descriptor");
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = own;
    hd.size = im->allocation_size;
    hd.flags = im->dedicated ? cudaExternalMemoryDedicated : 0;
    cudaError_t ce = cudaImportExternalMemory(&n.mem, &hd);
    if (ce != cudaSuccess) { cudaGetLastError(); close_if_open(own); return cuda_fail(c, ce, "cudaImportExternalMemory"); }
    if (im->tiling == VKRT_TILING_LINEAR) {
        cudaExternalMemoryBufferDesc bd{};
        bd.offset = im->offset; bd.size = need;
        void *p = nullptr;
        ce = cudaExternalMemoryGetMappedBuffer(&p, n.mem, &bd);
        if (ce == cudaSuccess) { n.kind = 1; n.ptr = (uchar4 *)p; n.pitch = pitch; }
    } else {
        cudaExternalMemoryMipmappedArrayDesc md{};
        md.offset = im->offset;
        md.formatDesc = cudaCreateChannelDesc<uchar4>();           // VK_FORMAT_R8G8B8A8_UNORM (:672)
        md.extent = make_cudaExtent(c->info.width, c->info.height, 0);
        md.flags = cudaArraySurfaceLoadStore;                       // VK_IMAGE_USAGE_STORAGE_BIT (:674)
        md.numLevels = 1;                                           // mipLevels = 1 (:685)
        ce = cudaExternalMemoryGetMappedMipmappedArray(&n.mip, n.mem, &md);
        if (ce == cudaSuccess) ce = cudaGetMipmappedArrayLevel(&n.arr, n.mip, 0);
        if (ce == cudaSuccess) {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeArray; rd.res.array.array = n.arr;
            ce = cudaCreateSurfaceObject(&n.surf, &rd);
        }
        if (ce == cudaSuccess) n.kind = 2;
    }
    if (ce != cudaSuccess) {
        cudaGetLastError();
        ext_release_target(n);                      // also closes the duplicate: the import itself had succeeded
        return cuda_fail(c, ce, "mapping the imported allocation");
    }
    ExtSlot &e = c->ext[slot];
    ext_release_target(e);
    e.kind = n.kind; e.mem = n.mem; e.mip = n.mip; e.arr = n.arr; e.own_arr = false; e.surf = n.surf; e.ptr = n.ptr; e.pitch = n.pitch;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_import_vk_semaphore(vkrt_ctx *c, uint32_t slot, uint32_t which, int32_t fd, uint32_t timeline)
{
    if (!c) return VKRT_BAD_ARG;
    if (slot >= c->ext.size()) return fail(c, VKRT_BAD_ARG, "slot >= frames_in_flight");
    if (which > VKRT_SEMAPHORE_RELEASE) return fail(c, VKRT_BAD_ARG, "unknown semaphore role");
    if (fd < 0) return fail(c, VKRT_BAD_ARG, "bad file descriptor");
    DeviceGuard g(c->info.device_id);
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    const int own = dup_fd(fd);
    if (own < 0) return fail(c, VKRT_BAD_ARG, "bad file descriptor");
    cudaExternalSemaphoreHandleDesc sd{};
    sd.type = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    sd.handle.fd = own;
    cudaExternalSemaphore_t sem = nullptr;
    const cudaError_t ce = cudaImportExternalSemaphore(&sem, &sd);
    if (ce != cudaSuccess) { cudaGetLastError(); close_if_open(own); return cuda_fail(c, ce, "cudaImportExternalSemaphore"); }
    ExtSlot &e = c->ext[slot];
    // the frame count both timelines are derived from starts with the slot's FIRST semaphore; importing the second
    // one (or replacing one) later does not move it back
    if (!e.sem[VKRT_SEMAPHORE_ACQUIRE] && !e.sem[VKRT_SEMAPHORE_RELEASE]) e.frames = 0;
    if (e.sem[which]) cudaDestroyExternalSemaphore(e.sem[which]);
    e.sem[which] = sem;
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_release_external(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    for (auto &e : c->ext) { ext_release_target(e); ext_release_semaphores(e); }
    return VKRT_SUCCESS;
}

// ---- sharding -------------------------------------------------------------------------------------
VKRT_API vkrt_error vkrt_shard_floats(vkrt_ctx *c, uint32_t tile_rank, size_t *n_floats)
{
    if (!c || !n_floats) return VKRT_BAD_ARG;
    const uint32_t cnt = c->info.tile_shard_count;
    const uint32_t owned = tile_rank < c->n_tiles ? (c->n_tiles - tile_rank + cnt - 1) / cnt : 0;
    *n_floats = (size_t)owned * TILE_PX * 4;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_pack_shard(vkrt_ctx *c, float **dev_ptr, size_t *n_floats)
{
    if (!c || !dev_ptr) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    // every rank packs the size of rank 0 (the largest shard) so that a plain gather works
    const uint32_t cnt = c->info.tile_shard_count;
    const size_t max_slots = (size_t)((c->n_tiles + cnt - 1) / cnt) * TILE_PX;
    if (!c->d_packed) {
        CU(c, cudaMalloc(&c->d_packed, max_slots * sizeof(float4)));
        CU(c, cudaMemsetAsync(c->d_packed, 0, max_slots * sizeof(float4), c->stream));
    }
    CU(c, join(c));
    RenderParams rp;
    fill_params(c, rp);
    if (rp.n_work) CU(c, launch_pack(rp, c->d_packed, c->stream));
    *dev_ptr = (float *)c->d_packed;
    if (n_floats) *n_floats = max_slots * 4;
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_pack_shard_into(vkrt_ctx *c, float *dev_dst, size_t n_floats)
{
    if (!c || !dev_dst) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    RenderParams rp;
    fill_params(c, rp);
    if (n_floats < (size_t)rp.n_work * 4) return fail(c, VKRT_BAD_ARG, "shard buffer too small");
    if (rp.n_work) CU(c, launch_pack(rp, (float4 *)dev_dst, c->stream));
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_unpack_shard(vkrt_ctx *c, const float *dev_packed, uint32_t tile_rank, uint32_t tile_count, int add)
{
    if (!c || !dev_packed || tile_count == 0 || tile_rank >= tile_count) return VKRT_BAD_ARG;
    DeviceGuard g(c->info.device_id);
    CU(c, join(c));
    CU(c, launch_unpack(c->d_accum, (const float4 *)dev_packed, c->info.width, c->info.height, tile_rank, tile_count, add, c->stream));
    c->accum_valid = true;
    return VKRT_SUCCESS;
}

// ---- mesh loading (ref: the TODO at Assets/Raytracer.comp:10, "Load mesh data off of disk and upload ... triangle lists") ----
// Wavefront OBJ, the subset a triangle list needs: `v x y z` and `f a b c ...` (1-based or negative indices, an optional
// /vt/vn suffix per corner is ignored, polygons are fanned around their first corner).  Context-free: the engine hands
// the triangles to vkrt_set_triangles like the reference hands its one triangle to the SSBO (Source/GraphicsDevice.cpp:796-822).
VKRT_API vkrt_error vkrt_load_obj(const char *path, const float xform[12], vkrt_triangle *out, uint32_t capacity, uint32_t *n_triangles)
{
    if (!path || !n_triangles) return VKRT_BAD_ARG;
    *n_triangles = 0;
    FILE *f = std::fopen(path, "r");
    if (!f) return fail(nullptr, VKRT_BAD_ARG, std::string("cannot open ") + path);
    std::vector<float> v;
    uint32_t n = 0;
    char line[1024];
    bool bad = false;
    while (std::fgets(line, sizeof line, f)) {
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            float x, y, z;
            if (std::sscanf(line + 2, "%f %f %f", &x, &y, &z) != 3) { bad = true; break; }
            if (xform) {                       // row-major 3x4: p' = M p + t
                const float X = xform[0] * x + xform[1] * y + xform[2] * z + xform[3];
                const float Y = xform[4] * x + xform[5] * y + xform[6] * z + xform[7];
                const float Z = xform[8] * x + xform[9] * y + xform[10] * z + xform[11];
                x = X; y = Y; z = Z;
            }
            v.push_back(x); v.push_back(y); v.push_back(z);
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            long idx[64]; int k = 0;
            const long nv = (long)(v.size() / 3);
            for (char *p = line + 2; *p && k < 64; ) {
                while (*p == ' ' || *p == '\t') ++p;
                if (!*p || *p == '\n' || *p == '\r') break;
                char *end = nullptr;
                long i = std::strtol(p, &end, 10);
                if (end == p) { bad = true; break; }
                i = i < 0 ? nv + i : i - 1;
                if (i < 0 || i >= nv) { bad = true; break; }
                idx[k++] = i;
                p = end;
                while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;      // skip /vt/vn
            }
            if (bad || k < 3) { bad = true; break; }
            for (int j = 1; j + 1 < k; ++j, ++n) {
                if (out && n < capacity) {
                    const long t3[3] = {idx[0], idx[j], idx[j + 1]};
                    vkrt_vec3a *dst[3] = {&out[n].v0, &out[n].v1, &out[n].v2};
                    for (int q = 0; q < 3; ++q) { dst[q]->x = v[3 * t3[q]]; dst[q]->y = v[3 * t3[q] + 1]; dst[q]->z = v[3 * t3[q] + 2]; dst[q]->_pad = 0.f; }
                }
            }
        }
    }
    std::fclose(f);
    if (bad) return fail(nullptr, VKRT_BAD_ARG, std::string("malformed OBJ: ") + path);
    *n_triangles = n;                          // the number in the file: call again with a larger buffer if it exceeds `capacity`
    return VKRT_SUCCESS;
}

// ---- frame exchange over peer memory (csrc/vkrt_exchange.cu) ---------------------------------------
static vkrt_error exchange_bind(vkrt_ctx *c, char *base, size_t bytes, bool owner, bool ipc)
{
    c->xch = vkrt_ctx::Exchange{};
    c->xch.base = base; c->xch.bytes = bytes; c->xch.owner = owner; c->xch.ipc = ipc;
    c->xch.T = c->info.tile_shard_count; c->xch.S = c->info.sample_shard_count;
    c->xch.rank = c->info.sample_shard_rank * c->info.tile_shard_count + c->info.tile_shard_rank;
    return VKRT_SUCCESS;
}
static size_t exchange_bytes(const vkrt_ctx *c)
{
    return (size_t)X_OFF_TARGETS + (size_t)2 * c->info.sample_shard_count * c->info.width * c->info.height * sizeof(float4);
}
VKRT_API vkrt_error vkrt_exchange_create(vkrt_ctx *c, vkrt_exchange_handle *out)
{
    if (!c || !out) return VKRT_BAD_ARG;
    if (c->info.tile_shard_rank != 0 || c->info.sample_shard_rank != 0)
        return fail(c, VKRT_BAD_ARG, "the exchange block belongs to the gathering context (tile rank 0, sample rank 0)");
    if ((uint64_t)c->info.tile_shard_count * c->info.sample_shard_count > X_MAX_RANKS) return fail(c, VKRT_BAD_ARG, "more than 64 ranks");
    if (c->xch.base) return fail(c, VKRT_BAD_ARG, "an exchange is already attached");
    DeviceGuard g(c->info.device_id);
    vkrt_error r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    const size_t bytes = exchange_bytes(c);
    char *base = nullptr;
    CU(c, cudaMalloc((void **)&base, bytes));
    cudaError_t e = cudaMemset(base, 0, bytes);
    std::memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) <= sizeof(out->ipc), "cudaIpcMemHandle_t is 64 bytes");
    // the handle is only needed by other processes; a device without IPC support still serves vkrt_exchange_attach
    if (e == cudaSuccess && cudaIpcGetMemHandle(&h, base) == cudaSuccess) std::memcpy(out->ipc, &h, sizeof h);
    else cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(base); return cuda_fail(c, e, "cudaMemset(exchange block)"); }
    out->bytes = bytes; out->width = c->info.width; out->height = c->info.height;
    out->tile_shard_count = c->info.tile_shard_count; out->sample_shard_count = c->info.sample_shard_count; out->magic = X_MAGIC;
    return exchange_bind(c, base, bytes, true, false);
}
static vkrt_error exchange_check_peer(vkrt_ctx *c, uint32_t w, uint32_t h, uint32_t T, uint32_t S)
{
    if (c->xch.base) return fail(c, VKRT_BAD_ARG, "an exchange is already attached");
    if (w != c->info.width || h != c->info.height || T != c->info.tile_shard_count || S != c->info.sample_shard_count)
        return fail(c, VKRT_BAD_ARG, "the exchange block was made for another frame size / shard layout");
    if (c->info.tile_shard_rank == 0 && c->info.sample_shard_rank == 0)
        return fail(c, VKRT_BAD_ARG, "rank (0, 0) is the gathering context: it creates the exchange block");
    if ((c->info.flags & VKRT_FLAG_PROGRESSIVE) != 0) return fail(c, VKRT_BAD_ARG, "VKRT_FLAG_PROGRESSIVE belongs to the gathering context only");
    return VKRT_SUCCESS;
}
VKRT_API vkrt_error vkrt_exchange_open(vkrt_ctx *c, const vkrt_exchange_handle *hd)
{
    if (!c || !hd) return VKRT_BAD_ARG;
    if (hd->magic != X_MAGIC) return fail(c, VKRT_BAD_ARG, "not an exchange handle");
    vkrt_error r = exchange_check_peer(c, hd->width, hd->height, hd->tile_shard_count, hd->sample_shard_count);
    if (r != VKRT_SUCCESS) return r;
    if (hd->bytes != exchange_bytes(c)) return fail(c, VKRT_BAD_ARG, "exchange block size mismatch");
    DeviceGuard g(c->info.device_id);
    r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hd->ipc, sizeof h);
    void *base = nullptr;
    CU(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    return exchange_bind(c, (char *)base, (size_t)hd->bytes, false, true);
}
VKRT_API vkrt_error vkrt_exchange_attach(vkrt_ctx *c, vkrt_ctx *owner)
{
    if (!c || !owner) return VKRT_BAD_ARG;
    if (!owner->xch.base || !owner->xch.owner) return fail(c, VKRT_BAD_ARG, "the gathering context has no exchange block (vkrt_exchange_create)");
    vkrt_error r = exchange_check_peer(c, owner->info.width, owner->info.height, owner->info.tile_shard_count, owner->info.sample_shard_count);
    if (r != VKRT_SUCCESS) return r;
    DeviceGuard g(c->info.device_id);
    r = make_idle(c);
    if (r != VKRT_SUCCESS) return r;
    if (c->info.device_id != owner->info.device_id) {
        int can = 0;
        CU(c, cudaDeviceCanAccessPeer(&can, c->info.device_id, owner->info.device_id));
        if (!can) return fail(c, VKRT_NO_SUITABLE_GPU, "no peer access between the devices (NVLink / PCIe P2P)");
        const cudaError_t e = cudaDeviceEnablePeerAccess(owner->info.device_id, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(c, e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    return exchange_bind(c, owner->xch.base, owner->xch.bytes, false, false);
}
VKRT_API vkrt_error vkrt_exchange_close(vkrt_ctx *c)
{
    if (!c) return VKRT_BAD_ARG;
    if (!c->xch.base) return VKRT_SUCCESS;
    DeviceGuard g(c->info.device_id);
    cudaDeviceSynchronize();
    cudaGetLastError();
    if (c->xch.owner) cudaFree(c->xch.base);
    else if (c->xch.ipc) cudaIpcCloseMemHandle(c->xch.base);
    c->xch = vkrt_ctx::Exchange{};
    return VKRT_SUCCESS;
}

VKRT_API vkrt_error vkrt_measure_fp32_peak(int device_id, float *tflops)
{
    if (!tflops) return VKRT_BAD_ARG;
    DeviceGuard g(device_id);
    return measure_fp32_peak(tflops) == cudaSuccess ? VKRT_SUCCESS : VKRT_CUDA_ERROR;
}
VKRT_API vkrt_error vkrt_measure_l2_bandwidth(int device_id, float *gbs)
{
    if (!gbs) return VKRT_BAD_ARG;
    DeviceGuard g(device_id);
    return measure_l2_bandwidth(gbs) == cudaSuccess ? VKRT_SUCCESS : VKRT_CUDA_ERROR;
}

} // extern "C"
