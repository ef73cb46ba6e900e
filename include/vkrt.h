/*
 * libvkrt_cuda -- C ABI of the B200-native ray-tracing hot path for vk-renderer.
 *
 * This is the drop-in boundary: it replaces what the reference does between
 * Source/GraphicsDevice.cpp:1234 and :1284 (barrier -> bind compute pipeline ->
 * push 96-byte FrameData -> vkCmdDispatch(RES/16, RES/16, 1) -> barrier), i.e.
 * the execution of Assets/Tracer.comp / Assets/Raytracer.comp, plus the scene
 * upload of Source/GraphicsDevice.cpp:796-822.
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All functions return a
 * vkrt_error; no exception crosses this boundary.  One host thread per context
 * (the reference is single-threaded too, Source/VulkanState.h:52-55).
 *
 * Citations "ref:" are relative to the reference repository root.
 */
#ifndef VKRT_H
#define VKRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKRT_API __declspec(dllexport)
#else
#define VKRT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------- */
/* Error convention -- ref: Include/GraphicsDevice.h:46-52                     */
/* (enum class Error : signed char { SUCCESS, NO_SUITABLE_GPU,                */
/*  NO_SUITABLE_SURFACE, UNKNOWN }); the first four values are identical.     */
/* ------------------------------------------------------------------------- */
typedef int8_t vkrt_error;
enum {
    VKRT_SUCCESS             = 0,
    VKRT_NO_SUITABLE_GPU     = 1,
    VKRT_NO_SUITABLE_SURFACE = 2, /* never produced; kept for value parity */
    VKRT_UNKNOWN             = 3,
    VKRT_CUDA_ERROR          = 4,
    VKRT_NCCL_ERROR          = 5, /* the multi-GPU frame exchange failed (a peer did not deliver its frame in time) */
    VKRT_BAD_ARG             = 6
};

/* ------------------------------------------------------------------------- */
/* Boundary PODs, byte-identical to the reference's host structs and to the   */
/* shader's push-constant / std430 blocks.                                    */
/* ------------------------------------------------------------------------- */

/* a glm::vec3 with alignas(16): 12 bytes of data + 4 bytes of padding */
typedef struct vkrt_vec3a { float x, y, z, _pad; } vkrt_vec3a;

/* ref: Include/Camera.h:5-12 (CameraData, 64 B); shader: Tracer.comp:22-35 */
typedef struct vkrt_camera_data {
    vkrt_vec3a pos;   /* @0  */
    vkrt_vec3a dir;   /* @16 */
    vkrt_vec3a right; /* @32 */
    vkrt_vec3a up;    /* @48 */
} vkrt_camera_data;

/* ref: Include/GraphicsDevice.h:20-29 (FrameData, 96 B);
 * shader: Tracer.comp:151-166 (offsets 0/4/16/32), Raytracer.comp:79-86 */
typedef struct vkrt_frame_data {
    float            aspect_ratio; /* @0  */
    float            seed;         /* @4  */
    float            _pad0[2];     /* @8  */
    vkrt_vec3a       light_pos;    /* @16 (used by the whitted integrator only) */
    vkrt_camera_data camera;       /* @32 */
} vkrt_frame_data;

/* ref: Include/GraphicsDevice.h:13-18 (Triangle, 48 B); shader: Tracer.comp:127-132,146-149 */
typedef struct vkrt_triangle {
    vkrt_vec3a v0, v1, v2;
} vkrt_triangle;

/* ref: Tracer.comp:59-74 (Material).  In the reference these are shader
 * constants (Tracer.comp:186-194); here they are a table indexed by mat_id.
 * The whitted integrator (Raytracer.comp:59-63, Material{bool reflective;
 * vec3 diffuse}) reads albedo as `diffuse` and (metalness >= 0.5) as `reflective`. */
enum { VKRT_MAT_DIFFUSE = 0, VKRT_MAT_DIELECTRIC = 1 };
typedef struct vkrt_material {
    float    albedo[3];
    float    roughness;   /* DIELECTRIC: used raw as the refraction ratio (Tracer.comp:532,537,539) */
    float    emissive[3];
    float    metalness;
    uint32_t type;        /* VKRT_MAT_* */
    uint32_t _pad[3];
} vkrt_material; /* 48 B */

/* ref: Tracer.comp:102-111 (Sphere: P, r) */
typedef struct vkrt_sphere { float cx, cy, cz, r; } vkrt_sphere;
/* ref: Tracer.comp:118-125 (Plane: dot(P,N) + len = 0) */
typedef struct vkrt_plane { float nx, ny, nz, len; } vkrt_plane;

/* ------------------------------------------------------------------------- */
/* Context creation -- generalises GraphicsDevice::CreateInfo                  */
/* (ref: Include/GraphicsDevice.h:57-72; Main.cpp:105-115 passes {3,2,1024,0}) */
/* ------------------------------------------------------------------------- */
enum { VKRT_INTEGRATOR_WHITTED = 0,   /* Assets/Raytracer.comp semantics */
       VKRT_INTEGRATOR_PATH    = 1 }; /* Assets/Tracer.comp semantics    */
enum { VKRT_VARIANT_MEGAKERNEL = 0,   /* persistent threads, per-lane path regeneration */
       VKRT_VARIANT_WAVEFRONT  = 1 }; /* raygen / extend / shade / compact queues       */
enum { VKRT_SCENE_TRACER = 0,         /* Tracer.comp:186-211    */
       VKRT_SCENE_RAYTRACER = 1 };    /* Raytracer.comp:98-127  */

enum {
    VKRT_FLAG_PROGRESSIVE = 1u << 0, /* keep summing frames into the accumulator (new; the
                                        reference renders every frame from scratch)       */
    VKRT_FLAG_HIT_IDS     = 1u << 1, /* write the primary nearest-hit id AOV               */
    VKRT_FLAG_STATS       = 1u << 2, /* also count BVH node visits / leaf tests            */
    VKRT_FLAG_NO_RESOLVE  = 1u << 3, /* skip the rgba8 resolve in vkrt_draw (shard ranks)  */
    VKRT_FLAG_SERIAL_WAVES = 1u << 4, /* measurement aid (wavefront): the same waves and launches, but one after the other on one
                                        stream instead of overlapping on two -- per-launch event times then belong to one kernel
                                        alone (bench.py's roofline figures); the image is bit-identical */
    VKRT_FLAG_LAUNCH_TIMING = 1u << 5 /* measurement aid (wavefront): bracket every kernel launch with a CUDA-event pair
                                        (vkrt_last_frame_traversal_timing, vkrt_debug_dump_timeline); off = no event in the hot path */
};

typedef struct vkrt_create_info {
    uint32_t struct_size;        /* = sizeof(vkrt_create_info) */
    uint32_t width, height;      /* reference: raytrace_resolution x raytrace_resolution */
    uint32_t spp;                /* reference: SAMPLES = 4   (Tracer.comp:180) */
    uint32_t max_depth;          /* reference: DEPTH = 4     (Tracer.comp:179); whitted: BOUNCES = 2 (Raytracer.comp:90) */
    uint32_t integrator;         /* VKRT_INTEGRATOR_* */
    uint32_t variant;            /* VKRT_VARIANT_*    */
    uint32_t frames_in_flight;   /* reference: 2 (Main.cpp:110); number of rgba8 targets */
    int32_t  device_id;          /* CUDA device ordinal */
    uint32_t flags;              /* VKRT_FLAG_* */
    /* screen-tile sharding (tile = 32x32 px, owner = tile_index % tile_shard_count) */
    uint32_t tile_shard_rank, tile_shard_count;
    /* sample-range sharding: this context renders samples
     * [sample_shard_rank*spp/count, (sample_shard_rank+1)*spp/count) of each frame */
    uint32_t sample_shard_rank, sample_shard_count;
    void    *stream;             /* cudaStream_t to launch on; NULL = library-owned stream */
    /* In-library multi-GPU (ref: the reference is single-GPU, Source/VulkanState.h:52-55; GraphicsDevice::Draw stays ONE call):
     * n_devices > 1 makes this one context render every frame on the CUDA devices device_ids[0 .. n_devices), sharded by
     * interleaved screen tiles; device_ids[0] gathers (every other GPU stores its pixels straight into its memory over
     * NVLink, "frame exchange" below) and owns the outputs.  n_devices <= 1: device_id is used, as before.  The shard
     * fields above must then be 0 / 1 -- the library shards by itself. */
    uint32_t n_devices;
    int32_t  device_ids[8];
} vkrt_create_info;

typedef struct vkrt_ctx vkrt_ctx;

/* ref: GraphicsDevice::Construct (Include/GraphicsDevice.h:83) */
VKRT_API vkrt_error vkrt_create(const vkrt_create_info *info, vkrt_ctx **out_ctx);
/* ref: GraphicsDevice::Destruct (Include/GraphicsDevice.h:92) */
VKRT_API vkrt_error vkrt_destroy(vkrt_ctx *ctx);
/* ref: GraphicsDevice::Draw (Include/GraphicsDevice.h:94, Source/GraphicsDevice.cpp:1215).
 * Takes the same 96-byte struct.  Unlike the reference's Draw it does NOT overwrite
 * aspect_ratio / seed (Source/GraphicsDevice.cpp:1260-1262); the GraphicsDevice shim does.
 * Asynchronous: enqueues on the context's stream and returns. */
VKRT_API vkrt_error vkrt_draw(vkrt_ctx *ctx, const vkrt_frame_data *frame);
/* ref: GraphicsDevice::WaitIdle (Include/GraphicsDevice.h:96) */
VKRT_API vkrt_error vkrt_wait_idle(vkrt_ctx *ctx);
/* Like the reference (FRAMES_IN_FLIGHT = 2, Draw only waits for the fence of the slot it reuses,
 * Source/GraphicsDevice.cpp:1217), consecutive vkrt_draw calls overlap on the device.  Every library call that
 * consumes a frame waits for it by itself; vkrt_flush makes the context's stream wait for the frames in flight
 * (device-side, no host sync) for callers that enqueue their own work or events on that stream. */
VKRT_API vkrt_error vkrt_flush(vkrt_ctx *ctx);

/* Runtime knobs that are compile-time constants in the shader (Tracer.comp:179-180). */
VKRT_API vkrt_error vkrt_set_sampling(vkrt_ctx *ctx, uint32_t spp, uint32_t max_depth);
/* Integer-RNG key material.  The shader's float hash (Tracer.comp:221-234) is replaced by
 * a counter-based PCG hash keyed by (seed64, float bits of FrameData.seed, frame index,
 * pixel, sample, dimension); see DESIGN.md "RNG". */
VKRT_API vkrt_error vkrt_set_seed(vkrt_ctx *ctx, uint64_t seed);
VKRT_API vkrt_error vkrt_set_frame_index(vkrt_ctx *ctx, uint32_t frame_index);
VKRT_API vkrt_error vkrt_reset_accum(vkrt_ctx *ctx);

/* ------------------------------------------------------------------------- */
/* Scene                                                                      */
/* ------------------------------------------------------------------------- */
/* ref: the binding-1 SSBO upload, Source/GraphicsDevice.cpp:796-822 (48 B stride). */
VKRT_API vkrt_error vkrt_set_triangles(vkrt_ctx *ctx, const vkrt_triangle *tris, uint32_t n);
/* All triangles share one material: `mirror` in Tracer.comp:386, materials[1] in Raytracer.comp:116. */
VKRT_API vkrt_error vkrt_set_triangle_material(vkrt_ctx *ctx, uint32_t mat_id);
/* Per-triangle materials (new: meshes; ref: the TODO at Raytracer.comp:10): one material id per triangle of the list set
 * last, or n = 0 to return to the single shared material. */
VKRT_API vkrt_error vkrt_set_triangle_materials(vkrt_ctx *ctx, const uint32_t *mat_ids, uint32_t n);
/* Minimal mesh loader: reads the `v` / `f` records of a Wavefront OBJ file into the reference's 48-byte Triangle layout
 * (polygons are fanned; xform = optional row-major 3x4 matrix applied to every vertex, or NULL).  Writes at most `capacity`
 * triangles and always reports the number the file holds. */
VKRT_API vkrt_error vkrt_load_obj(const char *path, const float xform[12], vkrt_triangle *out, uint32_t capacity,
                                  uint32_t *n_triangles);
VKRT_API vkrt_error vkrt_set_materials(vkrt_ctx *ctx, const vkrt_material *mats, uint32_t n);
VKRT_API vkrt_error vkrt_set_spheres(vkrt_ctx *ctx, const vkrt_sphere *spheres,
                                     const uint32_t *mat_id, uint32_t n);
VKRT_API vkrt_error vkrt_set_planes(vkrt_ctx *ctx, const vkrt_plane *planes,
                                    const uint32_t *mat_id, uint32_t n);
/* Loads the shader-constant scene of Tracer.comp:186-211 or Raytracer.comp:98-127
 * (+ the host's single triangle, Source/GraphicsDevice.cpp:798-803 / Raytracer.comp:114-117). */
VKRT_API vkrt_error vkrt_use_default_scene(vkrt_ctx *ctx, uint32_t which);
/* Builds the LBVH over the spheres on the device (Morton codes -> radix sort -> Karras
 * hierarchy -> bottom-up refit) and switches sphere queries from the literal in-order
 * loop (Tracer.comp:398-412) to the order-independent nearest-hit rule (DESIGN.md "S rule").
 * The triangles of the scene get a second tree of the same kind over their padded boxes and are queried by the
 * analogous rule T instead of the in-order loop of Tracer.comp:378-396 (DESIGN.md "Rule T"). */
VKRT_API vkrt_error vkrt_build_bvh(vkrt_ctx *ctx);
VKRT_API vkrt_error vkrt_clear_bvh(vkrt_ctx *ctx);

/* ------------------------------------------------------------------------- */
/* Outputs (device pointers are borrowed until the next vkrt_draw / destroy)   */
/* ------------------------------------------------------------------------- */
/* Resolved 8-bit image of the most recent frame, row 0 = bottom of the view like the shader's
 * imageStore (Tracer.comp:592; Fullscreen.frag:16 flips on display).  pitch in bytes. */
VKRT_API vkrt_error vkrt_get_rgba8(vkrt_ctx *ctx, void **dev_ptr, size_t *pitch);
/* Linear radiance accumulator: float4 {sum_r, sum_g, sum_b, n_samples} per pixel, row-major.
 * This is the parity tap ahead of Tracer.comp:585. */
VKRT_API vkrt_error vkrt_get_accum(vkrt_ctx *ctx, float **dev_ptr);
/* Primary nearest-hit ids (needs VKRT_FLAG_HIT_IDS): 0 = miss, else (kind << 28) | index,
 * kind 1 = triangle, 2 = sphere, 3 = plane. */
VKRT_API vkrt_error vkrt_get_hit_ids(vkrt_ctx *ctx, uint32_t **dev_ptr);
VKRT_API vkrt_error vkrt_get_stream(vkrt_ctx *ctx, void **stream);

/* Synchronous host copies (these wait for the stream). */
VKRT_API vkrt_error vkrt_read_rgba8(vkrt_ctx *ctx, void *host, size_t bytes);
VKRT_API vkrt_error vkrt_read_accum(vkrt_ctx *ctx, float *host, size_t bytes);
VKRT_API vkrt_error vkrt_read_hit_ids(vkrt_ctx *ctx, uint32_t *host, size_t bytes);
/* Same as vkrt_read_rgba8 but asynchronous on the context's stream into pinned host memory. */
VKRT_API vkrt_error vkrt_read_rgba8_async(vkrt_ctx *ctx, void *pinned_host, size_t bytes);
/* (Re)runs only the resolve (mean, Reinhard, gamma, dither, unorm8; Tracer.comp:585-592). */
VKRT_API vkrt_error vkrt_resolve(vkrt_ctx *ctx);

/* The present filter of the reference (ref: Assets/Fullscreen.frag:14-31 with the sampler of
 * Source/GraphicsDevice.cpp:770-794): temporal-variance-gated 4-tap blur of image slot 0 against image slot 1,
 * resampled to an out_w x out_h framebuffer (row 0 = top, like the swapchain image).  Synchronous host copy. */
VKRT_API vkrt_error vkrt_present(vkrt_ctx *ctx, void *host_rgba8, uint32_t out_w, uint32_t out_h);

/* ------------------------------------------------------------------------- */
/* Vulkan <-> CUDA interop: the traced images the engine presents             */
/* (ref: traced_images[FRAMES_IN_FLIGHT], Source/GraphicsDevice.cpp:664-699;  */
/*  the barriers around the dispatch, :1234-1252 and :1268-1284)              */
/* ------------------------------------------------------------------------- */
/* The engine allocates traced_images[slot] with VkExternalMemoryImageCreateInfo +
 * VkExportMemoryAllocateInfo (VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT), exports the memory with
 * vkGetMemoryFdKHR and hands the fd over here; from then on the resolve of every frame whose target is `slot`
 * (slot = frame number % frames_in_flight, like state.currentFrame, :1341) is written straight into that memory,
 * so Fullscreen.vert/.frag sample it unchanged and no copy crosses PCIe.  The descriptor stays the caller's in every
 * case (the library imports a duplicate of it): close it after the call. */
enum { VKRT_TILING_LINEAR  = 0,   /* a VkBuffer, or a VK_IMAGE_TILING_LINEAR image: rows of row_pitch bytes          */
       VKRT_TILING_OPTIMAL = 1 }; /* the reference's VK_IMAGE_TILING_OPTIMAL R8G8B8A8_UNORM image (:672-673):
                                     mapped as a CUDA mipmapped array (1 level), written through a surface object */
typedef struct vkrt_external_image {
    uint32_t struct_size;      /* = sizeof(vkrt_external_image) */
    int32_t  fd;               /* from vkGetMemoryFdKHR */
    uint64_t allocation_size;  /* VkMemoryAllocateInfo.allocationSize (:692) */
    uint64_t offset;           /* memoryOffset of vkBindImageMemory / vkBindBufferMemory (:697: 0) */
    uint32_t tiling;           /* VKRT_TILING_* */
    uint32_t row_pitch;        /* LINEAR: bytes per row (VkSubresourceLayout.rowPitch); 0 = width * 4 */
    uint32_t dedicated;        /* the allocation used VkMemoryDedicatedAllocateInfo */
    uint32_t _pad;
} vkrt_external_image;
VKRT_API vkrt_error vkrt_import_vk_image(vkrt_ctx *ctx, uint32_t slot, const vkrt_external_image *image);
/* Same write path for linear device memory the caller maps by itself (a CUDA allocation, or an external buffer it
 * imported on its own): the resolve of frames whose target is `slot` goes to dev_ptr, rows row_pitch bytes apart
 * (0 = width * 4).  The memory stays the caller's.  dev_ptr = NULL returns the slot to the library's own image. */
VKRT_API vkrt_error vkrt_bind_rgba8_target(vkrt_ctx *ctx, uint32_t slot, void *dev_ptr, size_t row_pitch);
/* Diagnostics: backs `slot` by a library-owned CUDA array written through a surface object, i.e. the write path of
 * a VKRT_TILING_OPTIMAL import without a Vulkan allocation behind it (what the GPU tests drive). */
VKRT_API vkrt_error vkrt_debug_bind_array_target(vkrt_ctx *ctx, uint32_t slot);
/* The two vkCmdPipelineBarriers become one semaphore pair per slot, exported by the engine with vkGetSemaphoreFdKHR
 * (OPAQUE_FD; binary or timeline):
 *   VKRT_SEMAPHORE_ACQUIRE  signalled by the engine in the last submit that samples image `slot` before the
 *                           library's next write to it; the n-th frame drawn into the slot (n >= 2) waits for it
 *                           (timeline: for value n - 1) before it writes -- replaces the barrier at :1234-1252;
 *   VKRT_SEMAPHORE_RELEASE  signalled by the library on its stream after the n-th frame drawn into the slot (timeline:
 *                           to value n); the engine's submit that samples the image waits for it -- :1268-1284.
 * n counts the vkrt_draw calls that targeted the slot since its FIRST semaphore was imported (import both before
 * drawing); an explicit vkrt_resolve of the same frame neither waits nor signals.
 * Without semaphores the caller orders the two APIs itself (vkrt_wait_idle / a fence).  Like the image's, the
 * descriptor stays the caller's. */
enum { VKRT_SEMAPHORE_ACQUIRE = 0, VKRT_SEMAPHORE_RELEASE = 1 };
VKRT_API vkrt_error vkrt_import_vk_semaphore(vkrt_ctx *ctx, uint32_t slot, uint32_t which, int32_t fd, uint32_t timeline);
/* Waits for the frames in flight, then destroys every imported object and binding; all slots return to the
 * library's own images. */
VKRT_API vkrt_error vkrt_release_external(vkrt_ctx *ctx);

/* Ray counters, cumulative since creation / vkrt_reset_counters.  A "ray" is one trace_ray
 * invocation of the reference algorithm (Tracer.comp:374, Raytracer.comp:224). */
typedef struct vkrt_counters {
    uint64_t closest_rays;  /* nearest-hit queries  */
    uint64_t shadow_rays;   /* any-hit queries      */
    uint64_t node_visits;   /* BVH nodes fetched (VKRT_FLAG_STATS) */
    uint64_t leaf_tests;    /* sphere tests made from BVH leaves (VKRT_FLAG_STATS) */
    uint64_t paths;         /* radiance() invocations */
    uint64_t frames;
    /* of the rays above, those that were answered without a traversal of their own: */
    uint64_t shared_primary_rays;   /* samples 2..S of a pixel reuse the pixel's one primary-ray query (Tracer.comp:574-581) */
    uint64_t zero_term_shadow_rays; /* shadow rays whose unoccluded contribution is exactly 0 (light behind the surface) */
} vkrt_counters;
VKRT_API vkrt_error vkrt_get_counters(vkrt_ctx *ctx, vkrt_counters *out);
VKRT_API vkrt_error vkrt_reset_counters(vkrt_ctx *ctx);
/* Device milliseconds (CUDA events on the context's stream) of the trace kernels of the most
 * recently completed vkrt_draw, and how many kernels that draw launched. */
VKRT_API vkrt_error vkrt_last_frame_timing(vkrt_ctx *ctx, float *trace_ms, float *total_ms,
                                           uint32_t *n_launches);

/* Device milliseconds summed over the traversal-kernel launches of the last frame (the megakernel's one
 * launch, or every extend / shadow launch of the wavefront), each bracketed by its own CUDA events, how many
 * launches that was, and the same sum over ALL kernel launches of the frame (the wavefront's two lanes overlap,
 * so these sums can exceed the frame time).  This is the duration bench.py's roofline divides by. */
VKRT_API vkrt_error vkrt_last_frame_traversal_timing(vkrt_ctx *ctx, float *traversal_ms, uint32_t *n_launches,
                                                     float *all_kernels_ms);

/* Diagnostics: per-launch start/end times of the most recent wavefront frame, as text lines
 * "lane kernel start_ms end_ms" (relative to the frame's first launch). */
VKRT_API vkrt_error vkrt_debug_dump_timeline(vkrt_ctx *ctx, const char *path);

/* ------------------------------------------------------------------------- */
/* BVH introspection (tests, DESIGN.md byte accounting)                       */
/* ------------------------------------------------------------------------- */
typedef struct vkrt_bvh_info {
    uint32_t n_spheres;
    uint32_t n_nodes;       /* internal nodes, 64 B each */
    uint32_t node_bytes;
    float    build_ms;      /* device time of the last vkrt_build_bvh */
    uint32_t build_launches;
    uint32_t depth;         /* inner-node levels on the longest root-to-leaf chain; <= 64 by construction (distinct 64-bit
                               (Morton code, index) keys), and the bound of the kernels' traversal stacks (128 entries) */
    uint32_t n_triangles;   /* the triangles' own tree (rule T): triangles, inner nodes, depth, device build time */
    uint32_t n_tri_nodes;
    uint32_t tri_depth;
    float    tri_build_ms;
    uint32_t traversal_depth;    /* the tree the wavefront's 32-byte traversal nodes are made from: the same leaves and exact
                                    union boxes as the LBVH (so rule S's answer is the same), splits chosen top-down by a binned
                                    surface-area heuristic on the device; its depth, whether it is that tree (1) or the LBVH
                                    itself (0), and its share of build_ms */
    uint32_t traversal_is_sah;
    float    traversal_build_ms;
} vkrt_bvh_info;
VKRT_API vkrt_error vkrt_get_bvh_info(vkrt_ctx *ctx, vkrt_bvh_info *out);
/* Copies the packed nodes (n_nodes * 16 floats) to the host. */
VKRT_API vkrt_error vkrt_read_bvh_nodes(vkrt_ctx *ctx, float *host, size_t bytes);
/* The same layout for the traversal tree (n_nodes * 16 floats, root = node 0, breadth-first ids). */
VKRT_API vkrt_error vkrt_read_bvh_traversal_nodes(vkrt_ctx *ctx, float *host, size_t bytes);
/* Copies the 32-byte traversal nodes (n_nodes * 8 uint32: two child records {x: lo | hi << 16, y, z, ref}, ref =
 * inner node index or ~sphere) and their 16-bit grid (scale[3], offset[3]: the coordinate of code q is
 * (2^23 + q) * scale + offset) to the host.  Introspection for the containment tests. */
VKRT_API vkrt_error vkrt_read_bvh_qnodes(vkrt_ctx *ctx, uint32_t *host, size_t bytes, float grid[6]);

/* ------------------------------------------------------------------------- */
/* Sharding (one context per GPU; the exchange itself is done by the caller,  */
/* e.g. torch.distributed/NCCL gather, on these device buffers)               */
/* ------------------------------------------------------------------------- */
/* Packs the accumulator of the tiles this context owns into a compact buffer
 * (n_owned_tiles * 1024 float4) and returns it. */
VKRT_API vkrt_error vkrt_pack_shard(vkrt_ctx *ctx, float **dev_ptr, size_t *n_floats);
/* Same, into a caller-owned device buffer (e.g. a torch tensor that NCCL will send);
 * n_floats must be at least what vkrt_shard_floats(ctx, 0, ..) reports (the largest shard). */
VKRT_API vkrt_error vkrt_pack_shard_into(vkrt_ctx *ctx, float *dev_dst, size_t n_floats);
/* Number of floats vkrt_pack_shard produces for shard `tile_rank` of this image. */
VKRT_API vkrt_error vkrt_shard_floats(vkrt_ctx *ctx, uint32_t tile_rank, size_t *n_floats);
/* On the gathering context: scatters the packed buffer of (tile_rank, sample_rank) into the
 * full accumulator; add = 0 overwrites, add = 1 sums (sample shards, applied in call order). */
VKRT_API vkrt_error vkrt_unpack_shard(vkrt_ctx *ctx, const float *dev_packed, uint32_t tile_rank,
                                      uint32_t tile_count, int add);

/* ------------------------------------------------------------------------- */
/* Frame exchange over peer memory (NVLink / NVSwitch) -- SURVEY.md 8(e)      */
/* ------------------------------------------------------------------------- */
/* With an exchange attached the sharded contexts of one frame (tile_shard_* x sample_shard_*, one per GPU, in one or in
 * several processes) need no pack / gather / unpack step: the last kernel of every context's vkrt_draw stores the pixels
 * it owns directly into a frame target in the memory of the GATHERING context's GPU (tile rank 0, sample rank 0), through
 * a peer mapping, and publishes a frame counter there; the gathering context's vkrt_draw waits for the counters of all
 * ranks on the device, adds the sample groups in rank order into its accumulator and resolves.  Everything is ordered on
 * the device (no host synchronisation, no collective call), frames stay two deep in flight.  Every context must draw
 * the same sequence of frames.  VKRT_FLAG_PROGRESSIVE is applied by the gathering context.
 *   vkrt_exchange_create  gathering context: allocates the exchange block and returns a handle (CUDA IPC) for it
 *   vkrt_exchange_open    a context in ANOTHER process: maps the block from the handle (cudaIpcOpenMemHandle)
 *   vkrt_exchange_attach  a context in the SAME process: maps it through cudaDeviceEnablePeerAccess
 *   vkrt_exchange_close   detaches (waits for the frames in flight) */
typedef struct vkrt_exchange_handle {
    uint8_t  ipc[64];            /* cudaIpcMemHandle_t */
    uint64_t bytes;
    uint32_t width, height;
    uint32_t tile_shard_count, sample_shard_count;
    uint32_t magic, _pad;
} vkrt_exchange_handle;
VKRT_API vkrt_error vkrt_exchange_create(vkrt_ctx *ctx, vkrt_exchange_handle *out);
VKRT_API vkrt_error vkrt_exchange_open(vkrt_ctx *ctx, const vkrt_exchange_handle *handle);
VKRT_API vkrt_error vkrt_exchange_attach(vkrt_ctx *ctx, vkrt_ctx *gathering_ctx);
VKRT_API vkrt_error vkrt_exchange_close(vkrt_ctx *ctx);

/* Microbenchmarks for the roofline denominators that MEASURED_PEAKS.json lacks. */
VKRT_API vkrt_error vkrt_measure_fp32_peak(int device_id, float *tflops);
VKRT_API vkrt_error vkrt_measure_l2_bandwidth(int device_id, float *gbs);

VKRT_API const char *vkrt_last_error_string(vkrt_ctx *ctx); /* ctx may be NULL: last create error */
VKRT_API const char *vkrt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VKRT_H */
