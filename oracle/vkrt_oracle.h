/*
 * CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A line-by-line CPU restatement of the reference's two compute shaders
 * (Assets/Tracer.comp, Assets/Raytracer.comp).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; nothing under
 * vk-renderer_b200/ includes, links or calls it.
 *
 * PINS.  The reference ships no tests, golden images or known-answer vectors, and no Vulkan
 * driver exists in this image -- but it does ship the compiled shaders its engine loads
 * (Assets/Compiled/{Raytracer.comp,Tracer.comp,Fullscreen.frag}.spv, Source/GraphicsDevice.cpp:1086-1091).
 * oracle/spirv_interp.py executes those binaries on the CPU; their outputs are committed as
 * tests/golden/spirv_vectors.npz (generator: tests/golden/make_spirv_vectors.py) and this
 * restatement is checked against them in tests/test_spirv_pins.py: the 8-bit images of both
 * compute shaders and the primary-hit primitive ids identical, linear radiance within 1e-3
 * relative for >= 99.9 % of the pixels (RMSE <= 1e-4), every hit / miss decision of the three
 * intersection routines identical.  (The interpreter is this repository's, and Tracer.comp's
 * rand() is answered with the integer RNG below while the binary runs: the pin is "the
 * reference's compiled code under an independent IEEE executor", not a vendor's Vulkan driver.)
 * Further pins: the reference's own Source/Camera.cpp compiled from where it lies (oracle/_ref,
 * see oracle/Makefile), the struct layouts of Include/GraphicsDevice.h / Include/Camera.h, and
 * libm for the transcendental routines.
 *
 * Stated deviation: rand() (Tracer.comp:221-234, an implementation-defined float hash)
 * is replaced by a counter-based integer hash (see oracle_rand_u01).
 */
#ifndef VKRT_ORACLE_H
#define VKRT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mirrors of the boundary PODs (own definitions on purpose; static_asserted in the .cpp) */
typedef struct { float x, y, z, _pad; } orc_vec3a;
typedef struct { orc_vec3a pos, dir, right, up; } orc_camera_data;           /* Camera.h:5-12       */
typedef struct { float aspect_ratio, seed, _pad0[2]; orc_vec3a light_pos;
                 orc_camera_data camera; } orc_frame_data;                   /* GraphicsDevice.h:20 */
typedef struct { orc_vec3a v0, v1, v2; } orc_triangle;                       /* GraphicsDevice.h:13 */
typedef struct { float albedo[3], roughness, emissive[3], metalness;
                 uint32_t type, _pad[3]; } orc_material;                     /* Tracer.comp:59-74   */
typedef struct { float cx, cy, cz, r; } orc_sphere;
typedef struct { float nx, ny, nz, len; } orc_plane;

enum { ORC_WHITTED = 0, ORC_PATH = 1 };
/* how the sphere list is queried */
enum { ORC_SPHERES_LITERAL  = 0,  /* in-order loop with the +EPSILON chain rule (Tracer.comp:398-412) */
       ORC_SPHERES_S_LINEAR = 1,  /* order-independent rule "S", evaluated by a linear scan          */
       ORC_SPHERES_S_BVH    = 2 };/* rule "S", evaluated by traversing the oracle's own CPU LBVH     */
/* In the two S modes (scenes with a hierarchy) triangles follow the analogous order-independent rule "T", evaluated by
 * a linear scan over all triangles; in LITERAL mode they keep the reference's in-order loop (Tracer.comp:378-396). */

typedef struct orc_scene orc_scene;

typedef struct {
    uint32_t width, height;
    uint32_t spp, max_depth;
    uint32_t integrator;
    uint32_t sphere_mode;
    uint64_t seed;
    uint32_t frame_index;
    uint32_t sample_begin, sample_end;   /* samples [begin,end) of the frame; 0,0 = [0,spp) */
    uint32_t x0, y0, x1, y1;             /* pixel rectangle [x0,x1) x [y0,y1); 0,0,0,0 = whole image.
                                            Output buffers are always full-size (width*height). */
    uint32_t n_threads;                  /* 0 = hardware_concurrency */
    uint32_t accumulate;                 /* 1 = add this frame's sums to accum (progressive) */
} orc_params;

typedef struct {
    uint64_t closest_rays, shadow_rays, node_visits, leaf_tests, paths;
    uint64_t literal_vs_s_mismatch;      /* primary rays whose literal-rule sphere result differs from rule S */
} orc_counters;

orc_scene *orc_scene_create(void);
void       orc_scene_destroy(orc_scene *);
int  orc_scene_set_materials(orc_scene *, const orc_material *, uint32_t n);
int  orc_scene_set_spheres(orc_scene *, const orc_sphere *, const uint32_t *mat_id, uint32_t n);
int  orc_scene_set_planes(orc_scene *, const orc_plane *, const uint32_t *mat_id, uint32_t n);
int  orc_scene_set_triangles(orc_scene *, const orc_triangle *, uint32_t n, uint32_t mat_id);
/* per-triangle materials (n = number of triangles, or 0 to go back to the single mat_id of set_triangles) */
int  orc_scene_set_triangle_materials(orc_scene *, const uint32_t *mat_ids, uint32_t n);
int  orc_scene_use_default(orc_scene *, uint32_t which /* 0 Tracer.comp, 1 Raytracer.comp */);
int  orc_scene_build_bvh(orc_scene *);          /* CPU LBVH (Morton / sort / Karras / refit) */
/* test hook: replace the hierarchy by any tree in orc_scene_read_bvh's layout (rule S must not depend on it) */
int  orc_scene_set_bvh(orc_scene *, const float *nodes, uint32_t n_nodes);
uint32_t orc_scene_bvh_nodes(const orc_scene *);
/* node i: 16 floats, same logical content as the device layout (child record 0, child record 1) */
int  orc_scene_read_bvh(const orc_scene *, float *out, size_t bytes);

/* Renders a frame.  accum: width*height*4 floats {sum rgb, n}; hit_ids / rgba8 may be NULL. */
int  orc_render(const orc_scene *, const orc_params *, const orc_frame_data *,
                float *accum, uint32_t *hit_ids, uint8_t *rgba8, orc_counters *counters);
/* Resolve only (Tracer.comp:585-592 / Raytracer.comp:398). */
int  orc_resolve(const orc_params *, const orc_frame_data *, const float *accum, uint8_t *rgba8);

/* Fullscreen.frag: temporal-variance-gated 4-tap blur of binding 0 vs binding 1 (both tw x th rgba8, row 0 =
 * bottom) into a W x H rgba8 framebuffer (row 0 = top). */
int  orc_present(const uint8_t *binding0, const uint8_t *binding1, uint32_t tw, uint32_t th, uint8_t *out, uint32_t W, uint32_t H);

/* unit-level entry points for known-answer tests */
float    orc_sin(float), orc_cos(float), orc_exp2(float), orc_log2(float), orc_pow(float, float);
uint32_t orc_pcg_hash(uint32_t);
uint32_t orc_frame_key(uint64_t seed, float frame_seed, uint32_t frame_index);
float    orc_rand_u01(uint32_t frame_key, uint32_t pixel, uint32_t sample, uint32_t dim);
float    orc_sphere_intersect(const float o[3], const float d[3], const orc_sphere *);
float    orc_plane_intersect_tracer(const float o[3], const float d[3], const orc_plane *);
float    orc_plane_intersect_raytracer(const float o[3], const float d[3], const orc_plane *);
float    orc_tri_intersect(const float o[3], const float d[3], const orc_triangle *, float epsilon);
void     orc_primary_ray(const orc_frame_data *, uint32_t w, uint32_t h, uint32_t x, uint32_t y,
                         float o[3], float d[3]);
/* slab test of rule S: returns 1 on hit and writes tn/tf */
int      orc_slab(const float o[3], const float d[3], const float lo[3], const float hi[3],
                  float *tn, float *tf);
/* one sphere query on an arbitrary ray, any mode; returns sphere index or -1, writes t */
int      orc_query_spheres(const orc_scene *, uint32_t mode, const float o[3], const float d[3],
                           float bound, float epsilon, float *t_out);

#ifdef __cplusplus
}
#endif
#endif
