// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See vkrt_oracle.h for the contract
// (what pins it: the reference's compiled SPIR-V run by oracle/spirv_interp.py; the one stated deviation: the RNG).
//
// Every function cites the reference lines it restates (paths relative to the reference
// root).  Arithmetic rule ("vkrt-f32"): every operation is one IEEE-754 binary32
// round-to-nearest operation (+ - * / sqrt), fused multiply-add ONLY where fmaf is written,
// transcendental functions are the polynomial routines below.  Build with
// -ffp-contract=off and no fast-math so the compiler adds or removes nothing.
#include "vkrt_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

static_assert(sizeof(orc_camera_data) == 64, "CameraData is 64 B (Include/Camera.h:5-12)");
static_assert(sizeof(orc_frame_data) == 96, "FrameData is 96 B (Include/GraphicsDevice.h:20-29)");
static_assert(offsetof(orc_frame_data, seed) == 4 && offsetof(orc_frame_data, light_pos) == 16 &&
              offsetof(orc_frame_data, camera) == 32, "FrameData offsets (Tracer.comp.spv: 0/4/16/32)");
static_assert(sizeof(orc_triangle) == 48, "Triangle is 48 B (Include/GraphicsDevice.h:13-18)");
static_assert(sizeof(orc_material) == 48, "material record");

namespace {

// ------------------------------------------------------------------------------------------
// vkrt-f32 arithmetic primitives
// ------------------------------------------------------------------------------------------
inline float fma_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// GLSL 4.50 8.3: max(x,y) = y if x < y else x;  min(x,y) = y if y < x else x.  The result for a
// NaN operand is undefined in GLSL; here (as on FMNMX hardware) a NaN x yields y, so the firefly
// clamp of Tracer.comp:441 scrubs a NaN accumulator to 0.
inline float gl_max(float x, float y) { return (x != x) ? y : (x < y ? y : x); }
inline float gl_min(float x, float y) { return (x != x) ? y : (y < x ? y : x); }
inline float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
inline float gl_sign(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
inline float gl_abs(float x) { return u2f(f2u(x) & 0x7fffffffu); }
// GLSL mix(x,y,a) = x*(1-a) + y*a ; the sum is fused
inline float gl_mix(float x, float y, float a) { return fma_(y, a, x * (1.0f - a)); }

const float PI = 3.14159265359f;        // Tracer.comp:176
const float TWO_PI = 2.0f * PI;         // "2.0 * PI" constant-folded (Tracer.comp:454,469)

// round-to-nearest-even to an integer-valued float, |x| < 2^22
inline float rne(float x) { return (x + 12582912.0f) - 12582912.0f; }

// sin/cos: Cody-Waite reduction by pi/2, degree-9/8 minimax polynomials on [-pi/4, pi/4]
inline void sincos_(float x, float &s, float &c)
{
    const float q = rne(x * 6.366197467e-01f);
    const int n = (int)q;
    float r = fma_(q, -1.570796371e+00f, x);        // pi/2 = hi + lo, both binary32
    r = fma_(q, 4.371138829e-08f, r);
    const float r2 = r * r;
    float sp = 2.723468469e-06f;
    sp = fma_(sp, r2, -1.983996626e-04f);
    sp = fma_(sp, r2, 8.333331905e-03f);
    sp = fma_(sp, r2, -1.666666716e-01f);
    const float sr = fma_(r * r2, sp, r);
    float cp = -2.728823461e-07f;
    cp = fma_(cp, r2, 2.480049989e-05f);
    cp = fma_(cp, r2, -1.388888806e-03f);
    cp = fma_(cp, r2, 4.166666791e-02f);
    const float cr = fma_(r2 * r2, cp, fma_(-0.5f, r2, 1.0f));
    const float s0 = (n & 1) ? cr : sr;
    const float c0 = (n & 1) ? sr : cr;
    s = (n & 2) ? -s0 : s0;
    c = ((n + 1) & 2) ? -c0 : c0;
}

// log2: x = 2^e * m, m in [2/3, 4/3); log2(m) = f * P(f), f = m - 1
inline float log2_(float x)
{
    uint32_t b = f2u(x);
    if (b == 0u || b == 0x80000000u) return -INFINITY;     // log2(+-0)
    if (b >> 31) return NAN;                                // negative
    if (b >= 0x7f800000u) return x;                         // inf, nan
    float bias = 0.0f;
    if (b < 0x00800000u) { x = x * 8388608.0f; b = f2u(x); bias = -23.0f; }   // subnormal
    const int e = (int)(b - 0x3f2aaaabu) >> 23;
    const float m = u2f(b - ((uint32_t)e << 23));
    const float f = m - 1.0f;
    float p = 2.032371908e-01f;
    p = fma_(p, f, -2.222205549e-01f);
    p = fma_(p, f, 2.016990036e-01f);
    p = fma_(p, f, -2.367866337e-01f);
    p = fma_(p, f, 2.887182832e-01f);
    p = fma_(p, f, -3.607895672e-01f);
    p = fma_(p, f, 4.808958173e-01f);
    p = fma_(p, f, -7.213465571e-01f);
    p = fma_(p, f, 1.442695022e+00f);
    return fma_(p, f, (float)e + bias);
}

// exp2: x = n + f, f in [-0.5, 0.5]; 2^f by a degree-6 polynomial, 2^n applied in two halves
inline float exp2_(float x)
{
    if (x != x) return x;
    if (x >= 128.0f) return INFINITY;
    if (x < -150.0f) return 0.0f;
    const float nf = rne(x);
    const float f = x - nf;
    const int n = (int)nf;
    float p = 1.546973217e-04f;
    p = fma_(p, f, 1.340043265e-03f);
    p = fma_(p, f, 9.618025273e-03f);
    p = fma_(p, f, 5.550327152e-02f);
    p = fma_(p, f, 2.402265072e-01f);
    p = fma_(p, f, 6.931471825e-01f);
    p = fma_(p, f, 1.0f);
    const int n1 = n >> 1, n2 = n - n1;
    return (p * u2f((uint32_t)(n1 + 127) << 23)) * u2f((uint32_t)(n2 + 127) << 23);
}

// GLSL pow(x,y) = exp2(y * log2(x)); undefined for x < 0 (NaN here), pow(0, y>0) = 0
inline float pow_(float x, float y) { return exp2_(y * log2_(x)); }

struct V3 { float x, y, z; };
inline V3 v3(float a) { return {a, a, a}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
// vec3 / scalar: one correctly rounded reciprocal, three multiplies (how GPUs lower it; GLSL allows 2.5 ULP on '/')
inline V3 operator/(V3 a, float s) { const float inv = 1.0f / s; return {a.x * inv, a.y * inv, a.z * inv}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot3(V3 a, V3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
inline V3 cross3(V3 a, V3 b)
{
    return {fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x))};
}
inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
inline V3 normalize3(V3 a) { const float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
inline V3 madd3(float t, V3 d, V3 o) { return {fma_(t, d.x, o.x), fma_(t, d.y, o.y), fma_(t, d.z, o.z)}; }
// GLSL reflect(I,N) = I - 2*dot(N,I)*N
inline V3 reflect3(V3 I, V3 N) { const float k = 2.0f * dot3(N, I); return madd3(-k, N, I); }
// GLSL refract(I,N,eta): k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 ? 0 : eta*I - (eta*dot(N,I) + sqrt(k))*N
inline V3 refract3(V3 I, V3 N, float eta)
{
    const float d = dot3(N, I);
    const float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return v3(0.0f);
    const float s = fma_(eta, d, sqrtf(k));
    return madd3(-s, N, eta * I);
}
inline V3 clamp3(V3 a, float lo, float hi) { return {gl_clamp(a.x, lo, hi), gl_clamp(a.y, lo, hi), gl_clamp(a.z, lo, hi)}; }
inline V3 mix3(V3 a, V3 b, float t) { return {gl_mix(a.x, b.x, t), gl_mix(a.y, b.y, t), gl_mix(a.z, b.z, t)}; }
inline V3 from3a(const orc_vec3a &v) { return {v.x, v.y, v.z}; }
inline V3 fromp(const float *p) { return {p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------
// RNG -- replaces Tracer.comp:221-234.  Counter-based: keyed by (frame key, pixel, sample,
// dimension), so the draw order and the launch geometry cannot change a path.
// ------------------------------------------------------------------------------------------
inline uint32_t pcg(uint32_t v)
{
    const uint32_t state = v * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
inline uint32_t frame_key_(uint64_t seed, float frame_seed, uint32_t frame_index)
{
    const uint32_t a = pcg(frame_index + 0x9E3779B9u);
    const uint32_t b = pcg(f2u(frame_seed) ^ a);
    const uint32_t c = pcg((uint32_t)(seed >> 32) ^ b);
    return pcg((uint32_t)seed ^ c);
}
inline uint32_t sample_key_(uint32_t frame_key, uint32_t pixel, uint32_t sample)
{
    const uint32_t kp = pcg(pixel + frame_key);
    return pcg(kp ^ (sample * 0x9E3779B9u));
}
inline float u01_(uint32_t sample_key, uint32_t dim)
{
    return (float)(pcg(sample_key + dim) >> 8) * 5.9604644775390625e-08f;   // [0,1)
}
// dimension layout: 32 per bounce; slot 0 = r2 (DIFFUSE, Tracer.comp:453) or the reflect/refract pick
// (DIELECTRIC, :541); 1 = phi (:454); 2 = Russian roulette (:547); 3+2l / 4+2l = light l (:468,:469)
enum { DIMS_PER_BOUNCE = 32, SLOT_R2 = 0, SLOT_PHI = 1, SLOT_RR = 2, SLOT_LIGHT = 3, MAX_LIGHTS = 14 };
const uint32_t DITHER_SAMPLE = 0xFFFFFFFFu;     // the final rand() of main (:590)

// ------------------------------------------------------------------------------------------
// Scene
// ------------------------------------------------------------------------------------------
struct Material { V3 albedo, emissive; float roughness, metalness; uint32_t type; };

struct BvhChild { float a[4]; float b[2]; int32_t index; int32_t kind; };   // 32 B
struct BvhNode { BvhChild c[2]; };                                           // 64 B
static_assert(sizeof(BvhNode) == 64, "node");

} // namespace

struct orc_scene {
    std::vector<Material> mats;
    std::vector<orc_sphere> spheres;
    std::vector<uint32_t> sphere_mat;
    std::vector<orc_plane> planes;
    std::vector<uint32_t> plane_mat;
    std::vector<orc_triangle> tris;
    uint32_t tri_mat = 0;
    std::vector<uint32_t> tri_mats;      // per-triangle material ids (empty: every triangle uses tri_mat, Tracer.comp:386)
    std::vector<uint32_t> lights;        // indices of spheres with emissive != 0 (Tracer.comp:462), ascending
    std::vector<BvhNode> bvh;
    bool has_bvh = false;
};

namespace {

const uint32_t KIND_TRI = 1, KIND_SPHERE = 2, KIND_PLANE = 3;

struct Ray { V3 o, d; };
struct Hit { float t; uint32_t kind, index; };

struct Stats { uint64_t closest = 0, shadow = 0, nodes = 0, leaves = 0, paths = 0, mism = 0; };

// ---- Tracer.comp:314-329 / Raytracer.comp:163-178 -----------------------------------------
inline float sphere_intersect(const Ray &ray, const orc_sphere &s)
{
    const V3 oc = ray.o - V3{s.cx, s.cy, s.cz};
    const float b = 2.0f * dot3(oc, ray.d);
    const float c = fma_(-s.r, s.r, dot3(oc, oc));
    const float h = fma_(b, b, -(4.0f * c));
    if (h < 0.0f) return -1.0f;
    return (-b - sqrtf(h)) * 0.5f;
}
// ---- Tracer.comp:331-338 -------------------------------------------------------------------
inline float plane_intersect_tracer(const Ray &ray, const orc_plane &p)
{
    const V3 N{p.nx, p.ny, p.nz};
    const float d = dot3(ray.d, N);
    const float dist = -(p.len + dot3(ray.o, N)) / d;
    const float when_neq = gl_abs(gl_sign(d - 0.0f));                 // :246-249
    return when_neq * gl_max(dist, 0.0f);
}
// ---- Raytracer.comp:180-192 ----------------------------------------------------------------
inline float plane_intersect_raytracer(const Ray &ray, const orc_plane &p)
{
    const V3 N{p.nx, p.ny, p.nz};
    const float d = dot3(ray.d, N);
    if (d == 0.0f) return 0.0f;
    const float dist = -(p.len + dot3(ray.o, N)) / d;
    return gl_max(dist, 0.0f);
}
// ---- Tracer.comp:340-372 / Raytracer.comp:129-161 (EPSILON differs: 1e-3 / 0.01) ------------
inline float tri_intersect(const Ray &ray, const orc_triangle &tri, float eps)
{
    const V3 v0 = from3a(tri.v0);
    const V3 v0v1 = from3a(tri.v1) - v0;
    const V3 v0v2 = from3a(tri.v2) - v0;
    const V3 pvec = cross3(ray.d, v0v2);
    const float det = dot3(v0v1, pvec);
    if (det < eps) return -1.0f;                                      // back-face cull
    const float inv_det = 1.0f / det;
    const V3 tvec = ray.o - v0;
    const float u = dot3(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return -1.0f;
    const V3 qvec = cross3(tvec, v0v1);
    const float v = dot3(ray.d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return -1.0f;
    return dot3(v0v2, qvec) * inv_det;
}

// ------------------------------------------------------------------------------------------
// Rule S (new; the reference has no BVH).  A sphere i is a CANDIDATE for
// ray (o,d), lower bound eps (exclusive) and upper bound B (exclusive) iff
//   box(i) = [c - rp, c + rp], rp = r*1.001 + 0.001, is hit by the slab test below with
//   tn <= tf, tf >= 0, and t_i = sphere_intersect satisfies eps < t_i < B and tn <= t_i.
// The nearest hit is the candidate with the smallest (t_i, i) lexicographically.  Because the
// slab test is monotone under box inclusion in floating point, ANY bounding hierarchy whose
// inner boxes are exact min/max unions of leaf boxes, traversed in ANY order with the cull
// rule "skip a box whose tn > current best", returns exactly this result.
// ------------------------------------------------------------------------------------------
struct SlabRay { V3 inv, oinv; };
inline float mn(float a, float b) { return a < b ? a : b; }
inline float mx(float a, float b) { return a > b ? a : b; }
inline float safe_inv(float d)
{
    const float ad = gl_abs(d);
    const float dd = ad > 1e-20f ? d : copysignf(1e-20f, d);
    return 1.0f / dd;
}
inline SlabRay slab_setup(const Ray &r)
{
    SlabRay s;
    s.inv = {safe_inv(r.d.x), safe_inv(r.d.y), safe_inv(r.d.z)};
    s.oinv = r.o * s.inv;
    return s;
}
inline bool slab_test(const SlabRay &s, V3 lo, V3 hi, float &tn, float &tf)
{
    const float t0x = fma_(lo.x, s.inv.x, -s.oinv.x), t1x = fma_(hi.x, s.inv.x, -s.oinv.x);
    const float t0y = fma_(lo.y, s.inv.y, -s.oinv.y), t1y = fma_(hi.y, s.inv.y, -s.oinv.y);
    const float t0z = fma_(lo.z, s.inv.z, -s.oinv.z), t1z = fma_(hi.z, s.inv.z, -s.oinv.z);
    // all six values are finite by construction (safe_inv), so plain compares == fminf/fmaxf
    tn = mx(mx(mn(t0x, t1x), mn(t0y, t1y)), mn(t0z, t1z));
    tf = mn(mn(mx(t0x, t1x), mx(t0y, t1y)), mx(t0z, t1z));
    return tn <= tf && tf >= 0.0f;
}
inline float sphere_pad_radius(float r) { return r * 1.001f + 0.001f; }
inline void sphere_box(const orc_sphere &s, V3 &lo, V3 &hi)
{
    const float rp = sphere_pad_radius(s.r);
    lo = {s.cx - rp, s.cy - rp, s.cz - rp};
    hi = {s.cx + rp, s.cy + rp, s.cz + rp};
}

struct SBest { float t; int idx; };
// tests sphere `i` under rule S against the running best; tn = slab entry of its own box
inline void s_consider(const Ray &ray, const orc_sphere &sp, int i, float tn, float eps, float B, SBest &best)
{
    const float t = sphere_intersect(ray, sp);
    if (!(t > eps) || !(tn <= t)) return;
    if (best.idx < 0) { if (t < B) { best.t = t; best.idx = i; } }
    else if (t < best.t || (t == best.t && i < best.idx)) { best.t = t; best.idx = i; }
}

inline SBest s_query_linear(const orc_scene &sc, const Ray &ray, float eps, float B, bool any, Stats &st)
{
    SBest best{B, -1};
    const SlabRay sr = slab_setup(ray);
    for (size_t i = 0; i < sc.spheres.size(); ++i) {
        V3 lo, hi; sphere_box(sc.spheres[i], lo, hi);
        float tn, tf;
        if (!slab_test(sr, lo, hi, tn, tf)) continue;
        if (!(tn <= best.t)) continue;
        ++st.leaves;
        s_consider(ray, sc.spheres[i], (int)i, tn, eps, B, best);
        if (any && best.idx >= 0) break;
    }
    return best;
}

inline SBest s_query_bvh(const orc_scene &sc, const Ray &ray, float eps, float B, bool any, Stats &st)
{
    SBest best{B, -1};
    if (sc.bvh.empty()) return best;
    const SlabRay sr = slab_setup(ray);
    int stack[96]; int sp = 0; int node = 0;
    for (;;) {
        const BvhNode &n = sc.bvh[node];
        ++st.nodes;
        int next[2]; float tnx[2]; int nn = 0;
        for (int k = 0; k < 2; ++k) {
            const BvhChild &c = n.c[k];
            float tn, tf;
            const V3 lo{c.a[0], c.a[1], c.a[2]}, hi{c.a[3], c.b[0], c.b[1]};
            if (!slab_test(sr, lo, hi, tn, tf) || !(tn <= best.t)) continue;
            if (c.kind == 1) {                                  // leaf: the record holds the sphere's own padded box
                ++st.leaves;
                s_consider(ray, sc.spheres[c.index], c.index, tn, eps, B, best);
                if (any && best.idx >= 0) return best;
            } else {
                next[nn] = c.index; tnx[nn] = tn; ++nn;
            }
        }
        if (nn == 2) {
            const int nearer = tnx[1] < tnx[0] ? 1 : 0;
            stack[sp++] = next[1 - nearer];
            node = next[nearer];
        } else if (nn == 1) node = next[0];
        else { if (sp == 0) break; node = stack[--sp]; }
    }
    return best;
}

// Rule T (new, like rule S): triangles of a scene that has a hierarchy.  Triangle j is a CANDIDATE iff its padded box
//   [min(v) - p, max(v) + p],  p = 0.001 + 0.001 * (largest extent of the unpadded box),
// passes the slab test and t_j = tri_intersect (Tracer.comp:340-372, back-face culled) satisfies eps < t_j < B and
// tn <= t_j; the nearest hit is the candidate with the smallest (t_j, j).  Order-independent, so any hierarchy over the
// padded boxes traversed in any order gives the same answer as this linear scan.  (The reference's in-order loop with
// its +EPSILON chain rule, :378-396, picks the LATER triangle among hits within 1e-3 -- the same kind of tie band as
// for spheres; scenes without a hierarchy keep that literal loop.)
inline void tri_box(const orc_triangle &t, V3 &lo, V3 &hi)
{
    const V3 a = from3a(t.v0), b = from3a(t.v1), c = from3a(t.v2);
    lo = {mn(mn(a.x, b.x), c.x), mn(mn(a.y, b.y), c.y), mn(mn(a.z, b.z), c.z)};
    hi = {mx(mx(a.x, b.x), c.x), mx(mx(a.y, b.y), c.y), mx(mx(a.z, b.z), c.z)};
    const float p = 0.001f + 0.001f * mx(mx(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
    lo = {lo.x - p, lo.y - p, lo.z - p};
    hi = {hi.x + p, hi.y + p, hi.z + p};
}
inline SBest t_query_linear(const orc_scene &sc, const Ray &ray, float eps, float B, bool any, Stats &st)
{
    SBest best{B, -1};
    const SlabRay sr = slab_setup(ray);
    for (size_t j = 0; j < sc.tris.size(); ++j) {
        V3 lo, hi; tri_box(sc.tris[j], lo, hi);
        float tn, tf;
        if (!slab_test(sr, lo, hi, tn, tf)) continue;
        if (!(tn <= best.t)) continue;
        ++st.leaves;
        const float t = tri_intersect(ray, sc.tris[j], eps);
        if (!(t > eps) || !(tn <= t)) continue;
        if (best.idx < 0) { if (t < B) { best.t = t; best.idx = (int)j; } }
        else if (t < best.t || (t == best.t && (int)j < best.idx)) { best.t = t; best.idx = (int)j; }
        if (any && best.idx >= 0) break;
    }
    return best;
}

// literal in-order loop with the chain rule (Tracer.comp:398-412): returns the LAST accepted sphere
inline int literal_spheres_tracer(const orc_scene &sc, const Ray &ray, float eps, float &cur)
{
    int idx = -1;
    for (size_t i = 0; i < sc.spheres.size(); ++i) {
        const float t = sphere_intersect(ray, sc.spheres[i]);
        if ((t > eps) && (t < cur + eps)) { cur = t; idx = (int)i; }
    }
    return idx;
}

// ------------------------------------------------------------------------------------------
// trace_ray -- Tracer.comp:374-431 (eps 1e-3, +eps / -eps acceptance) and
//              Raytracer.comp:224-278 (eps 0.01, strict first-wins), selected by `tracer_rules`
// ------------------------------------------------------------------------------------------
template <bool TRACER_RULES>
inline bool trace_ray(const orc_scene &sc, uint32_t mode, const Ray &ray, Hit &hit, bool shadow, Stats &st)
{
    const float EPS = TRACER_RULES ? 1e-3f : 0.01f;
    if (shadow) ++st.shadow; else ++st.closest;
    bool found = false;
    float cur = hit.t;
    if (mode != ORC_SPHERES_LITERAL && !sc.tris.empty()) {
        // scenes with a hierarchy: rule T (the same exclusive bound as the literal loop's first acceptance)
        const SBest b = t_query_linear(sc, ray, EPS, TRACER_RULES ? cur + EPS : cur, shadow, st);
        if (b.idx >= 0) { cur = b.t; hit.kind = KIND_TRI; hit.index = (uint32_t)b.idx; found = true; }
        if (shadow && found) { hit.t = cur; return true; }
    } else
    for (size_t i = 0; i < sc.tris.size(); ++i) {
        const float t = tri_intersect(ray, sc.tris[i], EPS);
        const bool acc = TRACER_RULES ? ((t > EPS) && (t < cur + EPS)) : (t > EPS && t < cur);
        if (acc) { cur = t; hit.kind = KIND_TRI; hit.index = (uint32_t)i; found = true; }
    }
    if (mode == ORC_SPHERES_LITERAL) {
        if (TRACER_RULES) {
            const int i = literal_spheres_tracer(sc, ray, EPS, cur);
            if (i >= 0) { hit.kind = KIND_SPHERE; hit.index = (uint32_t)i; found = true; }
        } else {
            for (size_t i = 0; i < sc.spheres.size(); ++i) {
                const float t = sphere_intersect(ray, sc.spheres[i]);
                if (t > EPS && t < cur) { cur = t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)i; found = true; }
            }
        }
    } else {
        const float B = TRACER_RULES ? cur + EPS : cur;
        const SBest b = (mode == ORC_SPHERES_S_BVH) ? s_query_bvh(sc, ray, EPS, B, shadow, st)
                                                    : s_query_linear(sc, ray, EPS, B, shadow, st);
        if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
    }
    for (size_t i = 0; i < sc.planes.size(); ++i) {
        const float t = TRACER_RULES ? plane_intersect_tracer(ray, sc.planes[i])
                                     : plane_intersect_raytracer(ray, sc.planes[i]);
        const bool acc = TRACER_RULES ? ((t > EPS) && (t < cur - EPS)) : (t > EPS && t < cur);
        if (acc) { cur = t; hit.kind = KIND_PLANE; hit.index = (uint32_t)i; found = true; }
    }
    hit.t = cur;
    return found;
}

struct Surface { V3 P, N; const Material *mat; };

// what Tracer.comp:384-392,404-408,420-424 / Raytracer.comp:234-242,253-257,268-272 store on accept
inline Surface surface_of(const orc_scene &sc, const Ray &ray, const Hit &hit)
{
    Surface s;
    s.P = madd3(hit.t, ray.d, ray.o);
    if (hit.kind == KIND_TRI) {
        const orc_triangle &t = sc.tris[hit.index];
        const V3 u = from3a(t.v1) - from3a(t.v0), v = from3a(t.v2) - from3a(t.v0);
        s.N = cross3(u, v);                                   // unnormalised, as in the shader
        s.mat = &sc.mats[sc.tri_mats.empty() ? sc.tri_mat : sc.tri_mats[hit.index]];
    } else if (hit.kind == KIND_SPHERE) {
        const orc_sphere &sp = sc.spheres[hit.index];
        s.N = (s.P - V3{sp.cx, sp.cy, sp.cz}) / sp.r;
        s.mat = &sc.mats[sc.sphere_mat[hit.index]];
    } else {
        const orc_plane &p = sc.planes[hit.index];
        s.N = {p.nx, p.ny, p.nz};
        s.mat = &sc.mats[sc.plane_mat[hit.index]];
    }
    return s;
}

// ---- Tracer.comp:256-263 -------------------------------------------------------------------
inline V3 jitter(V3 d, float phi, float sina, float cosa)
{
    const V3 w = normalize3(d);
    const V3 u = normalize3(cross3(V3{w.y, w.z, w.x}, w));
    const V3 v = cross3(w, u);
    float s, c; sincos_(phi, s, c);
    return (u * c + v * s) * sina + w * cosa;
}
// ---- Tracer.comp:265-272 -------------------------------------------------------------------
inline float schlick(float cosine, float ior)
{
    float r0 = (1.0f - ior) / (1.0f + ior);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * pow_(1.0f - cosine, 5.0f);
}
// ---- Tracer.comp:274-277 -------------------------------------------------------------------
inline V3 fresnel_schlick(float cosTheta, V3 F0)
{
    const float p = pow_(1.0f - cosTheta, 5.0f);
    return F0 + (v3(1.0f) - F0) * p;
}
// ---- Tracer.comp:279-292 -------------------------------------------------------------------
inline float distribution_ggx(V3 N, V3 H, float roughness)
{
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float NdotH = gl_max(dot3(N, H), 0.0f);
    const float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = PI * denom * denom;
    return a2 / denom;
}
// ---- Tracer.comp:294-312 -------------------------------------------------------------------
inline float geometry_schlick_ggx(float NdotV, float roughness)
{
    const float r = (roughness + 1.0f);
    const float k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
inline float geometry_smith(V3 N, V3 V, V3 L, float roughness)
{
    const float NdotV = gl_max(dot3(N, V), 0.0f);
    const float NdotL = gl_max(dot3(N, L), 0.0f);
    const float ggx2 = geometry_schlick_ggx(NdotV, roughness);
    const float ggx1 = geometry_schlick_ggx(NdotL, roughness);
    return ggx1 * ggx2;
}
inline float max3(V3 e) { return gl_max(gl_max(e.x, e.y), e.z); }          // :236-239

// ---- radiance(), Tracer.comp:433-553 -------------------------------------------------------
V3 radiance(const orc_scene &sc, uint32_t mode, Ray ray, V3 cam_pos, uint32_t max_depth, uint32_t skey,
            Stats &st, uint32_t *primary_id)
{
    V3 acc = v3(0.0f), mask = v3(1.0f);
    ++st.paths;
    for (uint32_t depth = 0; depth < max_depth; ++depth) {
        const uint32_t dim0 = depth * DIMS_PER_BOUNCE;
        acc = clamp3(acc, 0.0f, 1.0f);                                              // :441
        Hit hit{3000.0f / pow_((float)(depth + 1u), 2.0f), 0, 0};                   // :444
        const bool found = trace_ray<true>(sc, mode, ray, hit, false, st);
        if (depth == 0 && primary_id) *primary_id = found ? ((hit.kind << 28) | hit.index) : 0u;
        if (!found) break;                                                          // :445
        const Surface sf = surface_of(sc, ray, hit);
        const Material &mat = *sf.mat;
        if (mat.type == 0u) {                                                       // MAT_TYPE_DIFFUSE :451
            const float r2 = u01_(skey, dim0 + SLOT_R2);
            const V3 d = jitter(sf.N, TWO_PI * u01_(skey, dim0 + SLOT_PHI), sqrtf(r2), sqrtf(1.0f - r2)) *
                         (1.0f - mat.metalness);                                    // :454
            V3 e = v3(0.0f);
            for (size_t l = 0; l < sc.lights.size(); ++l) {                         // :458-462
                const orc_sphere &s = sc.spheres[sc.lights[l]];
                const Material &smat = sc.mats[sc.sphere_mat[sc.lights[l]]];
                const V3 sP{s.cx, s.cy, s.cz};
                const float t = length3(sP - sf.P) - s.r;                           // :464
                const V3 l0 = sP - sf.P;
                const float cos_a_max = sqrtf(1.0f - gl_clamp(s.r * s.r / dot3(l0, l0), 0.0f, 1.0f));
                const float cosa = gl_mix(cos_a_max, 1.0f, u01_(skey, dim0 + SLOT_LIGHT + 2 * (uint32_t)l));
                const V3 L = jitter(l0, TWO_PI * u01_(skey, dim0 + SLOT_LIGHT + 2 * (uint32_t)l + 1),
                                    sqrtf(1.0f - cosa * cosa), cosa);               // :469
                Hit sh{t, 0, 0};
                if (!trace_ray<true>(sc, mode, Ray{sf.P, L}, sh, true, st)) {       // :473
                    V3 attenuation = smat.emissive * 1.0f / pow_(t / s.r + 1.0f, 2.0f);
                    attenuation = (attenuation - v3(0.001f)) / (1.0f - 0.001f);     // :478
                    attenuation = {gl_max(attenuation.x, 0.0f), gl_max(attenuation.y, 0.0f), gl_max(attenuation.z, 0.0f)};
                    V3 F0 = v3(0.04f);
                    F0 = mix3(F0, mat.albedo, mat.metalness);
                    const V3 V = normalize3(cam_pos - sf.P);                        // :484
                    const V3 H = normalize3(V + L);
                    const float NDF = distribution_ggx(sf.N, H, mat.roughness);
                    const float G = geometry_smith(sf.N, V, L, mat.roughness);
                    const V3 F = fresnel_schlick(gl_max(dot3(H, V), 0.0f), F0);
                    const V3 kS = F;
                    V3 kD = v3(1.0f) - kS;
                    kD = kD * (1.0f - mat.metalness);
                    const V3 numerator = (NDF * G) * F;
                    const float denominator = 4.0f * gl_max(dot3(sf.N, V), 0.0f) * gl_max(dot3(sf.N, L), 0.0f);
                    const V3 specular = numerator / gl_max(denominator, 0.001f);
                    const float NdotL = gl_max(dot3(sf.N, L), 0.0f);
                    e = e + (kD * mat.albedo / PI + specular) * attenuation * NdotL; // :502
                }
            }
            // :507  all(emissive > 0) ? normalize(emissive) : 0
            const bool all_pos = mat.emissive.x > 0.0f && mat.emissive.y > 0.0f && mat.emissive.z > 0.0f;
            const V3 emissive = all_pos ? normalize3(mat.emissive) : v3(0.0f);
            acc = acc + mask * (emissive + e);
            mask = mask * mat.albedo;
            ray = Ray{sf.P, normalize3(reflect3(ray.d, sf.N) + d)};                 // :511
        } else {                                                                    // MAT_TYPE_DIELECTRIC :514
            acc = acc + mat.emissive * mask;
            mask = mask * mat.albedo;
            const V3 normal = sf.N;
            const float nint = mat.roughness;
            const float cosine = -dot3(ray.d, sf.N) / length3(ray.d);               // :533
            const V3 reflected = reflect3(ray.d, sf.N);
            const V3 refracted = refract3(ray.d, normal, nint);
            const bool is_zero = refracted.x == 0.0f && refracted.y == 0.0f && refracted.z == 0.0f;
            const float p_reflect = is_zero ? 1.0f : schlick(cosine, mat.roughness); // :539
            ray = Ray{sf.P, normalize3(u01_(skey, dim0 + SLOT_R2) < p_reflect ? reflected : refracted)};
        }
        const float p = max3(mask);                                                 // :545
        if (u01_(skey, dim0 + SLOT_RR) > p) break;
        mask = mask * (1.0f / p);
    }
    return acc;
}

// ---- render_scene(), Raytracer.comp:280-355 ------------------------------------------------
V3 render_scene(const orc_scene &sc, uint32_t mode, Ray &ray, uint32_t &bounce_depth, uint32_t bounces,
                V3 light_pos, V3 cam_pos, Stats &st, uint32_t *primary_id)
{
    V3 color = v3(0.0f);
    Hit hit{1000.0f, 0, 0};                                                         // MAX_DISTANCE :88,290
    const bool found = trace_ray<false>(sc, mode, ray, hit, false, st);
    if (primary_id) *primary_id = found ? ((hit.kind << 28) | hit.index) : 0u;
    if (!found) return color;
    const Surface sf = surface_of(sc, ray, hit);
    const V3 light_vec = normalize3(light_pos - sf.P);
    const float dist_to_light = length3(light_pos - sf.P);
    {
        const float li = 540.0f / ((4.0f * 3.14159268f) * dist_to_light);           // :308
        const V3 light_intensity = v3(li);
        const V3 diffuse = light_intensity * sf.mat->albedo * gl_max(dot3(sf.N, light_vec), 0.0f);
        const V3 half_vec = normalize3(light_vec + normalize3(cam_pos));             // :314 (position as a direction)
        const V3 specular = light_intensity * pow_(gl_clamp(dot3(sf.N, half_vec), 0.0f, 1.0f), 16.0f);
        color = diffuse + specular;
    }
    {
        Hit sh{dist_to_light, 0, 0};
        if (trace_ray<false>(sc, mode, Ray{sf.P, light_vec}, sh, true, st)) {        // :329
            color = color * 0.5f;
            bounce_depth = bounces + 1;
        }
    }
    if (sf.mat->metalness >= 0.5f) {                                                // material.reflective :343
        ray.d = reflect3(ray.d, sf.N);                                              // :194-197
        ray.o = sf.P;
    } else {
        bounce_depth = bounces + 1;
    }
    return color;
}

// ---- primary ray, Tracer.comp:561-574 == Raytracer.comp:361-376 ----------------------------
inline Ray primary_ray(const orc_frame_data &fd, uint32_t w, uint32_t h, uint32_t x, uint32_t y)
{
    const float u = (float)x / (float)w, v = (float)y / (float)h;
    const float tx = 2.0f * u - 1.0f, ty = 2.0f * v - 1.0f;
    V3 dir = from3a(fd.camera.dir) + from3a(fd.camera.right) * tx + from3a(fd.camera.up) * ty;
    dir = dir * V3{fd.aspect_ratio, 1.0f, fd.aspect_ratio};                          // world x,z (!)
    return Ray{from3a(fd.camera.pos), normalize3(dir)};
}

inline uint8_t unorm8(float x)
{
    if (!(x == x)) return 0;
    const float c = gl_clamp(x, 0.0f, 1.0f);
    return (uint8_t)(int)floorf(c * 255.0f + 0.5f);
}

// ---- CPU LBVH ------------------------------------------------------------------------------
inline uint32_t expand_bits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
inline uint32_t quant10(float c, float cmin, float scale)
{
    float q = (c - cmin) * scale;
    q = fminf(fmaxf(q, 0.0f), 1023.0f);
    return (uint32_t)q;
}
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

} // namespace

// ============================================================================================
// C API
// ============================================================================================
extern "C" {

orc_scene *orc_scene_create(void) { return new orc_scene(); }
void orc_scene_destroy(orc_scene *s) { delete s; }

static void rebuild_lights(orc_scene *s)
{
    s->lights.clear();
    for (size_t i = 0; i < s->spheres.size(); ++i) {
        if (s->sphere_mat[i] >= s->mats.size()) continue;
        const Material &m = s->mats[s->sphere_mat[i]];
        if (!(m.emissive.x == 0.0f && m.emissive.y == 0.0f && m.emissive.z == 0.0f)) s->lights.push_back((uint32_t)i);
    }
}

int orc_scene_set_materials(orc_scene *s, const orc_material *m, uint32_t n)
{
    s->mats.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (m[i].type > 1u) return 6;
        s->mats[i] = Material{{m[i].albedo[0], m[i].albedo[1], m[i].albedo[2]},
                              {m[i].emissive[0], m[i].emissive[1], m[i].emissive[2]},
                              m[i].roughness, m[i].metalness, m[i].type};
    }
    rebuild_lights(s);
    return 0;
}
int orc_scene_set_spheres(orc_scene *s, const orc_sphere *sp, const uint32_t *mat_id, uint32_t n)
{
    s->spheres.assign(sp, sp + n);
    s->sphere_mat.assign(mat_id, mat_id + n);
    s->has_bvh = false; s->bvh.clear();
    rebuild_lights(s);
    return s->lights.size() > MAX_LIGHTS ? 6 : 0;
}
int orc_scene_set_planes(orc_scene *s, const orc_plane *p, const uint32_t *mat_id, uint32_t n)
{
    s->planes.assign(p, p + n);
    s->plane_mat.assign(mat_id, mat_id + n);
    return 0;
}
int orc_scene_set_triangles(orc_scene *s, const orc_triangle *t, uint32_t n, uint32_t mat_id)
{
    s->tris.assign(t, t + n);
    s->tri_mat = mat_id;
    s->tri_mats.clear();
    return 0;
}
int orc_scene_set_triangle_materials(orc_scene *s, const uint32_t *mat_ids, uint32_t n)
{
    if (n != s->tris.size() && n != 0) return 6;
    s->tri_mats.assign(mat_ids, mat_ids + n);
    return 0;
}

int orc_scene_use_default(orc_scene *s, uint32_t which)
{
    auto M = [](float ar, float ag, float ab, float e, float rough, float metal, uint32_t type) {
        orc_material m{};
        m.albedo[0] = ar; m.albedo[1] = ag; m.albedo[2] = ab;
        m.emissive[0] = m.emissive[1] = m.emissive[2] = e;
        m.roughness = rough; m.metalness = metal; m.type = type;
        return m;
    };
    // the one triangle the host uploads (Source/GraphicsDevice.cpp:798-803) == Raytracer.comp:116
    orc_triangle tri{};
    tri.v0 = {10.0f, 10.0f, 0.0f, 0.0f}; tri.v1 = {0.0f, 20.0f, 0.0f, 0.0f}; tri.v2 = {-10.0f, 10.0f, 0.0f, 0.0f};
    if (which == 0) {
        // Tracer.comp:186-194
        const orc_material mats[8] = {
            M(1.0f, 1.0f, 1.0f, 0.0f, 0.3f, 0.7f, 0),      // 0 matte_white
            M(0.75f, 0.25f, 0.25f, 0.0f, 0.4f, 0.0f, 0),   // 1 matte_red
            M(0.25f, 0.75f, 0.25f, 0.0f, 0.4f, 0.0f, 0),   // 2 matte_green
            M(0.25f, 0.25f, 0.75f, 0.0f, 0.4f, 0.0f, 0),   // 3 matte_blue
            M(0.25f, 0.25f, 0.75f, 0.0f, 0.3f, 0.6f, 0),   // 4 plastic
            M(1.0f, 0.5f, 0.5f, 0.0f, 0.0f, 1.0f, 0),      // 5 mirror
            M(1.0f, 1.0f, 1.0f, 0.0f, 0.42f, 0.0f, 1),     // 6 glass
            M(1.0f, 1.0f, 1.0f, 128.0f, 0.6f, 0.0f, 0)};   // 7 light
        orc_scene_set_materials(s, mats, 8);
        // Tracer.comp:196-202
        const orc_sphere sp[4] = {{42.0f, 16.0f, 12.0f, 16.0f}, {0.0f, 96.0f, 0.0f, 12.0f},
                                  {-32.0f, 24.0f, 24.0f, 24.0f}, {-24.0f, 11.0f, -48.0f, 11.0f}};
        const uint32_t spm[4] = {6, 7, 5, 4};
        orc_scene_set_spheres(s, sp, spm, 4);
        // Tracer.comp:204-211
        const orc_plane pl[5] = {{0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, -1.0f, 0.0f, 128.0f}, {1.0f, 0.0f, 0.0f, 64.0f},
                                 {0.0f, 0.0f, -1.0f, 64.0f}, {-1.0f, 0.0f, 0.0f, 64.0f}};
        const uint32_t plm[5] = {0, 0, 1, 2, 3};
        orc_scene_set_planes(s, pl, plm, 5);
        orc_scene_set_triangles(s, &tri, 1, 5);             // every triangle is `mirror` (Tracer.comp:386)
        return 0;
    }
    if (which == 1) {
        // Raytracer.comp:119-127  {reflective, diffuse}; reflective -> metalness 1 roughness 0, else 0 / 0.4
        auto R = [&](bool refl, float r, float g, float b) { return M(r, g, b, 0.0f, refl ? 0.0f : 0.4f, refl ? 1.0f : 0.0f, 0); };
        const orc_material mats[6] = {R(true, 1, 1, 1), R(false, 1, 0, 0), R(true, 0, 1, 0),
                                      R(false, 0, 0, 1), R(false, 1, 1, 0), R(false, 1, 0, 1)};
        orc_scene_set_materials(s, mats, 6);
        // Raytracer.comp:98-102
        const orc_sphere sp[2] = {{-14.0f, 12.0f, 32.0f, 5.0f}, {32.0f, 24.0f, 25.0f, 12.0f}};
        const uint32_t spm[2] = {5, 4};
        orc_scene_set_spheres(s, sp, spm, 2);
        // Raytracer.comp:104-112
        const orc_plane pl[5] = {{0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, -1.0f, 0.0f, 128.0f}, {0.0f, 0.0f, -1.0f, 64.0f},
                                 {1.0f, 0.0f, 0.0f, 64.0f}, {-1.0f, 0.0f, 0.0f, 64.0f}};
        const uint32_t plm[5] = {0, 0, 2, 1, 3};
        orc_scene_set_planes(s, pl, plm, 5);
        orc_scene_set_triangles(s, &tri, 1, 1);             // Raytracer.comp:116 mat_id 1
        return 0;
    }
    return 6;
}

// LBVH: centroid bounds -> 30-bit Morton -> sort by (code, index) -> Karras 2012 hierarchy -> refit.
// (New functionality; SURVEY.md section 0: the reference brute-forces every primitive.)
int orc_scene_build_bvh(orc_scene *s)
{
    const int n = (int)s->spheres.size();
    s->bvh.clear();
    s->has_bvh = true;
    if (n == 0) return 0;
    auto leaf_child = [&](int sphere) {
        BvhChild c{};
        V3 lo, hi; sphere_box(s->spheres[sphere], lo, hi);
        c.a[0] = lo.x; c.a[1] = lo.y; c.a[2] = lo.z; c.a[3] = hi.x;
        c.b[0] = hi.y; c.b[1] = hi.z; c.index = sphere; c.kind = 1;
        return c;
    };
    if (n == 1) { BvhNode nd; nd.c[0] = leaf_child(0); nd.c[1] = leaf_child(0); s->bvh.push_back(nd); return 0; }

    V3 cmin = v3(INFINITY), cmax = v3(-INFINITY);
    for (const orc_sphere &sp : s->spheres) {
        cmin = {fminf(cmin.x, sp.cx), fminf(cmin.y, sp.cy), fminf(cmin.z, sp.cz)};
        cmax = {fmaxf(cmax.x, sp.cx), fmaxf(cmax.y, sp.cy), fmaxf(cmax.z, sp.cz)};
    }
    const V3 ext = cmax - cmin;
    const V3 scale{ext.x > 0.0f ? 1024.0f / ext.x : 0.0f, ext.y > 0.0f ? 1024.0f / ext.y : 0.0f,
                   ext.z > 0.0f ? 1024.0f / ext.z : 0.0f};
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; ++i) {
        const orc_sphere &sp = s->spheres[i];
        const uint32_t code = (expand_bits(quant10(sp.cx, cmin.x, scale.x)) << 2) |
                              (expand_bits(quant10(sp.cy, cmin.y, scale.y)) << 1) |
                              expand_bits(quant10(sp.cz, cmin.z, scale.z));
        keys[i] = ((uint64_t)code << 32) | (uint32_t)i;
    }
    std::sort(keys.begin(), keys.end());
    auto code_of = [&](int i) { return (uint32_t)(keys[i] >> 32); };
    auto delta = [&](int i, int j) -> int {
        if (j < 0 || j >= n) return -1;
        const uint32_t a = code_of(i), b = code_of(j);
        if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
        return clz32(a ^ b);
    };
    struct Tmp { int child[2]; bool leaf[2]; V3 lo[2], hi[2]; };
    std::vector<Tmp> tmp(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        const int d = (delta(i, i + 1) - delta(i, i - 1)) > 0 ? 1 : -1;
        const int dmin = delta(i, i - d);
        int lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (delta(i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        const int dnode = delta(i, j);
        int sft = 0;
        for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
            if (delta(i, i + (sft + t) * d) > dnode) sft += t;
            if (t == 1) break;
        }
        const int gamma = i + sft * d + std::min(d, 0);
        const int first = std::min(i, j), last = std::max(i, j);
        tmp[i].child[0] = gamma;     tmp[i].leaf[0] = (first == gamma);
        tmp[i].child[1] = gamma + 1; tmp[i].leaf[1] = (last == gamma + 1);
    }
    // refit: post-order from the root (node 0)
    std::vector<V3> nlo(n - 1), nhi(n - 1);
    std::vector<int> order; order.reserve(n - 1);
    { std::vector<int> st{0};
      while (!st.empty()) { const int i = st.back(); st.pop_back(); order.push_back(i);
        for (int k = 0; k < 2; ++k) if (!tmp[i].leaf[k]) st.push_back(tmp[i].child[k]); } }
    for (int q = (int)order.size() - 1; q >= 0; --q) {
        const int i = order[q];
        for (int k = 0; k < 2; ++k) {
            if (tmp[i].leaf[k]) sphere_box(s->spheres[(uint32_t)keys[tmp[i].child[k]]], tmp[i].lo[k], tmp[i].hi[k]);
            else { tmp[i].lo[k] = nlo[tmp[i].child[k]]; tmp[i].hi[k] = nhi[tmp[i].child[k]]; }
        }
        nlo[i] = {fminf(tmp[i].lo[0].x, tmp[i].lo[1].x), fminf(tmp[i].lo[0].y, tmp[i].lo[1].y), fminf(tmp[i].lo[0].z, tmp[i].lo[1].z)};
        nhi[i] = {fmaxf(tmp[i].hi[0].x, tmp[i].hi[1].x), fmaxf(tmp[i].hi[0].y, tmp[i].hi[1].y), fmaxf(tmp[i].hi[0].z, tmp[i].hi[1].z)};
    }
    s->bvh.resize(n - 1);
    for (int i = 0; i < n - 1; ++i)
        for (int k = 0; k < 2; ++k) {
            if (tmp[i].leaf[k]) s->bvh[i].c[k] = leaf_child((int)(uint32_t)keys[tmp[i].child[k]]);
            else {
                BvhChild c{};
                c.a[0] = tmp[i].lo[k].x; c.a[1] = tmp[i].lo[k].y; c.a[2] = tmp[i].lo[k].z; c.a[3] = tmp[i].hi[k].x;
                c.b[0] = tmp[i].hi[k].y; c.b[1] = tmp[i].hi[k].z; c.index = tmp[i].child[k]; c.kind = 0;
                s->bvh[i].c[k] = c;
            }
        }
    return 0;
}
// Test hook: the oracle walks ANY hierarchy handed to it in the node layout of orc_scene_read_bvh (root = node 0).  Rule S
//  claims the answer does not depend on the hierarchy as long as leaves carry the spheres' own padded boxes and
// inner boxes are exact unions; tests feed random trees and the device's SAH traversal tree through this.
int orc_scene_set_bvh(orc_scene *s, const float *nodes, uint32_t n_nodes)
{
    if (!nodes && n_nodes) return 6;
    s->bvh.resize(n_nodes);
    if (n_nodes) std::memcpy(s->bvh.data(), nodes, (size_t)n_nodes * sizeof(BvhNode));
    s->has_bvh = true;
    return 0;
}
uint32_t orc_scene_bvh_nodes(const orc_scene *s) { return (uint32_t)s->bvh.size(); }
int orc_scene_read_bvh(const orc_scene *s, float *out, size_t bytes)
{
    if (bytes < s->bvh.size() * sizeof(BvhNode)) return 6;
    std::memcpy(out, s->bvh.data(), s->bvh.size() * sizeof(BvhNode));
    return 0;
}

int orc_resolve(const orc_params *p, const orc_frame_data *fd, const float *accum, uint8_t *rgba8)
{
    const uint32_t fkey = frame_key_(p->seed, fd->seed, p->frame_index);
    for (uint32_t y = 0; y < p->height; ++y)
        for (uint32_t x = 0; x < p->width; ++x) {
            const uint32_t pix = y * p->width + x;
            const float *a = accum + 4 * (size_t)pix;
            uint8_t *o = rgba8 + 4 * (size_t)pix;
            if (p->integrator == ORC_WHITTED) {                                     // Raytracer.comp:398
                o[0] = unorm8(a[0]); o[1] = unorm8(a[1]); o[2] = unorm8(a[2]); o[3] = 255;
                continue;
            }
            V3 c = V3{a[0], a[1], a[2]} / a[3];                                     // Tracer.comp:585
            c = c / (c + v3(1.0f));                                                 // :587
            c = {pow_(c.x, 1.0f / 2.2f), pow_(c.y, 1.0f / 2.2f), pow_(c.z, 1.0f / 2.2f)}; // :588
            const float dither = u01_(sample_key_(fkey, pix, DITHER_SAMPLE), 0) / 64.0f;  // :590
            o[0] = unorm8(c.x + dither); o[1] = unorm8(c.y + dither); o[2] = unorm8(c.z + dither); o[3] = 255;
        }
    return 0;
}

int orc_render(const orc_scene *sc, const orc_params *pp, const orc_frame_data *fd, float *accum,
               uint32_t *hit_ids, uint8_t *rgba8, orc_counters *counters)
{
    const orc_params p = *pp;
    if (!sc || !accum || p.width == 0 || p.height == 0) return 6;
    if (p.sphere_mode == ORC_SPHERES_S_BVH && !sc->has_bvh) return 6;
    for (uint32_t m : sc->sphere_mat) if (m >= sc->mats.size()) return 6;
    for (uint32_t m : sc->plane_mat) if (m >= sc->mats.size()) return 6;
    if (!sc->tris.empty() && sc->tri_mat >= sc->mats.size()) return 6;
    for (uint32_t m : sc->tri_mats) if (m >= sc->mats.size()) return 6;
    const uint32_t x0 = p.x0, y0 = p.y0;
    const uint32_t x1 = (p.x1 == 0 && p.x0 == 0) ? p.width : p.x1, y1 = (p.y1 == 0 && p.y0 == 0) ? p.height : p.y1;
    const uint32_t s0 = p.sample_begin, s1 = (p.sample_end == 0 && p.sample_begin == 0) ? p.spp : p.sample_end;
    const uint32_t fkey = frame_key_(p.seed, fd->seed, p.frame_index);
    const V3 cam_pos = from3a(fd->camera.pos), light_pos = from3a(fd->light_pos);
    unsigned nt = p.n_threads ? p.n_threads : std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    std::atomic<uint32_t> next_row{y0};
    std::vector<Stats> stats(nt);
    auto worker = [&](unsigned tid) {
        Stats st;                       // thread-local (no false sharing); published at the end
        for (;;) {
            const uint32_t y = next_row.fetch_add(1);
            if (y >= y1) break;
            for (uint32_t x = x0; x < x1; ++x) {
                const uint32_t pix = y * p.width + x;
                const Ray pr = primary_ray(*fd, p.width, p.height, x, y);
                uint32_t pid = 0;
                float *a = accum + 4 * (size_t)pix;
                if (p.integrator == ORC_WHITTED) {
                    // main(), Raytracer.comp:357-399
                    Ray ray = pr;
                    uint32_t bounce = 0;
                    V3 fin = render_scene(*sc, p.sphere_mode, ray, bounce, p.max_depth, light_pos, cam_pos, st, &pid);
                    float strength = 0.4f;
                    while (++bounce <= p.max_depth) {
                        const V3 refl = render_scene(*sc, p.sphere_mode, ray, bounce, p.max_depth, light_pos, cam_pos, st, nullptr);
                        fin = (1.0f - strength) * fin + strength * mix3(refl, fin, 1.0f - strength);   // :390
                        strength *= 0.5f;
                    }
                    a[0] = fin.x; a[1] = fin.y; a[2] = fin.z; a[3] = 1.0f;
                } else {
                    // main(), Tracer.comp:557-593: one primary ray reused by every sample (:574-581)
                    V3 sum = v3(0.0f);
                    for (uint32_t s = s0; s < s1; ++s) {
                        const uint32_t skey = sample_key_(fkey, pix, s);
                        sum = sum + radiance(*sc, p.sphere_mode, pr, cam_pos, p.max_depth, skey, st, s == s0 ? &pid : nullptr);
                    }
                    if (p.accumulate) { a[0] += sum.x; a[1] += sum.y; a[2] += sum.z; a[3] += (float)(s1 - s0); }
                    else { a[0] = sum.x; a[1] = sum.y; a[2] = sum.z; a[3] = (float)(s1 - s0); }
                }
                if (hit_ids) {
                    hit_ids[pix] = pid;
                    // tie-band report: does the literal chain rule pick another sphere than rule S?
                    if (p.sphere_mode != ORC_SPHERES_LITERAL) {
                        Stats dummy;
                        const float tmax0 = p.integrator == ORC_WHITTED ? 1000.0f : 3000.0f / pow_(1.0f, 2.0f);
                        Hit hl{tmax0, 0, 0};
                        const bool fl = p.integrator == ORC_WHITTED
                                            ? trace_ray<false>(*sc, ORC_SPHERES_LITERAL, pr, hl, false, dummy)
                                            : trace_ray<true>(*sc, ORC_SPHERES_LITERAL, pr, hl, false, dummy);
                        const uint32_t lid = fl ? ((hl.kind << 28) | hl.index) : 0u;
                        if (lid != pid) ++st.mism;
                    }
                }
            }
        }
        stats[tid] = st;
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    if (counters) {
        orc_counters c{};
        for (const Stats &s : stats) {
            c.closest_rays += s.closest; c.shadow_rays += s.shadow; c.node_visits += s.nodes;
            c.leaf_tests += s.leaves; c.paths += s.paths; c.literal_vs_s_mismatch += s.mism;
        }
        *counters = c;
    }
    if (rgba8) orc_resolve(&p, fd, accum, rgba8);
    return 0;
}

// ---- Fullscreen.frag:14-31 -- the present filter ----------------------------------
// Sampler: Source/GraphicsDevice.cpp:770-794 (bilinear, clamp-to-border, opaque black), normalised coordinates.
// Bilinear weights are implementation-defined in Vulkan; this restatement uses exact binary32 weights
// (texel centre at +0.5, mix() for both lerps) -- one legal execution.
static V3 sample_bilinear(const uint8_t *img, uint32_t w, uint32_t h, float u, float v)
{
    const float s = u * (float)w - 0.5f, t = v * (float)h - 0.5f;
    const float fs0 = floorf(s), ft0 = floorf(t);
    const float fx = s - fs0, fy = t - ft0;
    const int i0 = (int)fs0, j0 = (int)ft0;
    auto texel = [&](int i, int j) -> V3 {
        if (i < 0 || j < 0 || i >= (int)w || j >= (int)h) return v3(0.0f);       // VK_BORDER_COLOR_INT_OPAQUE_BLACK
        const uint8_t *p = img + 4 * ((size_t)j * w + (size_t)i);
        return V3{(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f};
    };
    const V3 a = mix3(texel(i0, j0), texel(i0 + 1, j0), fx), b = mix3(texel(i0, j0 + 1), texel(i0 + 1, j0 + 1), fx);
    return mix3(a, b, fy);
}

int orc_present(const uint8_t *binding0, const uint8_t *binding1, uint32_t tw, uint32_t th, uint8_t *out, uint32_t W, uint32_t H)
{
    const float inv_size = 1.0f / 2048.0f;                                         // Fullscreen.frag:12
    for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            // Fullscreen.vert:8-10: uv_coords spans [0,1] over the framebuffer, sampled at the pixel centre
            const float us = ((float)x + 0.5f) / (float)W, ut = ((float)y + 0.5f) / (float)H;
            const float u = us, v = 1.0f - ut;                                     // :16
            const V3 c0 = sample_bilinear(binding0, tw, th, u, v);
            const V3 c1 = sample_bilinear(binding1, tw, th, u, v);
            const V3 N = sample_bilinear(binding0, tw, th, u + 0.0f * inv_size, v + 1.0f * inv_size);
            const V3 S = sample_bilinear(binding0, tw, th, u + 0.0f * inv_size, v + -1.0f * inv_size);
            const V3 E = sample_bilinear(binding0, tw, th, u + 1.0f * inv_size, v + 0.0f * inv_size);
            const V3 Wc = sample_bilinear(binding0, tw, th, u + -1.0f * inv_size, v + 0.0f * inv_size);
            const V3 dlt = c0 - c1;
            const float vT = dot3(dlt, dlt);                                       // :26 temporal variance
            const V3 fin = vT > 0.0005f ? (((N + S) + E) + Wc) / 4.0f : c0;         // :28
            uint8_t *o = out + 4 * ((size_t)y * W + x);
            o[0] = unorm8(fin.x); o[1] = unorm8(fin.y); o[2] = unorm8(fin.z); o[3] = 255;
        }
    return 0;
}

float orc_sin(float x) { float s, c; sincos_(x, s, c); return s; }
float orc_cos(float x) { float s, c; sincos_(x, s, c); return c; }
float orc_exp2(float x) { return exp2_(x); }
float orc_log2(float x) { return log2_(x); }
float orc_pow(float x, float y) { return pow_(x, y); }
uint32_t orc_pcg_hash(uint32_t v) { return pcg(v); }
uint32_t orc_frame_key(uint64_t seed, float frame_seed, uint32_t frame_index) { return frame_key_(seed, frame_seed, frame_index); }
float orc_rand_u01(uint32_t frame_key, uint32_t pixel, uint32_t sample, uint32_t dim)
{
    return u01_(sample_key_(frame_key, pixel, sample), dim);
}
float orc_sphere_intersect(const float o[3], const float d[3], const orc_sphere *s)
{
    return sphere_intersect(Ray{fromp(o), fromp(d)}, *s);
}
float orc_plane_intersect_tracer(const float o[3], const float d[3], const orc_plane *p)
{
    return plane_intersect_tracer(Ray{fromp(o), fromp(d)}, *p);
}
float orc_plane_intersect_raytracer(const float o[3], const float d[3], const orc_plane *p)
{
    return plane_intersect_raytracer(Ray{fromp(o), fromp(d)}, *p);
}
float orc_tri_intersect(const float o[3], const float d[3], const orc_triangle *t, float eps)
{
    return tri_intersect(Ray{fromp(o), fromp(d)}, *t, eps);
}
void orc_primary_ray(const orc_frame_data *fd, uint32_t w, uint32_t h, uint32_t x, uint32_t y, float o[3], float d[3])
{
    const Ray r = primary_ray(*fd, w, h, x, y);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z;
}
int orc_slab(const float o[3], const float d[3], const float lo[3], const float hi[3], float *tn, float *tf)
{
    const SlabRay sr = slab_setup(Ray{fromp(o), fromp(d)});
    return slab_test(sr, fromp(lo), fromp(hi), *tn, *tf) ? 1 : 0;
}
int orc_query_spheres(const orc_scene *sc, uint32_t mode, const float o[3], const float d[3], float bound,
                      float epsilon, float *t_out)
{
    const Ray ray{fromp(o), fromp(d)};
    Stats st;
    if (mode == ORC_SPHERES_LITERAL) {
        float cur = bound - epsilon;            // so that cur + eps == bound up to rounding; tests pass bound = cur + eps
        const int i = literal_spheres_tracer(*sc, ray, epsilon, cur);
        if (t_out) *t_out = cur;
        return i;
    }
    const SBest b = mode == ORC_SPHERES_S_BVH ? s_query_bvh(*sc, ray, epsilon, bound, false, st)
                                              : s_query_linear(*sc, ray, epsilon, bound, false, st);
    if (t_out) *t_out = b.t;
    return b.idx;
}

} // extern "C"
