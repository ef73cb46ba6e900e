// Device-side scene layout, intersection routines, LBVH traversal and the two integrators'
// per-bounce bodies.  Shared by the megakernel and the wavefront kernels.
//
// Reference semantics restated here (paths relative to the reference root):
//   Assets/Tracer.comp:314-431    calc_*_intersect, trace_ray   (path integrator)
//   Assets/Tracer.comp:256-312    jitter, schlick, GGX, Smith
//   Assets/Tracer.comp:433-553    radiance
//   Assets/Raytracer.comp:129-355 calc_*_intersect, trace_ray, render_scene (whitted)
#pragma once
#include "vkrt_arith.cuh"
#include "../../include/vkrt.h"

namespace vkrt {

// BVH_STACK: entries of the per-thread traversal stacks.  The near-first binary walk holds at most one entry per tree
// level (+ the sentinel of the wavefront's stacks); the LBVH has at most 64 levels (vkrt_bvh.cu::k_tree_depth), the SAH
// traversal tree at most 48 + 32 (median index splits from level 48 on), and vkrt_build_bvh refuses a tree with
// depth + 2 > BVH_STACK, so the kernels push without a bound check.
enum { MAX_PLANES = 16, BVH_STACK = 128 };
enum { KIND_TRI = 1, KIND_SPHERE = 2, KIND_PLANE = 3 };

// ---- HBM layout (DESIGN.md "Data layout") ----------------------------------------------------
// spheres : float4 {cx,cy,cz,r}             16 B, original order (shading fetch: normal)
// mats    : 3 x float4 per material         48 B {albedo.rgb, roughness | emissive.rgb, metalness | type,-,-,-}
// tris    : 3 x float4 per triangle         48 B, the reference's SSBO layout verbatim
// bvh     : 4 x float4 per inner node       64 B = two 32 B child records:
//             {lo.x lo.y lo.z hi.x | hi.y hi.z index kind}; kind 0: inner node `index`; kind 1: leaf =
//             sphere `index`, the box being the sphere's own padded box (so both children are tested by
//             the same branch-free code; the sphere is fetched only when its box is hit)
struct DevScene {
    const float4 *spheres;
    const uint32_t *sphere_mat;
    const float4 *bvh;            // binary 64-byte nodes (two child records)
    const float4 *bvh4;           // 4-wide 128-byte traversal nodes (four child records), indexed by binary node id
    const uint4 *qbvh;            // 32-byte traversal nodes (two child records with 16-bit box coordinates)
    const float4 *tris;
    const float4 *mats;
    const float4 *tbvh;           // the triangles' own binary 64-byte nodes (rule T), or null: literal triangle loop
    const uint32_t *tri_mats;     // per-triangle material ids, or null: every triangle uses tri_mat (Tracer.comp:386)
    uint32_t n_spheres, n_tris, tri_mat, n_planes, n_lights, n_nodes, n_mats, n_tnodes;
    float4 planes[MAX_PLANES];
    uint32_t plane_mat[MAX_PLANES];
    uint32_t lights[MAX_LIGHTS];
    float qs[3], qb2[3];          // the 16-bit grid of qbvh: coordinate of code q = (2^23 + q) * qs + qb2
};

struct Material { V3 albedo; float roughness; V3 emissive; float metalness; uint32_t type; };

struct Stats { uint32_t closest, shadow, nodes, leaves, paths, skipped, shared; };
VKRT_DEV void stats_zero(Stats &s) { s.closest = s.shadow = s.nodes = s.leaves = s.paths = s.skipped = s.shared = 0; }

struct Hit { float t; uint32_t kind, index; };

VKRT_DEV Material load_material(const DevScene &sc, uint32_t id)
{
    const float4 a = __ldg(sc.mats + 3 * id), b = __ldg(sc.mats + 3 * id + 1);
    const uint32_t type = __float_as_uint(__ldg(&sc.mats[3 * id + 2].x));
    return Material{{a.x, a.y, a.z}, a.w, {b.x, b.y, b.z}, b.w, type};
}

// ---- Tracer.comp:314-329 / Raytracer.comp:163-178 ---------------------------------------------
VKRT_DEV float sphere_intersect(V3 o, V3 d, float4 s)
{
    const V3 oc = o - xyz(s);
    const float b = 2.0f * dot3(oc, d);
    const float c = fma_(-s.w, s.w, dot3(oc, oc));
    const float h = fma_(b, b, -(4.0f * c));
    if (h < 0.0f) return -1.0f;
    return (-b - sqrtf(h)) * 0.5f;
}
// ---- Tracer.comp:331-338 ----------------------------------------------------------------------
VKRT_DEV float plane_intersect_tracer(V3 o, V3 d, float4 p)
{
    const V3 N = xyz(p);
    const float dn = dot3(d, N);
    const float dist = -(p.w + dot3(o, N)) / dn;
    const float when_neq = gl_abs(gl_sign(dn - 0.0f));
    return when_neq * gl_max(dist, 0.0f);
}
// ---- Raytracer.comp:180-192 -------------------------------------------------------------------
VKRT_DEV float plane_intersect_raytracer(V3 o, V3 d, float4 p)
{
    const V3 N = xyz(p);
    const float dn = dot3(d, N);
    if (dn == 0.0f) return 0.0f;
    const float dist = -(p.w + dot3(o, N)) / dn;
    return gl_max(dist, 0.0f);
}
// ---- Tracer.comp:340-372 / Raytracer.comp:129-161 ---------------------------------------------
VKRT_DEV float tri_intersect(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float eps)
{
    const V3 v0v1 = v1 - v0, v0v2 = v2 - v0;
    const V3 pvec = cross3(d, v0v2);
    const float det = dot3(v0v1, pvec);
    if (det < eps) return -1.0f;
    const float inv_det = 1.0f / det;
    const V3 tvec = o - v0;
    const float u = dot3(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return -1.0f;
    const V3 qvec = cross3(tvec, v0v1);
    const float v = dot3(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return -1.0f;
    return dot3(v0v2, qvec) * inv_det;
}

// ---- rule S: order-independent nearest sphere through the LBVH (DESIGN.md "Rule S") ------------
VKRT_DEV float safe_inv(float d)
{
    const float dd = gl_abs(d) > 1e-20f ? d : copysignf(1e-20f, d);
    return 1.0f / dd;
}
struct SlabRay { V3 inv, oinv; };
VKRT_DEV SlabRay slab_setup(V3 o, V3 d)
{
    SlabRay s;
    s.inv = {safe_inv(d.x), safe_inv(d.y), safe_inv(d.z)};
    s.oinv = o * s.inv;
    return s;
}
VKRT_DEV bool slab_test(const SlabRay &s, V3 lo, V3 hi, float &tn, float &tf)
{
    const float t0x = fma_(lo.x, s.inv.x, -s.oinv.x), t1x = fma_(hi.x, s.inv.x, -s.oinv.x);
    const float t0y = fma_(lo.y, s.inv.y, -s.oinv.y), t1y = fma_(hi.y, s.inv.y, -s.oinv.y);
    const float t0z = fma_(lo.z, s.inv.z, -s.oinv.z), t1z = fma_(hi.z, s.inv.z, -s.oinv.z);
    tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    return tn <= tf && tf >= 0.0f;
}
VKRT_DEV float sphere_pad_radius(float r) { return r * 1.001f + 0.001f; }
// rule T: the padded box of a triangle, [min(v) - p, max(v) + p] with p = 0.001 + 0.001 * (largest extent)
VKRT_DEV void tri_padded_box(V3 a, V3 b, V3 c, V3 &lo, V3 &hi)
{
    lo = v3(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z));
    hi = v3(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
    const float p = 0.001f + 0.001f * fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
    lo = v3(lo.x - p, lo.y - p, lo.z - p);
    hi = v3(hi.x + p, hi.y + p, hi.z + p);
}

struct SBest { float t; int idx; };

VKRT_DEV void s_consider(V3 o, V3 d, float4 sph, int i, float tn, float eps, float B, SBest &best)
{
    const float t = sphere_intersect(o, d, sph);
    if (!(t > eps) || !(tn <= t)) return;
    if (best.idx < 0) { if (t < B) { best.t = t; best.idx = i; } }
    else if (t < best.t || (t == best.t && i < best.idx)) { best.t = t; best.idx = i; }
}

#ifndef VKRT_BVH4
#define VKRT_BVH4 0
#endif
VKRT_DEV void ldg256(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// Resumable traversal state: one call of trav_step visits one inner node (both child records).
// node < 0 means the traversal is finished.  The stack is a separate local-memory array (L1-resident)
// so that the scalar state below stays in registers (an array member would drag the whole struct
// into local memory).
struct Trav {
    SlabRay sr;
    SBest best;
    float eps, B;
    int node, sp;
};
VKRT_DEV void trav_init(Trav &tv, const DevScene &sc, V3 o, V3 d, float eps, float B)
{
    tv.sr = slab_setup(o, d);
    tv.best.t = B; tv.best.idx = -1;
    tv.eps = eps; tv.B = B;
    tv.sp = 0;
    tv.node = sc.n_nodes ? 0 : -1;
}
// the exact rule-S leaf test: the sphere's own padded box, then the reference's intersection formula
// the rule-T leaf test: the triangle's padded box, then the reference's Moeller-Trumbore routine (Tracer.comp:340-372)
template <bool STATS>
VKRT_DEV void tri_leaf_test(Trav &tv, const DevScene &sc, V3 o, V3 d, int ti, Stats &st)
{
    const V3 v0 = xyz(__ldg(sc.tris + 3 * ti)), v1 = xyz(__ldg(sc.tris + 3 * ti + 1)), v2 = xyz(__ldg(sc.tris + 3 * ti + 2));
    V3 lo, hi;
    tri_padded_box(v0, v1, v2, lo, hi);
    float tn, tf;
    if (slab_test(tv.sr, lo, hi, tn, tf) && tn <= tv.best.t) {
        if (STATS) ++st.leaves;
        const float t = tri_intersect(o, d, v0, v1, v2, tv.eps);
        if (!(t > tv.eps) || !(tn <= t)) return;
        if (tv.best.idx < 0) { if (t < tv.B) { tv.best.t = t; tv.best.idx = ti; } }
        else if (t < tv.best.t || (t == tv.best.t && ti < tv.best.idx)) { tv.best.t = t; tv.best.idx = ti; }
    }
}
template <bool STATS>
VKRT_DEV void leaf_test(Trav &tv, const DevScene &sc, V3 o, V3 d, int si, Stats &st)
{
    const float4 sph = __ldg(sc.spheres + si);
    const float rp = sphere_pad_radius(sph.w);
    float tn, tf;
    if (slab_test(tv.sr, v3(sph.x - rp, sph.y - rp, sph.z - rp), v3(sph.x + rp, sph.y + rp, sph.z + rp), tn, tf) && tn <= tv.best.t) {
        if (STATS) ++st.leaves;
        s_consider(o, d, sph, si, tn, tv.eps, tv.B, tv.best);
    }
}

VKRT_DEV void cswap(float &ta, int &ia, float &tb, int &ib)
{
    const bool s = tb < ta;
    const float t = s ? tb : ta; tb = s ? ta : tb; ta = t;
    const int i = s ? ib : ia; ib = s ? ia : ib; ia = i;
}

template <bool ANY, bool STATS, bool TRI = false>
VKRT_DEV void trav_step(Trav &tv, int *__restrict__ stack, const DevScene &sc, V3 o, V3 d, Stats &st)
{
#if VKRT_BVH4
    static_assert(!TRI, "the 4-wide option has no triangle tree");
    // one 128-byte node = four child records {lo.xyz hi.x | hi.yz index kind} = four 256-bit loads in flight at once
    const float4 *np = sc.bvh4 + 8 * (size_t)tv.node;
    float4 a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ldg256(np + 2 * k, a[k], b[k]);
    if (STATS) ++st.nodes;
    const float INF = __int_as_float(0x7f800000);
    float tn[4]; int idx[4]; bool hit[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float tf;
        hit[k] = slab_test(tv.sr, v3(a[k].x, a[k].y, a[k].z), v3(a[k].w, b[k].x, b[k].y), tn[k], tf) && tn[k] <= tv.best.t &&
                 __float_as_int(b[k].w) != 2;
        idx[k] = __float_as_int(b[k].z);
    }
    // leaf children: the stored box is the sphere's own padded box; fetch the sphere and run the rule-S leaf test
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (hit[k] && __float_as_int(b[k].w) == 1) { leaf_test<STATS>(tv, sc, o, d, idx[k], st); hit[k] = false; }
    if (ANY && tv.best.idx >= 0) { tv.node = -1; return; }
    // inner children still within reach, nearest first: descend into the nearest, push the others far -> near
    int c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { hit[k] = hit[k] && tn[k] <= tv.best.t; tn[k] = hit[k] ? tn[k] : INF; c += hit[k] ? 1 : 0; }
    cswap(tn[0], idx[0], tn[1], idx[1]); cswap(tn[2], idx[2], tn[3], idx[3]);
    cswap(tn[0], idx[0], tn[2], idx[2]); cswap(tn[1], idx[1], tn[3], idx[3]);
    cswap(tn[1], idx[1], tn[2], idx[2]);
    if (c > 3) stack[tv.sp++] = idx[3];
    if (c > 2) stack[tv.sp++] = idx[2];
    if (c > 1) stack[tv.sp++] = idx[1];
    if (c > 0) tv.node = idx[0];
    else tv.node = tv.sp ? stack[--tv.sp] : -1;
#else
    const float4 *np = (TRI ? sc.tbvh : sc.bvh) + 4 * (size_t)tv.node;
    float4 a0, b0, a1, b1;
    ldg256(np, a0, b0);
    ldg256(np + 2, a1, b1);
    if (STATS) ++st.nodes;
    float tn0, tn1, tf;
    bool h0 = slab_test(tv.sr, v3(a0.x, a0.y, a0.z), v3(a0.w, b0.x, b0.y), tn0, tf) && tn0 <= tv.best.t;
    bool h1 = slab_test(tv.sr, v3(a1.x, a1.y, a1.z), v3(a1.w, b1.x, b1.y), tn1, tf) && tn1 <= tv.best.t;
    const int i0 = __float_as_int(b0.z), i1 = __float_as_int(b1.z);
    if (h0 && __float_as_int(b0.w) != 0) { if (TRI) tri_leaf_test<STATS>(tv, sc, o, d, i0, st); else leaf_test<STATS>(tv, sc, o, d, i0, st); h0 = false; }
    if (h1 && __float_as_int(b1.w) != 0) { if (TRI) tri_leaf_test<STATS>(tv, sc, o, d, i1, st); else leaf_test<STATS>(tv, sc, o, d, i1, st); h1 = false; }
    if (ANY && tv.best.idx >= 0) { tv.node = -1; return; }
    const bool both = h0 && h1;
    const bool take1 = both ? (tn1 < tn0) : h1;
    const int nearer = take1 ? i1 : i0;
    if (both) stack[tv.sp++] = take1 ? i0 : i1;
    if (h0 || h1) tv.node = nearer;
    else tv.node = tv.sp ? stack[--tv.sp] : -1;
#endif
}

// ---- phase-split traversal (wavefront trace kernel) ------------------------------------------------
// The same walk as trav_step, cut into two kinds of step so that a warp can run each kind with many lanes:
// an INNER step visits one inner node and only *schedules* the leaf children whose box it hit (they travel
// through `node` / the stack encoded as ~sphere), a LEAF step runs the exact rule-S leaf test of one scheduled
// sphere.  Per ray the sequence of box tests, leaf tests and culls is a fixed near-first order, so the result is
// the same as any other order (rule S); what changes is that leaf tests of different lanes execute together
// instead of one or two lanes at a time inside the node loop (measured there: 1.8 active lanes, 1/3 of the
// kernel's instructions).
enum : int { TRAV_DONE = (int)0x80000000 };
// Per-lane traversal stack: the first VKRT_SMEM_STACK entries live in shared memory laid out [entry][thread]
// (lane i always hits bank i: one wavefront per push/pop however the lanes' depths differ), deeper entries
// spill to a local array.
#ifndef VKRT_SMEM_STACK
#define VKRT_SMEM_STACK 0
#endif
template <int BLOCK>
struct TravStack {
    int *sm;                 // &shared[0][threadIdx.x]
    int lm[BVH_STACK];
    VKRT_DEV void push(int &sp, int v)
    {
        if (VKRT_SMEM_STACK > 0 && sp < VKRT_SMEM_STACK) sm[sp * BLOCK] = v;
        else lm[sp - VKRT_SMEM_STACK] = v;
        ++sp;
    }
    VKRT_DEV int pop(int &sp)
    {
        --sp;
        if (VKRT_SMEM_STACK > 0 && sp < VKRT_SMEM_STACK) return sm[sp * BLOCK];
        return lm[sp - VKRT_SMEM_STACK];
    }
};
// ---- 32-byte nodes (VKRT_QNODES) ----------------------------------------------------------------------
// The trace kernels are bound by L1TEX wavefronts of scattered node fetches (one per lane and 32-byte
// sector), so the inner step reads ONE sector per node: two child records {x: lo | hi << 16, y, z, ref} whose
// 16-bit codes q stand for the real coordinate X(q) = (2^23 + q) * qs + qb2.  The builder picks codes with
// X(q_lo) <= lo and X(q_hi) >= hi of the exact box (vkrt_bvh.cu::k_quantize), monotonically, so coded boxes
// nest like the exact ones.  The slab test works on the codes directly: one PRMT turns a code into the float
// 2^23 + q, one directed-rounding FMA with per-ray constants gives a LOWER bound of the entry parameter
// (fma.rd, C_dn) and an UPPER bound of the exit parameter (fma.ru, C_up) of the real-arithmetic slab test of
// the coded box -- hence of the float slab test of every exact box inside it (a float <= a real x is <= RN(x)).
// An ancestor is therefore never rejected when the exact rule-S test of a sphere below it passes, and rule S's
// answer is unchanged (DESIGN.md "Rule S"); a coded box only ever lets a few more rays through.
#ifndef VKRT_QNODES
#define VKRT_QNODES 1
#endif
#ifndef VKRT_PRMT_ASM
#define VKRT_PRMT_ASM 1      // 1: PRMT by inline PTX (no `& 0x7777` that __byte_perm's definition makes the compiler emit)
#endif
#ifndef VKRT_QSEL6
#define VKRT_QSEL6 0         // 1: the FAR plane's selectors live in registers too (no `^ 0x22` per node and axis)
#endif
#ifndef VKRT_SENTINEL
#define VKRT_SENTINEL 2      // > 0: the stack's bottom entry is TRAV_DONE: a pop needs no "stack empty?" test (2, the default: 26.54 ->
#endif                       //    26.47 ms/frame); 1: the inner step also stores its far child unconditionally -- no branches (slower)
#ifndef VKRT_SELECT2
#define VKRT_SELECT2 1       // 1: near/far child chosen by one predicate and two selects
#endif
struct QRay {
    V3 a, cdn, cup; uint32_t selx, sely, selz;
#if VKRT_QSEL6
    uint32_t fselx, fsely, fselz;
#endif
};
// float 2^23 + q of one 16-bit code of the packed pair `w`: sel = 0x7610 takes the low code, 0x7632 the high one.
// Every selector nibble is <= 7, so the raw PRMT equals __byte_perm (which is defined on `sel & 0x7777`).
VKRT_DEV float qcode(uint32_t w, uint32_t sel)
{
#if VKRT_PRMT_ASM
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(sel));
    return __uint_as_float(r);
#else
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel));
#endif
}
VKRT_DEV QRay qray_setup(const DevScene &sc, const SlabRay &sr)
{
    QRay q;
    const float inv[3] = {sr.inv.x, sr.inv.y, sr.inv.z}, oinv[3] = {sr.oinv.x, sr.oinv.y, sr.oinv.z};
    float a[3], cdn[3], cup[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = __fmul_rn(sc.qs[k], inv[k]);
        // |(2^23 + q) * (a - qs * inv)| < 0.51 |a| for q < 2^16
        const float m = __fmul_ru(0.51f, gl_abs(a[k]));
        cdn[k] = __fsub_rd(__fmaf_rd(sc.qb2[k], inv[k], -oinv[k]), m);
        cup[k] = __fadd_ru(__fmaf_ru(sc.qb2[k], inv[k], -oinv[k]), m);
    }
    q.a = v3(a[0], a[1], a[2]); q.cdn = v3(cdn[0], cdn[1], cdn[2]); q.cup = v3(cup[0], cup[1], cup[2]);
    // PRMT selector of the NEAR plane's code: the low half (lo) for a positive direction, the high half (hi) else
    q.selx = inv[0] < 0.0f ? 0x7632u : 0x7610u;
    q.sely = inv[1] < 0.0f ? 0x7632u : 0x7610u;
    q.selz = inv[2] < 0.0f ? 0x7632u : 0x7610u;
#if VKRT_QSEL6
    q.fselx = q.selx ^ 0x22u; q.fsely = q.sely ^ 0x22u; q.fselz = q.selz ^ 0x22u;
#endif
    return q;
}
VKRT_DEV void ldg256u(const uint4 *p, uint4 &a, uint4 &b)
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
VKRT_DEV bool qslab_test(const QRay &q, uint4 w, float &tn)
{
#if VKRT_QSEL6
    const uint32_t fsx = q.fselx, fsy = q.fsely, fsz = q.fselz;
#else
    const uint32_t fsx = q.selx ^ 0x22u, fsy = q.sely ^ 0x22u, fsz = q.selz ^ 0x22u;
#endif
    const float nx = qcode(w.x, q.selx), fx = qcode(w.x, fsx);
    const float ny = qcode(w.y, q.sely), fy = qcode(w.y, fsy);
    const float nz = qcode(w.z, q.selz), fz = qcode(w.z, fsz);
    tn = fmaxf(fmaxf(__fmaf_rd(nx, q.a.x, q.cdn.x), __fmaf_rd(ny, q.a.y, q.cdn.y)), __fmaf_rd(nz, q.a.z, q.cdn.z));
    const float tf = fminf(fminf(__fmaf_ru(fx, q.a.x, q.cup.x), __fmaf_ru(fy, q.a.y, q.cup.y)), __fmaf_ru(fz, q.a.z, q.cup.z));
    return tn <= tf && tf >= 0.0f;
}
template <bool STATS, class Stack>
VKRT_DEV void trav_inner_step_q(Trav &tv, const QRay &q, Stack &stack, const DevScene &sc, Stats &st)
{
    uint4 w0, w1;
    ldg256u(sc.qbvh + 2 * (size_t)tv.node, w0, w1);
    if (STATS) ++st.nodes;
    float tn0, tn1;
    const bool h0 = qslab_test(q, w0, tn0) && tn0 <= tv.best.t;
    const bool h1 = qslab_test(q, w1, tn1) && tn1 <= tv.best.t;
    const int c0 = (int)w0.w, c1 = (int)w1.w;          // inner node index, or ~sphere for a leaf
#if VKRT_SELECT2
    // the same choice as below (both: the nearer child first, ties -> child 0) from one predicate and two selects
    const bool take1 = h1 && (!h0 || tn1 < tn0);
    const int nearer = take1 ? c1 : c0, other = take1 ? c0 : c1;
#if VKRT_SENTINEL == 2 && !VKRT_SMEM_STACK
    if (h0 && h1) stack.push(tv.sp, other);
    if (h0 || h1) tv.node = nearer;
    else tv.node = stack.lm[--tv.sp];             // the bottom entry is TRAV_DONE: no "stack empty?" test
#elif VKRT_SENTINEL && !VKRT_SMEM_STACK
    stack.lm[tv.sp] = other;                      // kept only if both children were hit
    tv.sp += (h0 && h1) ? 1 : 0;
    tv.node = nearer;
    if (!(h0 || h1)) tv.node = stack.lm[--tv.sp]; // the bottom entry is TRAV_DONE
#else
    if (h0 && h1) stack.push(tv.sp, other);
    if (h0 || h1) tv.node = nearer;
    else tv.node = tv.sp ? stack.pop(tv.sp) : (int)TRAV_DONE;
#endif
#else
    const bool both = h0 && h1;
    const bool take1 = both ? (tn1 < tn0) : h1;
    if (both) stack.push(tv.sp, take1 ? c0 : c1);
    if (h0 || h1) tv.node = take1 ? c1 : c0;
    else tv.node = tv.sp ? stack.pop(tv.sp) : (int)TRAV_DONE;
#endif
}

template <bool STATS, class Stack>
VKRT_DEV void trav_inner_step(Trav &tv, Stack &stack, const DevScene &sc, Stats &st)
{
    const float4 *np = sc.bvh + 4 * (size_t)tv.node;
    float4 a0, b0, a1, b1;
    ldg256(np, a0, b0);
    ldg256(np + 2, a1, b1);
    if (STATS) ++st.nodes;
    float tn0, tn1, tf;
    const bool h0 = slab_test(tv.sr, v3(a0.x, a0.y, a0.z), v3(a0.w, b0.x, b0.y), tn0, tf) && tn0 <= tv.best.t;
    const bool h1 = slab_test(tv.sr, v3(a1.x, a1.y, a1.z), v3(a1.w, b1.x, b1.y), tn1, tf) && tn1 <= tv.best.t;
    // child reference: inner node index, or ~sphere for a leaf (kind is 0 / 1)
    const int c0 = __float_as_int(b0.z) ^ -__float_as_int(b0.w), c1 = __float_as_int(b1.z) ^ -__float_as_int(b1.w);
    const bool both = h0 && h1;
    const bool take1 = both ? (tn1 < tn0) : h1;
    if (both) stack.push(tv.sp, take1 ? c0 : c1);
    if (h0 || h1) tv.node = take1 ? c1 : c0;
    else tv.node = tv.sp ? stack.pop(tv.sp) : (int)TRAV_DONE;
}
template <bool STATS, class Stack>
VKRT_DEV void trav_leaf_step(Trav &tv, Stack &stack, const DevScene &sc, V3 o, V3 d, bool any, Stats &st)
{
    leaf_test<STATS>(tv, sc, o, d, ~tv.node, st);
    if (any && tv.best.idx >= 0) { tv.node = TRAV_DONE; return; }
#if VKRT_SENTINEL && !VKRT_SMEM_STACK
    tv.node = stack.lm[--tv.sp];
#else
    tv.node = tv.sp ? stack.pop(tv.sp) : (int)TRAV_DONE;
#endif
}

template <bool ANY, bool STATS>
VKRT_DEV SBest bvh_query(const DevScene &sc, V3 o, V3 d, float eps, float B, Stats &st)
{
    Trav tv;
    int stack[BVH_STACK];
    trav_init(tv, sc, o, d, eps, B);
    while (tv.node >= 0) trav_step<ANY, STATS>(tv, stack, sc, o, d, st);
    return tv.best;
}
// the same walk through the triangles' tree (rule T)
template <bool ANY>
VKRT_DEV SBest tri_bvh_query(const DevScene &sc, V3 o, V3 d, float eps, float B)
{
    Trav tv;
    int stack[BVH_STACK];
    Stats st;
    trav_init(tv, sc, o, d, eps, B);
    tv.node = sc.n_tnodes ? 0 : -1;
    while (tv.node >= 0) trav_step<ANY, false, true>(tv, stack, sc, o, d, st);
    return tv.best;
}

// ---- trace_ray: Tracer.comp:374-431 (TRACER = true) / Raytracer.comp:224-278 (false) ----------
// SHADOW queries only need the boolean, so they may leave early.
// Scenes with a hierarchy (vkrt_build_bvh) hold their triangles in a tree of their own and query it by the
// order-independent rule T (same exclusive bound as the literal loop's first acceptance); ANY = a shadow query, which
// may stop at the first hit.  Scenes without one run the reference's literal loop.
// TB = the kernel was instantiated for scenes that MAY hold a triangle tree: the hot kernels exist in both forms and
// scenes without one (every BASELINE configuration) run the form that does not carry the tree walk's registers, spills
// and second stack (measured: 1.8 % of the cfg4 frame, 7 % of the megakernel on the default scene).
template <bool TRACER, bool ANY = false, bool TB = true>
VKRT_DEV bool trace_tris(const DevScene &sc, V3 o, V3 d, float &cur, Hit &hit)
{
    const float EPS = TRACER ? 1e-3f : 0.01f;
#ifndef VKRT_NO_TRI_BVH      // (measurement aid: the hot kernels without the triangle tree's code)
    if (TB && sc.tbvh) {
        const SBest b = tri_bvh_query<ANY>(sc, o, d, EPS, TRACER ? cur + EPS : cur);
        if (b.idx < 0) return false;
        cur = b.t; hit.kind = KIND_TRI; hit.index = (uint32_t)b.idx;
        return true;
    }
#endif
    bool found = false;
    for (uint32_t i = 0; i < sc.n_tris; ++i) {
        const V3 v0 = xyz(__ldg(sc.tris + 3 * i)), v1 = xyz(__ldg(sc.tris + 3 * i + 1)), v2 = xyz(__ldg(sc.tris + 3 * i + 2));
        const float t = tri_intersect(o, d, v0, v1, v2, EPS);
        const bool acc = TRACER ? ((t > EPS) && (t < cur + EPS)) : (t > EPS && t < cur);
        if (acc) { cur = t; hit.kind = KIND_TRI; hit.index = i; found = true; }
    }
    return found;
}
#ifndef VKRT_PLANE_ILP
#define VKRT_PLANE_ILP 0
#endif
template <bool TRACER>
VKRT_DEV bool trace_planes(const DevScene &sc, V3 o, V3 d, float &cur, Hit &hit)
{
    const float EPS = TRACER ? 1e-3f : 0.01f;
    bool found = false;
#if VKRT_PLANE_ILP
    // the same tests on the same values, four planes at a time: the two dot products of a plane do not depend on `cur`,
    // so the constant-bank loads and FMA chains of four planes overlap; only the acceptance chain stays sequential
    for (uint32_t i0 = 0; i0 < sc.n_planes; i0 += 4u) {
        float dn[4], num[4];
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t i = i0 + k < sc.n_planes ? i0 + k : sc.n_planes - 1u;
            const float4 p = sc.planes[i];
            const V3 N = xyz(p);
            dn[k] = dot3(d, N);
            num[k] = -(p.w + dot3(o, N));
        }
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            if (i0 + k >= sc.n_planes) break;
            const bool same_sign = (num[k] > 0.0f && dn[k] > 0.0f) || (num[k] < 0.0f && dn[k] < 0.0f);
            if (same_sign && gl_abs(num[k]) > (cur * gl_abs(dn[k])) * 1.000001f) continue;
            if (!same_sign && num[k] == num[k] && dn[k] == dn[k]) continue;
            float t;
            if (TRACER) t = gl_abs(gl_sign(dn[k] - 0.0f)) * gl_max(num[k] / dn[k], 0.0f);      // plane_intersect_tracer on the same dn, num
            else t = dn[k] == 0.0f ? 0.0f : gl_max(num[k] / dn[k], 0.0f);                          // plane_intersect_raytracer
            const bool acc = TRACER ? ((t > EPS) && (t < cur - EPS)) : (t > EPS && t < cur);
            if (acc) { cur = t; hit.kind = KIND_PLANE; hit.index = i0 + k; found = true; }
        }
    }
#else
    for (uint32_t i = 0; i < sc.n_planes; ++i) {
        // t = -(len + o.N) / (d.N) clamped at 0 is accepted iff EPS < t < cur (-EPS).  Two rejections need no
        // IEEE division and are exact: (a) the quotient's sign is the XOR of the operands' signs, so unless
        // num and dn are both non-zero with equal signs t is 0 / NaN -> rejected; (b) if |num| exceeds
        // cur * |dn| by more than a few ulps the correctly rounded quotient is >= cur -> rejected.
        const float4 p = sc.planes[i];
        const V3 N = xyz(p);
        const float dn = dot3(d, N);
        const float num = -(p.w + dot3(o, N));
        const bool same_sign = (num > 0.0f && dn > 0.0f) || (num < 0.0f && dn < 0.0f);
        if (same_sign && gl_abs(num) > (cur * gl_abs(dn)) * 1.000001f) continue;
        if (!same_sign && num == num && dn == dn) continue;
        const float t = TRACER ? plane_intersect_tracer(o, d, p) : plane_intersect_raytracer(o, d, p);
        const bool acc = TRACER ? ((t > EPS) && (t < cur - EPS)) : (t > EPS && t < cur);
        if (acc) { cur = t; hit.kind = KIND_PLANE; hit.index = i; found = true; }
    }
#endif
    return found;
}
// upper bound handed to the sphere query after the triangle loop (rule S: exclusive)
template <bool TRACER> VKRT_DEV float sphere_bound(float cur) { return TRACER ? cur + 1e-3f : cur; }
template <bool TRACER> VKRT_DEV float trace_eps() { return TRACER ? 1e-3f : 0.01f; }

template <bool TRACER, bool BVH, bool SHADOW, bool STATS>
VKRT_DEV bool trace_ray(const DevScene &sc, V3 o, V3 d, Hit &hit, Stats &st)
{
    const float EPS = trace_eps<TRACER>();
    if (SHADOW) ++st.shadow; else ++st.closest;
    float cur = hit.t;
    bool found = trace_tris<TRACER, SHADOW, BVH>(sc, o, d, cur, hit);      // a triangle tree only exists next to the spheres' (vkrt_build_bvh)
    if (SHADOW && found) return true;
    if (BVH) {
        const SBest b = bvh_query<SHADOW, STATS>(sc, o, d, EPS, sphere_bound<TRACER>(cur), st);
        if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
    } else {
        for (uint32_t i = 0; i < sc.n_spheres; ++i) {
            const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
            const bool acc = TRACER ? ((t > EPS) && (t < cur + EPS)) : (t > EPS && t < cur);
            if (acc) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
        }
    }
    if (SHADOW && found) return true;
    found = trace_planes<TRACER>(sc, o, d, cur, hit) || found;
    hit.t = cur;
    return found;
}

struct Surface { V3 P, N; uint32_t mat; };

VKRT_DEV Surface surface_of(const DevScene &sc, V3 o, V3 d, const Hit &hit)
{
    Surface s;
    s.P = madd3(hit.t, d, o);
    if (hit.kind == KIND_TRI) {
        const V3 v0 = xyz(__ldg(sc.tris + 3 * hit.index)), v1 = xyz(__ldg(sc.tris + 3 * hit.index + 1)),
                 v2 = xyz(__ldg(sc.tris + 3 * hit.index + 2));
        s.N = cross3(v1 - v0, v2 - v0);      // unnormalised (Tracer.comp:389-392)
        s.mat = sc.tri_mats ? __ldg(sc.tri_mats + hit.index) : sc.tri_mat;
    } else if (hit.kind == KIND_SPHERE) {
        const float4 sp = __ldg(sc.spheres + hit.index);
        s.N = (s.P - xyz(sp)) / sp.w;         // Tracer.comp:408
        s.mat = __ldg(sc.sphere_mat + hit.index);
    } else {
        s.N = xyz(sc.planes[hit.index]);
        s.mat = sc.plane_mat[hit.index];
    }
    return s;
}

// ---- Tracer.comp:256-312 ----------------------------------------------------------------------
VKRT_DEV V3 jitter(V3 d, float phi, float sina, float cosa)
{
    const V3 w = normalize3(d);
    const V3 u = normalize3(cross3(v3(w.y, w.z, w.x), w));
    const V3 v = cross3(w, u);
    float s, c;
    sincos_(phi, s, c);
    return (u * c + v * s) * sina + w * cosa;
}
VKRT_DEV float schlick(float cosine, float ior)
{
    float r0 = (1.0f - ior) / (1.0f + ior);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * pow_(1.0f - cosine, 5.0f);
}
VKRT_DEV V3 fresnel_schlick(float cosTheta, V3 F0)
{
    const float p = pow_(1.0f - cosTheta, 5.0f);
    return F0 + (v3(1.0f) - F0) * p;
}
VKRT_DEV float distribution_ggx(V3 N, V3 H, float roughness)
{
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float NdotH = gl_max(dot3(N, H), 0.0f);
    const float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = VKRT_PI * denom * denom;
    return a2 / denom;
}
VKRT_DEV float geometry_schlick_ggx(float NdotV, float roughness)
{
    const float r = (roughness + 1.0f);
    const float k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
VKRT_DEV float geometry_smith(V3 N, V3 V, V3 L, float roughness)
{
    const float NdotV = gl_max(dot3(N, V), 0.0f);
    const float NdotL = gl_max(dot3(N, L), 0.0f);
    const float ggx2 = geometry_schlick_ggx(NdotV, roughness);
    const float ggx1 = geometry_schlick_ggx(NdotL, roughness);
    return ggx1 * ggx2;
}
VKRT_DEV float max3(V3 e) { return gl_max(gl_max(e.x, e.y), e.z); }

// ---- one iteration of radiance()'s depth loop, Tracer.comp:438-550 ----------------------------
struct PathState { V3 o, d, acc, mask; uint32_t depth; };

VKRT_DEV void path_begin(PathState &ps, V3 o, V3 d)
{
    ps.o = o; ps.d = d; ps.acc = v3(0.0f); ps.mask = v3(1.0f); ps.depth = 0;
}
VKRT_DEV float path_tmax(uint32_t depth) { return 3000.0f / pow_((float)(depth + 1u), 2.0f); }   // :444

// the cone sample towards emissive sphere l (Tracer.comp:464-469): direction L and distance bound t
VKRT_DEV void nee_sample(const DevScene &sc, V3 P, uint32_t l, uint32_t skey, uint32_t dim0, V3 &L, float &t)
{
    const float4 s = __ldg(sc.spheres + sc.lights[l]);
    const V3 sP = xyz(s);
    t = length3(sP - P) - s.w;
    const V3 l0 = sP - P;
    const float cos_a_max = sqrtf(1.0f - gl_clamp(s.w * s.w / dot3(l0, l0), 0.0f, 1.0f));
    const float cosa = gl_mix(cos_a_max, 1.0f, u01(skey, dim0 + SLOT_LIGHT + 2 * l));
    L = jitter(l0, VKRT_TWO_PI * u01(skey, dim0 + SLOT_LIGHT + 2 * l + 1), sqrtf(1.0f - cosa * cosa), cosa);
}

// what light l adds to `e` when it is not occluded (Tracer.comp:475-502)
VKRT_DEV V3 light_term(const DevScene &sc, const Surface &sf, const Material &mat, V3 cam_pos, uint32_t l, V3 L, float t)
{
    const uint32_t li = sc.lights[l];
    const float sr = __ldg(&sc.spheres[li].w);
    const V3 semis = xyz(__ldg(sc.mats + 3 * __ldg(sc.sphere_mat + li) + 1));
    V3 attenuation = semis * 1.0f / pow_(t / sr + 1.0f, 2.0f);
    attenuation = (attenuation - v3(0.001f)) / (1.0f - 0.001f);
    attenuation = v3(gl_max(attenuation.x, 0.0f), gl_max(attenuation.y, 0.0f), gl_max(attenuation.z, 0.0f));
    V3 F0 = v3(0.04f);
    F0 = mix3(F0, mat.albedo, mat.metalness);
    const V3 V = normalize3(cam_pos - sf.P);
    const V3 H = normalize3(V + L);
    const float NDF = distribution_ggx(sf.N, H, mat.roughness);
    const float G = geometry_smith(sf.N, V, L, mat.roughness);
    const V3 F = fresnel_schlick(gl_max(dot3(H, V), 0.0f), F0);
    V3 kD = v3(1.0f) - F;
    kD = kD * (1.0f - mat.metalness);
    const V3 numerator = (NDF * G) * F;
    const float denominator = 4.0f * gl_max(dot3(sf.N, V), 0.0f) * gl_max(dot3(sf.N, L), 0.0f);
    const V3 specular = numerator / gl_max(denominator, 0.001f);
    const float NdotL = gl_max(dot3(sf.N, L), 0.0f);
    return (kD * mat.albedo / VKRT_PI + specular) * attenuation * NdotL;
}
// A term that is exactly zero in all three channels (typically N.L <= 0: the light is behind the surface)
// cannot change `e` whether the shadow ray is occluded or not, so that ray need not be traced; a NaN term
// is not "zero" and is traced like any other.  The ray still counts as a trace_ray invocation.
VKRT_DEV bool term_is_zero(V3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }

// light evaluation used by the megakernel: sample, evaluate, trace the shadow ray in place (Tracer.comp:464-503).
// Returns what is added to `e` (+0 when occluded: e + 0 == e bit for bit).
template <bool BVH, bool STATS>
struct LightTrace {
    const DevScene &sc; Stats &st; V3 cam_pos; uint32_t skey, dim0;
    VKRT_DEV V3 operator()(uint32_t l, const Surface &sf, const Material &mat) const
    {
        V3 L; float t;
        nee_sample(sc, sf.P, l, skey, dim0, L, t);
        Hit sh{t, 0, 0};
        if (!BVH)     // 10-primitive scenes: the ray is cheaper than the BRDF, keep the shader's order
            return trace_ray<true, BVH, true, STATS>(sc, sf.P, L, sh, st) ? v3(0.0f) : light_term(sc, sf, mat, cam_pos, l, L, t);
        const V3 term = light_term(sc, sf, mat, cam_pos, l, L, t);
        if (term_is_zero(term)) { ++st.shadow; ++st.skipped; return v3(0.0f); }
        return trace_ray<true, BVH, true, STATS>(sc, sf.P, L, sh, st) ? v3(0.0f) : term;
    }
};

// Shades the hit (the light loop asks `lights(l, surface, material)` for every emissive sphere's
// contribution) and rolls Russian roulette.  Returns true when the path continues into the next depth iteration.
// `emissive_used` receives the emissive term that was added next to the light sum (DIFFUSE branch only; the
// fused wavefront needs it to form the accumulator for the other outcome of a pending shadow ray).
template <class Lights>
VKRT_DEV bool path_shade(const DevScene &sc, V3 cam_pos, uint32_t max_depth, uint32_t skey, PathState &ps,
                         const Hit &hit, const Lights &lights, V3 *emissive_used = nullptr)
{
    const uint32_t dim0 = ps.depth * DIMS_PER_BOUNCE;
    const Surface sf = surface_of(sc, ps.o, ps.d, hit);
    const Material mat = load_material(sc, sf.mat);
    if (mat.type == 0u) {                                                            // MAT_TYPE_DIFFUSE
        const float r2 = u01(skey, dim0 + SLOT_R2);
        const V3 dj = jitter(sf.N, VKRT_TWO_PI * u01(skey, dim0 + SLOT_PHI), sqrtf(r2), sqrtf(1.0f - r2)) *
                      (1.0f - mat.metalness);
        V3 e = v3(0.0f);
        for (uint32_t l = 0; l < sc.n_lights; ++l) e = e + lights(l, sf, mat);
        const bool all_pos = mat.emissive.x > 0.0f && mat.emissive.y > 0.0f && mat.emissive.z > 0.0f;
        const V3 emissive = all_pos ? normalize3(mat.emissive) : v3(0.0f);
        if (emissive_used) *emissive_used = emissive;
        ps.acc = ps.acc + ps.mask * (emissive + e);
        ps.mask = ps.mask * mat.albedo;
        const V3 nd = normalize3(reflect3(ps.d, sf.N) + dj);
        ps.o = sf.P; ps.d = nd;
    } else {                                                                         // MAT_TYPE_DIELECTRIC
        ps.acc = ps.acc + mat.emissive * ps.mask;
        ps.mask = ps.mask * mat.albedo;
        const float cosine = -dot3(ps.d, sf.N) / length3(ps.d);
        const V3 reflected = reflect3(ps.d, sf.N);
        const V3 refracted = refract3(ps.d, sf.N, mat.roughness);
        const bool is_zero = refracted.x == 0.0f && refracted.y == 0.0f && refracted.z == 0.0f;
        const float p_reflect = is_zero ? 1.0f : schlick(cosine, mat.roughness);
        const V3 nd = normalize3(u01(skey, dim0 + SLOT_R2) < p_reflect ? reflected : refracted);
        ps.o = sf.P; ps.d = nd;
    }
    const float p = max3(ps.mask);
    if (u01(skey, dim0 + SLOT_RR) > p) return false;
    ps.mask = ps.mask * (1.0f / p);
    ++ps.depth;
    return ps.depth < max_depth;
}

// full bounce = firefly clamp + closest hit + shade.  Returns true while the path is alive.
template <bool BVH, bool STATS>
VKRT_DEV bool path_bounce(const DevScene &sc, V3 cam_pos, uint32_t max_depth, uint32_t skey, PathState &ps,
                          Stats &st, uint32_t *primary_id)
{
    ps.acc = clamp3(ps.acc, 0.0f, 1.0f);                                             // :441
    Hit hit{path_tmax(ps.depth), 0, 0};
    const bool found = trace_ray<true, BVH, false, STATS>(sc, ps.o, ps.d, hit, st);
    if (primary_id) *primary_id = found ? ((hit.kind << 28) | hit.index) : 0u;
    if (!found) return false;
    const LightTrace<BVH, STATS> lights{sc, st, cam_pos, skey, ps.depth * DIMS_PER_BOUNCE};
    return path_shade(sc, cam_pos, max_depth, skey, ps, hit, lights);
}

// ---- primary ray: Tracer.comp:561-574 == Raytracer.comp:361-376 -------------------------------
VKRT_DEV void primary_ray(const vkrt_frame_data &fd, uint32_t w, uint32_t h, uint32_t x, uint32_t y, V3 &o, V3 &d)
{
    const float u = (float)x / (float)w, v = (float)y / (float)h;
    const float tx = 2.0f * u - 1.0f, ty = 2.0f * v - 1.0f;
    const V3 cd = v3(fd.camera.dir.x, fd.camera.dir.y, fd.camera.dir.z);
    const V3 cr = v3(fd.camera.right.x, fd.camera.right.y, fd.camera.right.z);
    const V3 cu = v3(fd.camera.up.x, fd.camera.up.y, fd.camera.up.z);
    V3 dir = cd + cr * tx + cu * ty;
    dir = dir * v3(fd.aspect_ratio, 1.0f, fd.aspect_ratio);
    o = v3(fd.camera.pos.x, fd.camera.pos.y, fd.camera.pos.z);
    d = normalize3(dir);
}

// ---- whitted: render_scene + main, Raytracer.comp:280-399 -------------------------------------
template <bool BVH, bool STATS>
VKRT_DEV V3 whitted_render_scene(const DevScene &sc, V3 &ro, V3 &rd, uint32_t &bounce_depth, uint32_t bounces,
                                 V3 light_pos, V3 cam_pos, Stats &st, uint32_t *primary_id)
{
    V3 color = v3(0.0f);
    Hit hit{1000.0f, 0, 0};
    const bool found = trace_ray<false, BVH, false, STATS>(sc, ro, rd, hit, st);
    if (primary_id) *primary_id = found ? ((hit.kind << 28) | hit.index) : 0u;
    if (!found) return color;
    const Surface sf = surface_of(sc, ro, rd, hit);
    const Material mat = load_material(sc, sf.mat);
    const V3 light_vec = normalize3(light_pos - sf.P);
    const float dist_to_light = length3(light_pos - sf.P);
    {
        const float li = 540.0f / ((4.0f * 3.14159268f) * dist_to_light);
        const V3 light_intensity = v3(li);
        const V3 diffuse = light_intensity * mat.albedo * gl_max(dot3(sf.N, light_vec), 0.0f);
        const V3 half_vec = normalize3(light_vec + normalize3(cam_pos));
        const V3 specular = light_intensity * pow_(gl_clamp(dot3(sf.N, half_vec), 0.0f, 1.0f), 16.0f);
        color = diffuse + specular;
    }
    {
        Hit sh{dist_to_light, 0, 0};
        if (trace_ray<false, BVH, true, STATS>(sc, sf.P, light_vec, sh, st)) {
            color = color * 0.5f;
            bounce_depth = bounces + 1;
        }
    }
    if (mat.metalness >= 0.5f) { rd = reflect3(rd, sf.N); ro = sf.P; }
    else bounce_depth = bounces + 1;
    return color;
}

template <bool BVH, bool STATS>
VKRT_DEV V3 whitted_pixel(const DevScene &sc, const vkrt_frame_data &fd, V3 ro, V3 rd, uint32_t bounces, Stats &st,
                          uint32_t *primary_id)
{
    const V3 light_pos = v3(fd.light_pos.x, fd.light_pos.y, fd.light_pos.z);
    const V3 cam_pos = v3(fd.camera.pos.x, fd.camera.pos.y, fd.camera.pos.z);
    uint32_t bounce = 0;
    V3 fin = whitted_render_scene<BVH, STATS>(sc, ro, rd, bounce, bounces, light_pos, cam_pos, st, primary_id);
    float strength = 0.4f;
    while (++bounce <= bounces) {
        const V3 refl = whitted_render_scene<BVH, STATS>(sc, ro, rd, bounce, bounces, light_pos, cam_pos, st, nullptr);
        fin = (1.0f - strength) * fin + strength * mix3(refl, fin, 1.0f - strength);
        strength *= 0.5f;
    }
    return fin;
}

VKRT_DEV uint8_t unorm8(float x)
{
    if (!(x == x)) return 0;
    const float c = gl_clamp(x, 0.0f, 1.0f);
    return (uint8_t)(int)floorf(c * 255.0f + 0.5f);
}

} // namespace vkrt
