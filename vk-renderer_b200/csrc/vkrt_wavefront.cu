// Wavefront variant of the path integrator (Assets/Tracer.comp::radiance, :433-553).
//
// One "wave" holds S samples of every owned pixel as path records in HBM: one 64-byte record per path =
// two 32-byte sectors {origin.xyz t_hit | dir.xyz hit id} {acc.xyz pixel | mask.xyz sample<<8|depth}, each moved by
// ONE 256-bit load / store.  Records are sample-major (path = sample_in_wave * n_slots + pixel_slot); as the
// depth iterations thin the paths out, a surviving path still costs exactly its own two sectors of DRAM traffic
// (with one float4 array per field every access dragged in half a sector of a dead neighbour).
// Per depth iteration the queues are processed by small kernels:
//
//   extend   persistent warps pull rays from the active queue (one atomicAdd per warp for all lanes
//            that need a ray, found with __ballot_sync, base broadcast with __shfl_sync); a lane whose
//            traversal ends is refilled as soon as fewer than REFILL lanes are still traversing, so the
//            heavy-tailed LBVH traversal keeps the warp full.  Triangles + spheres only.
//   classify firefly clamp (:441), the plane loop of trace_ray (:414-428) on top of extend's result,
//            misses end the path, hits are binned by material type into the dielectric queue and the
//            diffuse queue (warp-aggregated pushes)
//   dielectric  shades the dielectric bin (:514-542) + Russian roulette, pushes survivors
//   nee      light-sample shadow rays of the diffuse bin (:464-469); triangle / plane occluders are
//            resolved here, only rays that still need the sphere any-hit query are queued
//   shadow   the same persistent traversal kernel in any-hit mode (spheres only)
//   diffuse  shades the diffuse bin (:451-513) with the occlusion flags + Russian roulette
//
// and a wave ends with `reduce`, which adds the per-sample radiances of every pixel IN SAMPLE ORDER,
// so the result is bit-identical to the megakernel and to the oracle whatever the scheduling was.
#include <cstdlib>

#include "vkrt_device.cuh"
#include "vkrt_internal.h"

#ifndef VKRT_REFILL
#define VKRT_REFILL 20      // refill the warp when fewer than this many lanes are still traversing
#endif
#ifndef VKRT_FETCH_CHUNK
#define VKRT_FETCH_CHUNK 64   // ray indices a warp reserves per atomicAdd on the queue head
#endif
#ifndef VKRT_NESTED_STEPS
#define VKRT_NESTED_STEPS 1   // the round's inner steps form ONE divergent region (no reconvergence point per step): 26.41 -> 26.01 ms/frame;
#endif                        // steps per round 2 .. 12 measured: 3, 6 and 9 are equal (25.9 - 26.0), the others 0.4 - 1.2 ms slower
#ifndef VKRT_TRAV_UNROLL
#define VKRT_TRAV_UNROLL 6    // inner steps per node-loop round (one nested region, one vote): 3, 6 and 9 are the good values, 6 is 0.5 % ahead of 3
#endif
#ifndef VKRT_LEAF_BATCH
#define VKRT_LEAF_BATCH 8    // run a leaf phase once this many lanes wait at a leaf (0: leaf tests inside the node step)
#endif
#ifndef VKRT_FAST_INNER
#define VKRT_FAST_INNER 20    // lanes at inner nodes from which the node loop skips its other votes (0 = never; 12..25 measured, 19-20 best:
                              // cfg4 26.98 -> 26.64 ms/frame)
#endif
#ifndef VKRT_TRACE_BLOCK
#define VKRT_TRACE_BLOCK 128
#endif
#ifndef VKRT_TRACE_MINBLOCKS
#define VKRT_TRACE_MINBLOCKS 12    // 40 registers, 48 warps/SM: latency hiding beats the few spills (9/10/11/12/14/16 measured)
#endif
#ifndef VKRT_TRACE_RESIDENT
#define VKRT_TRACE_RESIDENT 0       // cap of resident trace blocks per SM (0: as many as fit)
#endif
#ifndef VKRT_LOGIC_MINBLOCKS
#define VKRT_LOGIC_MINBLOCKS 4      // __launch_bounds__(256, n) of the fused classify + shade kernel
#endif
#ifndef VKRT_FUSED
#define VKRT_FUSED 1               // scenes with <= 1 light: logic + mixed trace (2 launches per depth) instead of 4
#endif
#ifndef VKRT_FUSED_GENERATE
#define VKRT_FUSED_GENERATE 1      // fused pipeline: depth 0 is shaded inside generate (one primary hit per pixel)
#endif
#ifndef VKRT_SPEC_LEAF
#define VKRT_SPEC_LEAF 0           // trace: park the first scheduled leaf and keep walking inner nodes
#endif
#ifndef VKRT_COLD_SMEM
#define VKRT_COLD_SMEM (VKRT_LEAF_BATCH && VKRT_QNODES && !VKRT_SPEC_LEAF)   // trace: leaf-test-only ray state in shared memory
#endif
#ifndef VKRT_RAY_LDCG
#define VKRT_RAY_LDCG 1            // trace: ray / shadow records and queue items are read past L1 (ld.global.cg): they are used once,
#endif                             //        L1 is kept for BVH nodes and the traversal stacks
#ifndef VKRT_COLD_FLOATS
#define VKRT_COLD_FLOATS 12        // trace: 12 = origin, direction and the exact slab constants wait in shared memory for the leaf
#endif                             //        tests; 6 = origin and direction only, the slab constants are recomputed there (bit-identical)
#ifndef VKRT_OCTANT_BIN
#define VKRT_OCTANT_BIN 0          // logic: a block's surviving paths enter the next queue grouped by the direction octant of their
#endif                             //        next ray (block-local counting sort; exact, measured slower: 30.9 vs 30.3 ms/frame on cfg4 -- off)
#ifndef VKRT_BLOCK_PUSH
#define VKRT_BLOCK_PUSH 1          // classify / shade: one queue-counter atomicAdd per block instead of per warp
#endif
#ifndef VKRT_DENSE
#define VKRT_DENSE 1               // fused pipeline: survivors of a depth are written DENSELY (ping-pong ray / state arrays, position =
#endif                             //                 queue index): no path-id indirection, every record access is a contiguous stream
#ifndef VKRT_GEN_PER_SAMPLE
#define VKRT_GEN_PER_SAMPLE 0      // 1: dense pipeline, depth 0 by one thread per (pixel, sample) behind a primary-hit kernel (one thread
#endif                             // per pixel): the S samples of a pixel become neighbours in the depth-1 arrays (shared origin).  Exact;
                                   // measured on cfg4: the traversal launches gain 4.7 % (20.70 -> 19.73 ms) from the coherence, but depth 0
                                   // costs 4.2 ms instead of 1.95 (every thread redoes what the per-pixel loop hoists): 29.4 vs 28.2 ms -- off
#ifndef VKRT_STAGGER
#define VKRT_STAGGER 0             // > 0: a wave starts when the wave before it has launched this depth's logic (staggered lanes)
#endif
#ifndef VKRT_GEN_PIXEL_MAJOR
#define VKRT_GEN_PIXEL_MAJOR 1     // dense pipeline, depth 0: 1 = a pixel's surviving samples are adjacent in the depth-1 arrays (the Russian
#endif                             //   roulette outcome of every sample is known ahead of the shading: one hash per sample), 0 = sample-major
                                   //   (measured: cfg4 27.85 -> 26.97 ms/frame, traversal launches alone 20.78 -> 19.84 ms)
#ifndef VKRT_WARP_RESERVE
#define VKRT_WARP_RESERVE 0        // dense pipeline: 1 = a warp (not the block) reserves its survivors' range of the next depth's arrays
#endif
#ifndef VKRT_SHADE_BLOCK
#define VKRT_SHADE_BLOCK 256
#endif
#ifndef VKRT_SHADE_MINBLOCKS
#define VKRT_SHADE_MINBLOCKS 4      // __launch_bounds__(256, n) of the streaming shade/classify kernels
#endif

namespace vkrt {

// one counter set per depth iteration (no reset launches): paths entering the depth, the material bins, the
// shadow-ray queue and the two work-fetch heads; survivors are counted in the NEXT depth's set
enum { C_ACTIVE = 0, C_DIEL = 2, C_DIFF = 3, C_SHADOW = 4, C_HEAD_EXTEND = 5, C_HEAD_SHADOW = 6, C_ZERO = 7, C_N = 8, C_SETS = 257 };
static_assert(C_SETS >= 255 + 2, "one counter set per depth 0..max_depth (max_depth <= 255, checked by vkrt_create / vkrt_set_sampling)");

struct WaveParams {
    float4 *rec, *sh, *term, *rad;   // rec: 4 float4 per path (see the top of the file)
    float4 *shrec;                   // fused pipeline: 4 float4 per path {P.xyz t | L.xyz dead | acc_if_visible.xyz pixel | -}
    uint32_t *q_active[2], *q_diel, *q_diff, *q_shadow;
    uint8_t *occ;
    uint32_t *cnt, *cnt_next;    // counter set of this depth / of the next depth
    uint32_t s0, S;          // first sample of the wave, samples per pixel in the wave
    uint32_t n_slots;        // = RenderParams.n_work
    uint32_t n_lights;
    // dense fused pipeline (VKRT_DENSE): the arrays of the depth being processed (x_*) and of the next depth (n_*).
    //   ray   2 float4 per path  {o.xyz - | d.xyz -}
    //   state 2 float4 per path  {acc.xyz slot | mask.xyz sample<<8|depth}
    //   hit   float2 per path    {t_hit, hit id}            written by trace, read by logic
    //   shr   4 float4 per shadow ray {P.xyz t | L.xyz dst | acc_if_visible.xyz slot | -}; dst = index of the path's state in
    //         the NEXT depth's arrays, or 0x80000000 | index into rad when the path ended at this bounce
    const float4 *x_ray; float4 *x_st, *n_ray, *n_st, *x_shr;
    float2 *x_hit;
    // every sample of a pixel starts from the same primary hit (Tracer.comp:574-581): the frame's first wave stores it
    // per pixel slot (prim_mode 1), the later waves of the frame load it instead of traversing again (prim_mode 2)
    float2 *prim; uint32_t prim_mode;
};

VKRT_DEV void ld256(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p) : "memory");
}
// the same load past L1 (cached in L2 only)
VKRT_DEV void ld256cg(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p) : "memory");
}
VKRT_DEV void st256(float4 *p, float4 a, float4 b)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
VKRT_DEV float4 *rec_ray(const WaveParams &wp, uint32_t path) { return wp.rec + 4 * (size_t)path; }
VKRT_DEV float4 *rec_state(const WaveParams &wp, uint32_t path) { return wp.rec + 4 * (size_t)path + 2; }

VKRT_DEV bool slot_to_pixel_w(const RenderParams &rp, uint32_t w, uint32_t &px, uint32_t &py)
{
    const uint32_t lt = w >> 10, in = w & 1023u;
    const uint32_t gt = rp.tile_rank + lt * rp.tile_count;
    const uint32_t tx = gt % rp.tiles_x, ty = gt / rp.tiles_x;
    const uint32_t blk = in >> 5, l = in & 31u;
    px = tx * TILE + (blk & 3u) * 8u + (l & 7u);
    py = ty * TILE + (blk >> 2) * 4u + (l >> 3);
    return px < rp.width && py < rp.height;
}

// warp-aggregated queue push: one atomicAdd per warp, __ballot_sync/__shfl_sync compaction
VKRT_DEV void push(uint32_t *queue, uint32_t *count, bool want, uint32_t value)
{
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u;
    const unsigned m = __ballot_sync(full, want);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(full, base, leader);
    if (want) queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// block-aggregated queue push: the warps' counts meet in shared memory and ONE atomicAdd per block and
// queue reserves the slots (queue counters are single addresses: with one atomic per warp they become the
// bottleneck of the streaming kernels -- same-address atomics serialise in L2).  All threads of the block
// must call it (two __syncthreads inside).
template <int NQ>
VKRT_DEV void push_block(uint32_t *const (&queue)[NQ], uint32_t *const (&count)[NQ], const bool (&want)[NQ], const uint32_t (&value)[NQ])
{
    __shared__ uint32_t s_cnt[NQ][32];
    __shared__ uint32_t s_base[NQ];
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
    unsigned m[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        m[q] = __ballot_sync(full, want[q]);
        if (lane == 0) s_cnt[q][warp] = (uint32_t)__popc(m[q]);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        uint32_t tot = 0;
        for (unsigned w = 0; w < n_warps; ++w) { const uint32_t c = s_cnt[threadIdx.x][w]; s_cnt[threadIdx.x][w] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomicAdd(count[threadIdx.x], tot) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q)
        if (want[q]) queue[q][s_base[q] + s_cnt[q][warp] + (uint32_t)__popc(m[q] & ((1u << lane) - 1u))] = value[q];
    __syncthreads();      // s_cnt / s_base are reused by the next call
}

// block-aggregated push of ONE queue with the block's items grouped by a small key (`bin` < NB): the block reserves
// its range with one atomicAdd like push_block and lays the items out bin by bin, warps in order inside a bin.  Only
// the order of the queue changes (every path's result is independent of it; `reduce` adds in sample order), the warps
// of the next trace launch then hold rays that agree in the key.  All threads of the block must call it.
template <int NB>
VKRT_DEV void push_block_binned(uint32_t *queue, uint32_t *count, bool want, uint32_t value, uint32_t bin)
{
    __shared__ uint32_t s_hist[NB][32];      // [bin][warp]: count, then the exclusive prefix over the warps of the bin
    __shared__ uint32_t s_bin[NB + 1];       // [bin]: total, then the exclusive prefix over the bins; [NB]: the block's queue base
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
    uint32_t rank = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const unsigned m = __ballot_sync(full, want && bin == (uint32_t)b);
        if (lane == 0) s_hist[b][warp] = (uint32_t)__popc(m);
        if (want && bin == (uint32_t)b) rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < NB) {
        uint32_t tot = 0;
        for (unsigned w = 0; w < n_warps; ++w) { const uint32_t c = s_hist[threadIdx.x][w]; s_hist[threadIdx.x][w] = tot; tot += c; }
        s_bin[threadIdx.x] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int b = 0; b < NB; ++b) { const uint32_t c = s_bin[b]; s_bin[b] = tot; tot += c; }
        s_bin[NB] = tot ? atomicAdd(count, tot) : 0u;
    }
    __syncthreads();
    if (want) queue[s_bin[NB] + s_bin[bin] + s_hist[bin][warp] + rank] = value;
    __syncthreads();      // the shared arrays are reused by the next call
}
// direction octant of a ray: the three sign bits (what selects the near planes and the near child in the node loop)
VKRT_DEV uint32_t octant_of(V3 d)
{
    return (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
}

VKRT_DEV void wf_flush(const Stats &st, unsigned long long *counters, bool stats)
{
    const unsigned full = 0xffffffffu;
    const uint32_t c = __reduce_add_sync(full, st.closest), s = __reduce_add_sync(full, st.shadow),
                   p = __reduce_add_sync(full, st.paths), k = __reduce_add_sync(full, st.skipped),
                   h = __reduce_add_sync(full, st.shared);
    uint32_t n = 0, l = 0;
    if (stats) { n = __reduce_add_sync(full, st.nodes); l = __reduce_add_sync(full, st.leaves); }
    if ((threadIdx.x & 31) == 0) {
        if (c) atomicAdd(counters + CNT_CLOSEST, (unsigned long long)c);
        if (s) atomicAdd(counters + CNT_SHADOW, (unsigned long long)s);
        if (p) atomicAdd(counters + CNT_PATHS, (unsigned long long)p);
        if (k) atomicAdd(counters + CNT_SKIPPED, (unsigned long long)k);
        if (h) atomicAdd(counters + CNT_SHARED, (unsigned long long)h);
        if (stats && n) atomicAdd(counters + CNT_NODES, (unsigned long long)n);
        if (stats && l) atomicAdd(counters + CNT_LEAVES, (unsigned long long)l);
    }
}

// ---- generate: Tracer.comp:561-581 ------------------------------------------------------------------
// Every sample of a pixel reuses ONE primary ray (:574-581, no sub-pixel jitter), so its triangle + sphere
// query is done once per pixel here and the result is copied into the S path records of the pixel; the
// ray counter still advances by S (it counts the reference algorithm's trace_ray invocations).
template <bool BVH, bool STATS>
__global__ void __launch_bounds__(256) k_wf_generate(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                      const __grid_constant__ WaveParams wp)
{
    Stats st; stats_zero(st);
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px = 0, py = 0;
    const bool valid = slot < wp.n_slots && slot_to_pixel_w(rp, slot, px, py);
    float4 fo = make_float4(0.f, 0.f, 0.f, 0.f), fd = fo;
    uint32_t pix = 0;
    if (valid) {
        V3 o, d;
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        Hit hit{0.f, 0, 0};
        float cur = path_tmax(0);
        bool found = trace_tris<true>(sc, o, d, cur, hit);
        if (BVH) {
            const SBest b = bvh_query<false, STATS>(sc, o, d, 1e-3f, sphere_bound<true>(cur), st);
            if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
        } else {
            for (uint32_t i = 0; i < sc.n_spheres; ++i) {                         // literal loop, Tracer.comp:398-412
                const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
                if ((t > 1e-3f) && (t < cur + 1e-3f)) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
            }
        }
        fo = make_float4(o.x, o.y, o.z, cur);
        fd = make_float4(d.x, d.y, d.z, __uint_as_float(found ? ((hit.kind << 28) | hit.index) : 0u));
        pix = py * rp.width + px;
        st.closest += wp.S;
        st.shared += wp.S - 1u;
        st.paths += wp.S;
    }
    for (uint32_t sl = 0; sl < wp.S; ++sl) {
        const uint32_t p = sl * wp.n_slots + slot;
        if (valid) {
            st256(rec_ray(wp, p), fo, fd);
            st256(rec_state(wp, p), make_float4(0.f, 0.f, 0.f, __uint_as_float(pix)), make_float4(1.f, 1.f, 1.f, __uint_as_float(sl << 8)));
        }
        uint32_t *const qs[1] = {wp.q_active[0]}; uint32_t *const cs[1] = {wp.cnt + C_ACTIVE};
        const bool ws[1] = {valid}; const uint32_t vs[1] = {p};
        push_block<1>(qs, cs, ws, vs);
    }
    wf_flush(st, rp.counters, STATS);
}

// ---- persistent trace kernel (nearest-hit for `extend`, any-hit for `shadow`) --------------------
// nearest-hit item: path queue[i]; ray = the record's (origin, dir), bound 3000/(depth+1)^2; triangles, then spheres;
//                   result -> the record's t_hit / hit id; the plane loop follows in classify / logic
// any-hit item, MODE 1 (four-kernel pipeline): queue[i] = path * 16 + light; ray (P, sh.xyz), bound sh.w; spheres only; -> occ
// any-hit item, MODE 2 (fused pipeline): queue_sh[i] = path; ray and bound from shrec; an UNOCCLUDED ray stores the
//                   accumulator `logic` prepared for that outcome into the path record (or the path's final radiance)
// MODE 0: nearest-hit items only.  MODE 2: ONE launch traces the nearest-hit rays of depth d and the shadow rays of
// depth d-1 -- item i < n_ext is a nearest-hit item, the rest are shadow items (the longer rays go first); lanes of
// one warp may hold either kind, the traversal is the same code.
// MODE 3 (dense fused pipeline): like MODE 2, but item i IS ray i of the depth's dense ray array and shadow item k is
//                   shadow record k -- no queues; the nearest hit goes to hit[i]
#if VKRT_QNODES
template <int N, bool STATS, class Stack>
VKRT_DEV void nested_inner_steps(Trav &tv, const QRay &qr, Stack &stack, const DevScene &sc, Stats &st)
{
    if (tv.node >= 0) {
        trav_inner_step_q<STATS>(tv, qr, stack, sc, st);
        if (N > 1) nested_inner_steps<(N > 1 ? N - 1 : 1), STATS>(tv, qr, stack, sc, st);
    }
}
#endif
enum { TRACE_EXTEND = 0, TRACE_SHADOW = 1, TRACE_MIXED = 2, TRACE_DENSE = 3 };
template <int MODE, bool BVH, bool STATS, bool TB = true>
__global__ void __launch_bounds__(VKRT_TRACE_BLOCK, VKRT_TRACE_MINBLOCKS) k_wf_trace(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                                const __grid_constant__ WaveParams wp, const uint32_t *__restrict__ queue,
                                                                const uint32_t *__restrict__ n_items_ptr, const uint32_t *__restrict__ queue_sh,
                                                                const uint32_t *__restrict__ n_sh_ptr, uint32_t *head, uint32_t depth)
{
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u;
    constexpr bool MIX = MODE == TRACE_MIXED || MODE == TRACE_DENSE, DENSE = MODE == TRACE_DENSE;
    Stats st; stats_zero(st);
    const uint32_t n_ext = MODE == TRACE_SHADOW ? 0u : *n_items_ptr;
    const uint32_t n_items = MODE == TRACE_SHADOW ? *n_items_ptr : (MIX ? n_ext + *n_sh_ptr : n_ext);
    bool any = MODE == TRACE_SHADOW;
    static_assert(!MIX || VKRT_LEAF_BATCH != 0, "the mixed trace kernel needs the phase-split traversal");
    const float EPS = 1e-3f;
    const float tmax = path_tmax(depth);       // every ray of one extend launch is at the same depth (:444)

    // items a warp reserves per atomicAdd on the queue head: VKRT_FETCH_CHUNK when there is plenty of work, smaller when
    // the launch has fewer items than that per warp -- a thin launch (deep bounces, small tile shards) is then spread over
    // all SMs in many partly filled warps instead of a few full ones, and ends after about one ray's latency
    uint32_t chunk = (uint32_t)VKRT_FETCH_CHUNK;
    {
        const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
        while (chunk > 8u && n_items < chunk * n_warps) chunk >>= 1;
    }
    const uint32_t refill_at = chunk < (uint32_t)VKRT_REFILL ? chunk : (uint32_t)VKRT_REFILL;
    bool has = false, drained = false;
    uint32_t res_base = 0, res_left = 0;       // this warp's current reservation of queue items (warp-uniform)
    uint32_t path = 0, light = 0;
    V3 o = v3(0.f), d = v3(0.f);
    Hit hit{0.f, 0, 0};
    bool found = false;
    float cur = 0.f;
#if VKRT_LEAF_BATCH
    const int FIN = TRAV_DONE;
    Trav tv; tv.node = FIN; tv.sp = 0;
#if VKRT_SMEM_STACK
    __shared__ int s_stack[VKRT_SMEM_STACK][VKRT_TRACE_BLOCK];
#endif
    TravStack<VKRT_TRACE_BLOCK> stack;
#if VKRT_QNODES
    QRay qr;
#endif
#if VKRT_SMEM_STACK
    stack.sm = &s_stack[0][threadIdx.x];
#else
    stack.sm = nullptr;
#endif
#else
    const int FIN = -1;
    Trav tv; tv.node = FIN; tv.sp = 0;
    int stack[BVH_STACK];
#endif
#if VKRT_COLD_SMEM
    __shared__ float s_cold[VKRT_COLD_FLOATS][VKRT_TRACE_BLOCK];
#endif
    int pend = FIN;                            // VKRT_SPEC_LEAF: the parked leaf (~sphere); FIN = none

    for (;;) {
        // ---- refill the lanes that have no ray -------------------------------------------------
        const bool need = !has && !drained;
        const unsigned m = __ballot_sync(full, need);
        if (m) {
            // the warp reserves ray indices VKRT_FETCH_CHUNK at a time (one atomicAdd by one lane, broadcast
            // with __shfl_sync) and hands them to the lanes that need work from that reservation
            // (a thin launch hands out at most one chunk per round: the other idle lanes stay idle for other warps' sake)
            const uint32_t idle = (uint32_t)__popc(m);
            const uint32_t cnt = idle < chunk ? idle : chunk;
            uint32_t new_base = 0;
            if (res_left < cnt) {
                const int leader = __ffs(m) - 1;
                if ((int)lane == leader) new_base = atomicAdd(head, chunk);
                new_base = __shfl_sync(full, new_base, leader);
            }
            const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
            const uint32_t my_item = rank < res_left ? res_base + rank : new_base + (rank - res_left);
            if (res_left < cnt) { res_base = new_base + (cnt - res_left); res_left = chunk - (cnt - res_left); }
            else { res_base += cnt; res_left -= cnt; }
            if (need && rank < cnt) {
                const uint32_t item = my_item;
                if (item >= n_items) drained = true;
                else {
                    has = true;
                    if (MIX) any = item >= n_ext;
                    float4 fo, fd;
                    if (DENSE) {
                        path = any ? item - n_ext : item;              // the item's own index: there is no queue
                        ld256cg(any ? wp.x_shr + 4 * (size_t)path : wp.x_ray + 2 * (size_t)path, fo, fd);
                    } else {
#if VKRT_RAY_LDCG
                        const uint32_t q = __ldcg((MIX && any) ? queue_sh + (item - n_ext) : queue + item);
#else
                        const uint32_t q = (MIX && any) ? queue_sh[item - n_ext] : queue[item];
#endif
                        path = MODE == TRACE_SHADOW ? (q >> 4) : q;
                        light = MODE == TRACE_SHADOW ? (q & 15u) : 0u;
#if VKRT_RAY_LDCG
                        ld256cg(MIX && any ? wp.shrec + 4 * (size_t)path : rec_ray(wp, path), fo, fd);
#else
                        ld256(MIX && any ? wp.shrec + 4 * (size_t)path : rec_ray(wp, path), fo, fd);
#endif
                    }
                    found = false; hit.kind = 0; hit.index = 0;
                    if (MODE == TRACE_SHADOW) {
                        const float4 s = wp.sh[(size_t)path * wp.n_lights + light];
                        o = madd3(fo.w, xyz(fd), xyz(fo));           // the hit point P (== surface_of's P)
                        d = xyz(s); cur = s.w;
                    } else if (MIX && any) {
                        o = xyz(fo); d = xyz(fd); cur = fo.w;        // shrec: {P.xyz t | L.xyz dead}   (dense: {P.xyz t | L.xyz dst})
                        light = __float_as_uint(fd.w);               // "the path ended at this bounce"  (dense: where acc_if_visible goes)
                    } else {
                        o = xyz(fo); d = xyz(fd);
                        cur = tmax;
                        ++st.closest;
                        found = trace_tris<true, false, TB && BVH>(sc, o, d, cur, hit);
                    }
                    if (BVH) {
                        trav_init(tv, sc, o, d, EPS, sphere_bound<true>(cur)); if (tv.node < 0) tv.node = FIN;
#if VKRT_LEAF_BATCH && VKRT_SENTINEL && !VKRT_SMEM_STACK
                        stack.lm[0] = TRAV_DONE; tv.sp = 1;          // sentinel: popping it ends the traversal
#endif
#if VKRT_LEAF_BATCH && VKRT_QNODES
                        qr = qray_setup(sc, tv.sr);
#endif
#if VKRT_COLD_SMEM
                        {
                            float *cs = &s_cold[0][threadIdx.x];
                            cs[0] = o.x; cs[VKRT_TRACE_BLOCK] = o.y; cs[2 * VKRT_TRACE_BLOCK] = o.z;
                            cs[3 * VKRT_TRACE_BLOCK] = d.x; cs[4 * VKRT_TRACE_BLOCK] = d.y; cs[5 * VKRT_TRACE_BLOCK] = d.z;
#if VKRT_COLD_FLOATS >= 9
                            cs[6 * VKRT_TRACE_BLOCK] = tv.sr.inv.x; cs[7 * VKRT_TRACE_BLOCK] = tv.sr.inv.y; cs[8 * VKRT_TRACE_BLOCK] = tv.sr.inv.z;
#endif
#if VKRT_COLD_FLOATS >= 12
                            cs[9 * VKRT_TRACE_BLOCK] = tv.sr.oinv.x; cs[10 * VKRT_TRACE_BLOCK] = tv.sr.oinv.y; cs[11 * VKRT_TRACE_BLOCK] = tv.sr.oinv.z;
#endif
                        }
#endif
                    }
                    else tv.node = FIN;
                }
            }
        }
        if (__all_sync(full, !has)) break;

        // ---- traverse; leave as soon as the warp has thinned out and there are rays left to fetch --
        if (BVH) {
#if VKRT_LEAF_BATCH && VKRT_SPEC_LEAF
            // a lane that reaches a scheduled leaf parks it in `pend` and keeps walking inner nodes (the leaf test is
            // only postponed: any order of box and leaf tests gives rule S's answer, a later test merely culls later);
            // it waits only when a second leaf arrives while the first is still parked.  Leaf tests run when
            // VKRT_LEAF_BATCH lanes have one parked, or nobody has an inner node left.
            for (;;) {
                const bool trav = has && (tv.node != FIN || pend != FIN);
                const unsigned tm = __ballot_sync(full, trav);
                if (tm == 0) break;
                if (__popc(tm) < (int)refill_at && __any_sync(full, !drained && !trav)) break;
                const unsigned im = __ballot_sync(full, trav && tv.node >= 0);
                const unsigned pm = __ballot_sync(full, trav && pend != FIN);
                if (im == 0 || __popc(pm) >= VKRT_LEAF_BATCH) {
                    if (trav && pend != FIN) {
                        leaf_test<STATS>(tv, sc, o, d, ~pend, st);
                        pend = FIN;
                        if (any && tv.best.idx >= 0) tv.node = FIN;                  // any-hit: answered
                    }
                    if (trav && tv.node < 0 && tv.node != FIN) { pend = tv.node; tv.node = tv.sp ? stack.pop(tv.sp) : FIN; }
                } else {
#pragma unroll
                    for (int u = 0; u < VKRT_TRAV_UNROLL; ++u)
                        if (has && tv.node >= 0) {
#if VKRT_QNODES
                            trav_inner_step_q<STATS>(tv, qr, stack, sc, st);
#else
                            trav_inner_step<STATS>(tv, stack, sc, st);
#endif
                            if (tv.node < 0 && tv.node != FIN && pend == FIN) { pend = tv.node; tv.node = tv.sp ? stack.pop(tv.sp) : FIN; }
                        }
                }
            }
#elif VKRT_LEAF_BATCH
            for (;;) {
                // lanes at inner nodes (a lane without a ray holds FIN, a lane at a scheduled leaf ~sphere: both negative)
                const unsigned im = __ballot_sync(full, tv.node >= 0);
                const int n_inner = __popc(im);
                bool leaf_phase = false;
                // With VKRT_FAST_INNER or more lanes at inner nodes the warp just keeps walking: one vote per round instead of
                // four (with >= 25 the decisions below could not come out differently anyway; 20 measured best)
                if (n_inner < (VKRT_FAST_INNER ? VKRT_FAST_INNER : 33)) {
                    const bool trav = has && tv.node != FIN;
                    const unsigned tm = __ballot_sync(full, trav);
                    if (tm == 0) break;
                    if (__popc(tm) < (int)refill_at && __any_sync(full, !drained && !trav)) break;
                    // lanes whose next item is a scheduled leaf wait until VKRT_LEAF_BATCH of them can run the leaf test
                    // together (or nobody has an inner node left); everybody else keeps visiting inner nodes
                    leaf_phase = n_inner == 0 || __popc(tm) - n_inner >= VKRT_LEAF_BATCH;
                }
                if (leaf_phase) {
                    if (has && tv.node < 0 && tv.node != FIN) {
#if VKRT_COLD_SMEM
                        // the ray's origin, direction and exact slab constants are only needed here: they live in shared
                        // memory ([field][thread], conflict-free) so the node loop's registers hold nothing cold
                        const float *cs = &s_cold[0][threadIdx.x];
                        const V3 co = v3(cs[0], cs[VKRT_TRACE_BLOCK], cs[2 * VKRT_TRACE_BLOCK]);
                        const V3 cd = v3(cs[3 * VKRT_TRACE_BLOCK], cs[4 * VKRT_TRACE_BLOCK], cs[5 * VKRT_TRACE_BLOCK]);
#if VKRT_COLD_FLOATS >= 12
                        tv.sr.inv = v3(cs[6 * VKRT_TRACE_BLOCK], cs[7 * VKRT_TRACE_BLOCK], cs[8 * VKRT_TRACE_BLOCK]);
                        tv.sr.oinv = v3(cs[9 * VKRT_TRACE_BLOCK], cs[10 * VKRT_TRACE_BLOCK], cs[11 * VKRT_TRACE_BLOCK]);
#elif VKRT_COLD_FLOATS >= 9
                        tv.sr.inv = v3(cs[6 * VKRT_TRACE_BLOCK], cs[7 * VKRT_TRACE_BLOCK], cs[8 * VKRT_TRACE_BLOCK]);
                        tv.sr.oinv = co * tv.sr.inv;     // the same three multiplications as slab_setup
#else
                        tv.sr = slab_setup(co, cd);      // the same operations on the same values as at the ray's start
#endif
                        trav_leaf_step<STATS>(tv, stack, sc, co, cd, any, st);
#else
                        trav_leaf_step<STATS>(tv, stack, sc, o, d, any, st);
#endif
                    }
                } else {
#if VKRT_NESTED_STEPS && VKRT_QNODES
                    // ONE divergent region for the round's steps: a lane that runs out of inner nodes leaves it for good
                    // (no reconvergence point per step)
                    nested_inner_steps<VKRT_TRAV_UNROLL, STATS>(tv, qr, stack, sc, st);
#else
#pragma unroll
                    for (int u = 0; u < VKRT_TRAV_UNROLL; ++u)     // the warp votes above cost ~10 instructions: amortise them
#if VKRT_QNODES
                        if (tv.node >= 0) trav_inner_step_q<STATS>(tv, qr, stack, sc, st);
#else
                        if (tv.node >= 0) trav_inner_step<STATS>(tv, stack, sc, st);
#endif
#endif
                }
            }
#else
            for (;;) {
                const bool trav = has && tv.node >= 0;
                const unsigned tm = __ballot_sync(full, trav);
                if (tm == 0) break;
                if (__popc(tm) < (int)refill_at && __any_sync(full, !drained && !trav)) break;
#pragma unroll
                for (int u = 0; u < VKRT_TRAV_UNROLL; ++u)     // the warp votes above cost ~10 instructions: amortise them
                    if (has && tv.node >= 0) trav_step<MODE == TRACE_SHADOW, STATS>(tv, stack, sc, o, d, st);
            }
#endif
        }

        // ---- finish the lanes whose traversal is over ---------------------------------------------
        if (has && tv.node == FIN && pend == FIN) {
            if (BVH) {
                if (tv.best.idx >= 0) { cur = tv.best.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)tv.best.idx; found = true; }
            } else {
                for (uint32_t i = 0; i < sc.n_spheres; ++i) {                     // literal loop, Tracer.comp:398-412
                    const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
                    if ((t > EPS) && (t < cur + EPS)) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
                }
            }
            if (MODE == TRACE_SHADOW) wp.occ[(size_t)path * wp.n_lights + light] = found ? 1 : 0;
            else if (DENSE && any) {
                if (!found) {        // unoccluded: the light's term counts (Tracer.comp:473-503)
                    const float4 a1 = __ldcg(wp.x_shr + 4 * (size_t)path + 2);
                    if (light & 0x80000000u) __stcg(wp.rad + (light & 0x7fffffffu), make_float4(a1.x, a1.y, a1.z, 0.f));
                    else __stcg(wp.x_st + 2 * (size_t)light, a1);
                }
            } else if (DENSE) {
                __stcg(wp.x_hit + path, make_float2(cur, __uint_as_float(found ? ((hit.kind << 28) | hit.index) : 0u)));
            } else if (MIX && any) {
                if (!found) {        // unoccluded: the light's term counts (Tracer.comp:473-503)
#if VKRT_RAY_LDCG
                    const float4 a1 = __ldcg(wp.shrec + 4 * (size_t)path + 2);
#else
                    const float4 a1 = wp.shrec[4 * (size_t)path + 2];
#endif
#if VKRT_RAY_LDCG
                    if (light) __stcg(wp.rad + path, make_float4(a1.x, a1.y, a1.z, 0.f));
                    else __stcg(rec_state(wp, path), a1);
#else
                    if (light) wp.rad[path] = make_float4(a1.x, a1.y, a1.z, 0.f);
                    else rec_state(wp, path)[0] = a1;
#endif
                }
            } else {
#if VKRT_RAY_LDCG
                __stcg(&rec_ray(wp, path)[0].w, cur);
                __stcg(&rec_ray(wp, path)[1].w, __uint_as_float(found ? ((hit.kind << 28) | hit.index) : 0u));
#else
                rec_ray(wp, path)[0].w = cur;
                rec_ray(wp, path)[1].w = __uint_as_float(found ? ((hit.kind << 28) | hit.index) : 0u);
#endif
            }
            has = false;
        }
    }
    wf_flush(st, rp.counters, STATS);
}

VKRT_DEV void load_path(const WaveParams &wp, uint32_t path, PathState &ps, Hit &hit, uint32_t &pix, uint32_t &sl)
{
    float4 fo, fd, fa, fm;
    ld256(rec_ray(wp, path), fo, fd);
    ld256(rec_state(wp, path), fa, fm);
    ps.o = xyz(fo); ps.d = xyz(fd); ps.acc = xyz(fa); ps.mask = xyz(fm);
    const uint32_t sd = __float_as_uint(fm.w);
    ps.depth = sd & 255u; sl = sd >> 8;
    pix = __float_as_uint(fa.w);
    const uint32_t id = __float_as_uint(fd.w);
    hit.t = fo.w; hit.kind = id >> 28; hit.index = id & 0x0fffffffu;
}
VKRT_DEV void store_path(const WaveParams &wp, uint32_t path, const PathState &ps, uint32_t pix, uint32_t sl)
{
    st256(rec_ray(wp, path), make_float4(ps.o.x, ps.o.y, ps.o.z, 0.f), make_float4(ps.d.x, ps.d.y, ps.d.z, 0.f));
    st256(rec_state(wp, path), make_float4(ps.acc.x, ps.acc.y, ps.acc.z, __uint_as_float(pix)),
          make_float4(ps.mask.x, ps.mask.y, ps.mask.z, __uint_as_float((sl << 8) | ps.depth)));
}

struct LightsNone {      // the dielectric branch never evaluates a light
    VKRT_DEV V3 operator()(uint32_t, const Surface &, const Material &) const { return v3(0.0f); }
};

// ---- classify: one pass over the active paths after `extend` ----------------------------------------
//   firefly clamp (:441); the plane loop of trace_ray (:414-428) on top of extend's triangle/sphere result;
//   miss -> the path ends (:445); DIELECTRIC hit -> shaded here completely (:514-549), survivors are pushed to
//   the next active queue; DIFFUSE hit -> the light-sample shadow rays (:464-469) are generated here, their
//   triangle/plane occluders resolved, the rest queued for the sphere any-hit kernel, and the path goes to
//   the diffuse bin.  (Binning by material type = which queue a path is pushed to.)
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_SHADE_MINBLOCKS) k_wf_classify(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                         const __grid_constant__ WaveParams wp, const uint32_t *__restrict__ queue,
                                                         const uint32_t *__restrict__ n_ptr)
{
    Stats st; stats_zero(st);
    const uint32_t n = *n_ptr;
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool diff = false, diel = false;
        uint32_t path = 0, skey = 0, dim0 = 0;
        Surface sf; sf.P = v3(0.f); sf.N = v3(0.f); sf.mat = 0;
        Material dmat{};
        if (i < n) {
            path = queue[i];
            PathState ps; Hit hit; uint32_t pix, sl;
            load_path(wp, path, ps, hit, pix, sl);
            const V3 acc0 = ps.acc;
            ps.acc = clamp3(ps.acc, 0.0f, 1.0f);                                                        // :441
            const uint32_t id0 = (hit.kind << 28) | hit.index;
            float cur = hit.t;
            const bool found = trace_planes<true>(sc, ps.o, ps.d, cur, hit) || id0 != 0u;             // :414-428
            hit.t = cur;
            const uint32_t id = found ? ((hit.kind << 28) | hit.index) : 0u;
            if (rp.hit_ids && ps.depth == 0u && sl == 0u && wp.s0 == rp.s_begin) rp.hit_ids[pix] = id;
            if (!found) wp.rad[path] = make_float4(ps.acc.x, ps.acc.y, ps.acc.z, 0.f);               // miss: break (:445)
            else {
                const uint32_t mat = hit.kind == KIND_TRI ? sc.tri_mat : hit.kind == KIND_SPHERE ? __ldg(sc.sphere_mat + hit.index) : sc.plane_mat[hit.index];
                const uint32_t type = __float_as_uint(__ldg(&sc.mats[3 * mat + 2].x));
                skey = sample_key(rp.fkey, pix, wp.s0 + sl);
                // the record keeps the final hit and the clamped accumulator for the shade kernels
                if (id != id0) { rec_ray(wp, path)[0].w = cur; rec_ray(wp, path)[1].w = __uint_as_float(id); }
                if (__float_as_uint(acc0.x) != __float_as_uint(ps.acc.x) || __float_as_uint(acc0.y) != __float_as_uint(ps.acc.y) ||
                    __float_as_uint(acc0.z) != __float_as_uint(ps.acc.z))
                    rec_state(wp, path)[0] = make_float4(ps.acc.x, ps.acc.y, ps.acc.z, __uint_as_float(pix));
                if (type != 0u) diel = true;                                                           // DIELECTRIC bin
                else {                                                                                 // DIFFUSE bin
                    diff = true;
                    dim0 = ps.depth * DIMS_PER_BOUNCE;
                    sf = surface_of(sc, ps.o, ps.d, hit);
                    dmat = load_material(sc, sf.mat);
                }
            }
        }
#if VKRT_BLOCK_PUSH
        {
            uint32_t *const qs[2] = {wp.q_diel, wp.q_diff}; uint32_t *const cs[2] = {wp.cnt + C_DIEL, wp.cnt + C_DIFF};
            const bool ws[2] = {diel, diff}; const uint32_t vs[2] = {path, path};
            push_block<2>(qs, cs, ws, vs);
        }
#else
        push(wp.q_diel, wp.cnt + C_DIEL, diel, path);
        push(wp.q_diff, wp.cnt + C_DIFF, diff, path);
#endif
        for (uint32_t l = 0; l < sc.n_lights; ++l) {
            bool queue_it = false;
            if (diff) {
                V3 L; float t;
                nee_sample(sc, sf.P, l, skey, dim0, L, t);
                ++st.shadow;
                // the unoccluded contribution is evaluated first: an exactly-zero term (N.L <= 0) needs no ray
                const V3 term = light_term(sc, sf, dmat, cam_pos, l, L, t);
                bool occluded = term_is_zero(term);
                if (occluded) ++st.skipped;
                // any-hit is a plain OR over the primitive classes: triangles and planes here,
                // spheres (bound t + EPSILON, unchanged because nothing was accepted before them) in `shadow`
                Hit h{t, 0, 0};
                float cur = t;
                if (!occluded) occluded = trace_tris<true>(sc, sf.P, L, cur, h);
                if (!occluded) { cur = t; occluded = trace_planes<true>(sc, sf.P, L, cur, h); }
                wp.occ[(size_t)path * sc.n_lights + l] = occluded ? 1 : 0;     // `shadow` overwrites the 0 of queued rays
                if (!occluded) {
                    wp.sh[(size_t)path * sc.n_lights + l] = make_float4(L.x, L.y, L.z, t);
                    wp.term[(size_t)path * sc.n_lights + l] = make_float4(term.x, term.y, term.z, 0.f);
                }
                queue_it = !occluded;
            }
#if VKRT_BLOCK_PUSH
            {
                uint32_t *const qs[1] = {wp.q_shadow}; uint32_t *const cs[1] = {wp.cnt + C_SHADOW};
                const bool ws[1] = {queue_it}; const uint32_t vs[1] = {(path << 4) | l};
                push_block<1>(qs, cs, ws, vs);
            }
#else
            push(wp.q_shadow, wp.cnt + C_SHADOW, queue_it, (path << 4) | l);
#endif
        }
    }
    wf_flush(st, rp.counters, false);
}

struct LightsStored {    // the terms were evaluated by classify, the occlusion flags by classify / shadow
    const uint8_t *occ; const float4 *term;
    VKRT_DEV V3 operator()(uint32_t l, const Surface &, const Material &) const { return occ[l] != 0 ? v3(0.0f) : xyz(term[l]); }
};
// ---- shade both material bins + Russian roulette (:545-549): the DIELECTRIC bin (:514-542) needs no light, the
// DIFFUSE bin (:451-513) uses the stored light terms and occlusion flags; survivors go to the next depth's queue
template <bool DIFFUSE>
VKRT_DEV void shade_bin(const DevScene &sc, const RenderParams &rp, const WaveParams &wp, uint32_t next, V3 cam_pos)
{
    const uint32_t *queue = DIFFUSE ? wp.q_diff : wp.q_diel;
    const uint32_t n = wp.cnt[DIFFUSE ? C_DIFF : C_DIEL];
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool alive = false;
        uint32_t path = 0;
        if (i < n) {
            path = queue[i];
            PathState ps; Hit hit; uint32_t pix, sl;
            load_path(wp, path, ps, hit, pix, sl);
            const uint32_t skey = sample_key(rp.fkey, pix, wp.s0 + sl);
            if (DIFFUSE) {
                const LightsStored lights{wp.occ + (size_t)path * sc.n_lights, wp.term + (size_t)path * sc.n_lights};
                alive = path_shade(sc, cam_pos, rp.max_depth, skey, ps, hit, lights);
            } else {
                alive = path_shade(sc, cam_pos, rp.max_depth, skey, ps, hit, LightsNone{});
            }
            if (alive) store_path(wp, path, ps, pix, sl);
            else wp.rad[path] = make_float4(ps.acc.x, ps.acc.y, ps.acc.z, 0.f);
        }
#if VKRT_BLOCK_PUSH
        {
            uint32_t *const qs[1] = {wp.q_active[next]}; uint32_t *const cs[1] = {wp.cnt_next + C_ACTIVE};
            const bool ws[1] = {alive}; const uint32_t vs[1] = {path};
            push_block<1>(qs, cs, ws, vs);
        }
#else
        push(wp.q_active[next], wp.cnt_next + C_ACTIVE, alive, path);
#endif
    }
}
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_SHADE_MINBLOCKS) k_wf_shade(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                      const __grid_constant__ WaveParams wp, uint32_t next)
{
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    shade_bin<false>(sc, rp, wp, next, cam_pos);
    shade_bin<true>(sc, rp, wp, next, cam_pos);
}

// ---- fused pipeline (scenes with at most one emissive sphere): `logic` = classify + shade in one pass ------------
// With one light the light sum of a DIFFUSE hit is  e = 0 + (occluded ? 0 : term)  (Tracer.comp:457-503), so both
// possible accumulators are known before the shadow ray is traced:
//     acc_occluded = acc + mask * (emissive + (0 + 0))          acc_visible = acc + mask * (emissive + (0 + term))
// `logic` shades the path completely with acc_occluded (next ray, mask, Russian roulette, the record or the final
// radiance) and leaves {P, t, L, "path ended", acc_visible} for the shadow ray; the trace launch of the NEXT depth
// (TRACE_MIXED) answers it next to that depth's nearest-hit rays and, if the ray is unoccluded, stores acc_visible.
// The firefly clamp of the next iteration (:441) reads the accumulator after that launch, so every value is formed
// by the same operations in the same order as in radiance(); per depth the wave costs two launches instead of four
// and one pass over the path records instead of two.
template <bool TB>
struct LightsDeferred {
    const DevScene &sc; V3 cam_pos; uint32_t skey, dim0; Stats &st;
    bool *need_ray; V3 *L, *term; float *t;
    VKRT_DEV V3 operator()(uint32_t l, const Surface &sf, const Material &mat) const
    {
        V3 Ld; float td;
        nee_sample(sc, sf.P, l, skey, dim0, Ld, td);
        ++st.shadow;
        const V3 tm = light_term(sc, sf, mat, cam_pos, l, Ld, td);
        bool occluded = term_is_zero(tm);
        if (occluded) ++st.skipped;
        Hit h{td, 0, 0};
        float cur = td;
        if (!occluded) occluded = trace_tris<true, true, TB>(sc, sf.P, Ld, cur, h);
        if (!occluded) { cur = td; occluded = trace_planes<true>(sc, sf.P, Ld, cur, h); }
        *need_ray = !occluded; *L = Ld; *term = tm; *t = td;
        return v3(0.0f);          // shade as if occluded; the visible outcome is formed by the caller
    }
};
// one path after its nearest hit is final (`found`, `hit` include the plane loop; ps.acc is already clamped, :441):
// shades it as if the light sample were occluded (ps: the next ray, mask, accumulator) and, when a shadow ray is
// needed, returns that ray {P, L, t} and the accumulator for the other outcome (acc_v)
struct ShadowOut { V3 P, L, acc_v; float t; };
template <bool TB = true>
VKRT_DEV void logic_compute_k(const DevScene &sc, const RenderParams &rp, V3 cam_pos, PathState &ps,
                              const Hit &hit, bool found, uint32_t skey, Stats &st, bool &alive, bool &need_ray, ShadowOut &so)
{
    alive = false; need_ray = false;
    if (found) {
        const V3 acc_b = ps.acc, mask_b = ps.mask;
        so.P = madd3(hit.t, ps.d, ps.o);                                                // P == surface_of's P
        V3 term = v3(0.0f), emis = v3(0.0f);
        so.L = v3(0.0f); so.t = 0.0f;
        const LightsDeferred<TB> lights{sc, cam_pos, skey, ps.depth * DIMS_PER_BOUNCE, st, &need_ray, &so.L, &term, &so.t};
        alive = path_shade(sc, cam_pos, rp.max_depth, skey, ps, hit, lights, &emis);
        if (need_ray) so.acc_v = acc_b + mask_b * (emis + (v3(0.0f) + term));
    }
}
template <bool TB = true>
VKRT_DEV void logic_compute(const DevScene &sc, const RenderParams &rp, const WaveParams &wp, V3 cam_pos, PathState &ps,
                            const Hit &hit, bool found, uint32_t pix, uint32_t sl, Stats &st, bool &alive, bool &need_ray, ShadowOut &so)
{
    logic_compute_k<TB>(sc, rp, cam_pos, ps, hit, found, sample_key(rp.fkey, pix, wp.s0 + sl), st, alive, need_ray, so);
}
VKRT_DEV void logic_path(const DevScene &sc, const RenderParams &rp, const WaveParams &wp, V3 cam_pos, uint32_t path, PathState &ps,
                         const Hit &hit, bool found, uint32_t pix, uint32_t sl, Stats &st, bool &alive, bool &need_ray)
{
    ShadowOut so;
    logic_compute(sc, rp, wp, cam_pos, ps, hit, found, pix, sl, st, alive, need_ray, so);
    if (need_ray) {
        float4 *sr = wp.shrec + 4 * (size_t)path;
        st256(sr, make_float4(so.P.x, so.P.y, so.P.z, so.t), make_float4(so.L.x, so.L.y, so.L.z, __uint_as_float(alive ? 0u : 1u)));
        sr[2] = make_float4(so.acc_v.x, so.acc_v.y, so.acc_v.z, __uint_as_float(pix));
    }
    if (alive) store_path(wp, path, ps, pix, sl);
    else wp.rad[path] = make_float4(ps.acc.x, ps.acc.y, ps.acc.z, 0.f);               // miss (:445) or the path ended
}
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_LOGIC_MINBLOCKS) k_wf_logic(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                      const __grid_constant__ WaveParams wp, const uint32_t *__restrict__ queue,
                                                      const uint32_t *__restrict__ n_ptr, uint32_t next)
{
    Stats st; stats_zero(st);
    const uint32_t n = *n_ptr;
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool alive = false, need_ray = false;
        uint32_t path = 0, oct = 0;
        if (i < n) {
            path = queue[i];
            PathState ps; Hit hit; uint32_t pix, sl;
            load_path(wp, path, ps, hit, pix, sl);
            ps.acc = clamp3(ps.acc, 0.0f, 1.0f);                                                        // :441
            const uint32_t id0 = (hit.kind << 28) | hit.index;
            float cur = hit.t;
            const bool found = trace_planes<true>(sc, ps.o, ps.d, cur, hit) || id0 != 0u;             // :414-428
            hit.t = cur;
            if (rp.hit_ids && ps.depth == 0u && sl == 0u && wp.s0 == rp.s_begin) rp.hit_ids[pix] = found ? ((hit.kind << 28) | hit.index) : 0u;
            logic_path(sc, rp, wp, cam_pos, path, ps, hit, found, pix, sl, st, alive, need_ray);
            oct = octant_of(ps.d);              // the NEXT ray's direction when the path is alive
        }
#if VKRT_OCTANT_BIN
        push_block_binned<8>(wp.q_active[next], wp.cnt_next + C_ACTIVE, alive, path, oct);
        {
            uint32_t *const qs[1] = {wp.q_shadow}; uint32_t *const cs[1] = {wp.cnt + C_SHADOW};
            const bool ws[1] = {need_ray}; const uint32_t vs[1] = {path};
            push_block<1>(qs, cs, ws, vs);
        }
#else
        (void)oct;
        uint32_t *const qs[2] = {wp.q_active[next], wp.q_shadow}; uint32_t *const cs[2] = {wp.cnt_next + C_ACTIVE, wp.cnt + C_SHADOW};
        const bool ws[2] = {alive, need_ray}; const uint32_t vs[2] = {path, path};
        push_block<2>(qs, cs, ws, vs);
#endif
    }
    wf_flush(st, rp.counters, false);
}

// generate + logic of depth 0 in one pass (fused pipeline): every sample of a pixel starts from the same primary hit
// (Tracer.comp:574-581), so the query, the plane loop, the surface and the material are found once per pixel and
// the S samples are shaded from registers; no depth-0 path record is written and read back.
template <bool BVH, bool STATS>
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_LOGIC_MINBLOCKS) k_wf_generate_logic(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                                                      const __grid_constant__ WaveParams wp)
{
    Stats st; stats_zero(st);
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px = 0, py = 0;
    const bool valid = slot < wp.n_slots && slot_to_pixel_w(rp, slot, px, py);
    V3 o = v3(0.f), d = v3(0.f);
    Hit hit{0.f, 0, 0};
    bool found = false;
    uint32_t pix = 0;
    if (valid) {
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        float cur = path_tmax(0);
        found = trace_tris<true>(sc, o, d, cur, hit);
        if (BVH) {
            const SBest b = bvh_query<false, STATS>(sc, o, d, 1e-3f, sphere_bound<true>(cur), st);
            if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
        } else {
            for (uint32_t i = 0; i < sc.n_spheres; ++i) {                         // literal loop, Tracer.comp:398-412
                const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
                if ((t > 1e-3f) && (t < cur + 1e-3f)) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
            }
        }
        found = trace_planes<true>(sc, o, d, cur, hit) || found;                 // :414-428
        hit.t = cur;
        pix = py * rp.width + px;
        if (rp.hit_ids && wp.s0 == rp.s_begin) rp.hit_ids[pix] = found ? ((hit.kind << 28) | hit.index) : 0u;
        st.closest += wp.S;
        st.shared += wp.S - 1u;
        st.paths += wp.S;
    }
    for (uint32_t sl = 0; sl < wp.S; ++sl) {
        const uint32_t path = sl * wp.n_slots + slot;
        bool alive = false, need_ray = false;
        uint32_t oct = 0;
        if (valid) {
            PathState ps;
            path_begin(ps, o, d);                   // acc = 0: the firefly clamp of :441 leaves it unchanged
            logic_path(sc, rp, wp, cam_pos, path, ps, hit, found, pix, sl, st, alive, need_ray);
            oct = octant_of(ps.d);
        }
#if VKRT_OCTANT_BIN
        push_block_binned<8>(wp.q_active[1], wp.cnt_next + C_ACTIVE, alive, path, oct);
        {
            uint32_t *const qs[1] = {wp.q_shadow}; uint32_t *const cs[1] = {wp.cnt + C_SHADOW};
            const bool ws[1] = {need_ray}; const uint32_t vs[1] = {path};
            push_block<1>(qs, cs, ws, vs);
        }
#else
        (void)oct;
        uint32_t *const qs[2] = {wp.q_active[1], wp.q_shadow}; uint32_t *const cs[2] = {wp.cnt_next + C_ACTIVE, wp.cnt + C_SHADOW};
        const bool ws[2] = {alive, need_ray}; const uint32_t vs[2] = {path, path};
        push_block<2>(qs, cs, ws, vs);
#endif
    }
    wf_flush(st, rp.counters, STATS);
}

// ---- dense fused pipeline (VKRT_DENSE) ---------------------------------------------------------------------
// The same two kernels per depth, but the survivors of a depth are written densely: the block reserves its
// range of the next depth's arrays with one atomicAdd (reserve_block) and every thread stores its path's ray and
// state at its position there.  Item i of a depth is ray i / state i / hit i -- no path-id queues, no indirection:
// `logic` reads and writes contiguous streams, the trace kernel fetches its ray records from a contiguous array.
// A shadow record carries where its "visible" accumulator has to go (the path's state in the next depth's array, or
// the finished radiance of (sample, slot) when the path ended).
template <int NQ>
VKRT_DEV void reserve_block(uint32_t *const (&count)[NQ], const bool (&want)[NQ], uint32_t (&pos)[NQ])
{
#if VKRT_WARP_RESERVE
    // one atomicAdd per warp and output: no block-wide barrier, so a warp never waits for the slowest warp of its block
    // (the barrier of the block-wide form was 14 % of the logic kernels' stall samples); the L2 atomic unit takes about
    // one same-address operation per cycle, a launch issues one per 32 paths
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const unsigned m = __ballot_sync(full, want[q]);
        uint32_t b = 0;
        if (lane == 0 && m) b = atomicAdd(count[q], (uint32_t)__popc(m));
        pos[q] = __shfl_sync(full, b, 0) + (uint32_t)__popc(m & ((1u << lane) - 1u));
    }
#else
    __shared__ uint32_t s_cnt[NQ][32];
    __shared__ uint32_t s_base[NQ];
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
    unsigned m[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        m[q] = __ballot_sync(full, want[q]);
        if (lane == 0) s_cnt[q][warp] = (uint32_t)__popc(m[q]);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        uint32_t tot = 0;
        for (unsigned w = 0; w < n_warps; ++w) { const uint32_t c = s_cnt[threadIdx.x][w]; s_cnt[threadIdx.x][w] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomicAdd(count[threadIdx.x], tot) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) pos[q] = s_base[q] + s_cnt[q][warp] + (uint32_t)__popc(m[q] & ((1u << lane) - 1u));
    __syncthreads();      // s_cnt / s_base are reused by the next call
#endif
}
// every thread asks for `c` consecutive positions; the block reserves the sum with one atomicAdd and hands out
// the ranges in thread order
VKRT_DEV uint32_t reserve_block_ranges(uint32_t *count, uint32_t c)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_b;
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(full, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31u) s_w[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (unsigned w = 0; w < n_warps; ++w) { const uint32_t t = s_w[w]; s_w[w] = tot; tot += t; }
        s_b = tot ? atomicAdd(count, tot) : 0u;
    }
    __syncthreads();
    const uint32_t r = s_b + s_w[warp] + x - c;
    __syncthreads();
    return r;
}
// stores what logic_compute produced for one path: the next depth's ray / state at position j, the shadow record at
// position k, or the finished radiance of (sample, slot)
VKRT_DEV void dense_store(const WaveParams &wp, const PathState &ps, const ShadowOut &so, bool alive, bool need_ray, uint32_t j, uint32_t k,
                          uint32_t slot, uint32_t sl, uint32_t skey)
{
    const uint32_t rad_i = sl * wp.n_slots + slot;
    if (alive) {
        // the ray record's spare word carries the path's sample key: `logic` needs neither the pixel of the slot (two
        // integer divisions) nor the two hashes of sample_key again
        st256(wp.n_ray + 2 * (size_t)j, make_float4(ps.o.x, ps.o.y, ps.o.z, __uint_as_float(skey)), make_float4(ps.d.x, ps.d.y, ps.d.z, 0.f));
        st256(wp.n_st + 2 * (size_t)j, make_float4(ps.acc.x, ps.acc.y, ps.acc.z, __uint_as_float(slot)),
              make_float4(ps.mask.x, ps.mask.y, ps.mask.z, __uint_as_float((sl << 8) | ps.depth)));
    } else {
        wp.rad[rad_i] = make_float4(ps.acc.x, ps.acc.y, ps.acc.z, 0.f);                          // miss (:445) or the path ended
    }
    if (need_ray) {
        float4 *sr = wp.x_shr + 4 * (size_t)k;
        st256(sr, make_float4(so.P.x, so.P.y, so.P.z, so.t), make_float4(so.L.x, so.L.y, so.L.z, __uint_as_float(alive ? j : (0x80000000u | rad_i))));
        sr[2] = make_float4(so.acc_v.x, so.acc_v.y, so.acc_v.z, __uint_as_float(slot));
    }
}
template <bool TB>
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_LOGIC_MINBLOCKS) k_wfd_logic(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                       const __grid_constant__ WaveParams wp, const uint32_t *__restrict__ n_ptr)
{
    Stats st; stats_zero(st);
    const uint32_t n = *n_ptr;
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool alive = false, need_ray = false;
        PathState ps; ShadowOut so;
        uint32_t slot = 0, sl = 0, skey = 0;
        if (i < n) {
            float4 fo, fd, fa, fm;
            ld256cg(wp.x_ray + 2 * (size_t)i, fo, fd);
            ld256cg(wp.x_st + 2 * (size_t)i, fa, fm);
            const float2 h = __ldcg(wp.x_hit + i);
            ps.o = xyz(fo); ps.d = xyz(fd); ps.acc = xyz(fa); ps.mask = xyz(fm);
            const uint32_t sd = __float_as_uint(fm.w);
            ps.depth = sd & 255u; sl = sd >> 8;
            slot = __float_as_uint(fa.w);
            skey = __float_as_uint(fo.w);
            const uint32_t id0 = __float_as_uint(h.y);
            Hit hit{h.x, id0 >> 28, id0 & 0x0fffffffu};
            ps.acc = clamp3(ps.acc, 0.0f, 1.0f);                                                        // :441
            float cur = hit.t;
            const bool found = trace_planes<true>(sc, ps.o, ps.d, cur, hit) || id0 != 0u;             // :414-428
            hit.t = cur;
            logic_compute_k<TB>(sc, rp, cam_pos, ps, hit, found, skey, st, alive, need_ray, so);
        }
        uint32_t *const cs[2] = {wp.cnt_next + C_ACTIVE, wp.cnt + C_SHADOW};
        const bool ws[2] = {alive, need_ray};
        uint32_t pos[2];
        reserve_block<2>(cs, ws, pos);
        if (i < n) dense_store(wp, ps, so, alive, need_ray, pos[0], pos[1], slot, sl, skey);
    }
    wf_flush(st, rp.counters, false);
}
// generate + logic of depth 0, dense: like k_wf_generate_logic, the survivors go to the depth-1 arrays
template <bool BVH, bool STATS, bool TB>
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_LOGIC_MINBLOCKS) k_wfd_generate_logic(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                                                       const __grid_constant__ WaveParams wp)
{
    Stats st; stats_zero(st);
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px = 0, py = 0;
    const bool valid = slot < wp.n_slots && slot_to_pixel_w(rp, slot, px, py);
    V3 o = v3(0.f), d = v3(0.f);
    Hit hit{0.f, 0, 0};
    bool found = false;
    uint32_t pix = 0;
    if (valid) {
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        if (wp.prim_mode == 2u) {                   // a later wave of the frame: the pixel's primary hit is known
            const float2 ph = __ldcg(wp.prim + slot);
            const uint32_t id = __float_as_uint(ph.y);
            hit.t = ph.x; hit.kind = id >> 28; hit.index = id & 0x0fffffffu; found = id != 0u;
        } else {
            float cur = path_tmax(0);
            found = trace_tris<true, false, TB && BVH>(sc, o, d, cur, hit);
            if (BVH) {
                const SBest b = bvh_query<false, STATS>(sc, o, d, 1e-3f, sphere_bound<true>(cur), st);
                if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
            } else {
                for (uint32_t i = 0; i < sc.n_spheres; ++i) {                         // literal loop, Tracer.comp:398-412
                    const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
                    if ((t > 1e-3f) && (t < cur + 1e-3f)) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
                }
            }
            found = trace_planes<true>(sc, o, d, cur, hit) || found;                 // :414-428
            hit.t = cur;
            if (wp.prim_mode == 1u) __stcg(wp.prim + slot, make_float2(cur, __uint_as_float(found ? ((hit.kind << 28) | hit.index) : 0u)));
        }
        pix = py * rp.width + px;
        if (rp.hit_ids && wp.s0 == rp.s_begin) rp.hit_ids[pix] = found ? ((hit.kind << 28) | hit.index) : 0u;
        st.closest += wp.S;
        st.shared += wp.S - (wp.prim_mode == 2u ? 0u : 1u);
        st.paths += wp.S;
    }
#if VKRT_GEN_PIXEL_MAJOR
    // Whether sample sl of this pixel survives depth 0 is known ahead of its shading: path_shade's Russian roulette compares
    // u01(sample key, SLOT_RR) with max3(mask * albedo), mask = (1, 1, 1), of the pixel's one primary hit (Tracer.comp:547).
    // So a thread can reserve one contiguous range for all its surviving samples: the depth-1 rays of a pixel -- same
    // origin -- are then neighbours in the dense arrays, hence in the warps of the traversal kernel and of `logic`.
    float p_rr = 0.f;
    bool can_live = false;
    if (valid && found) {
        const Surface sf = surface_of(sc, o, d, hit);
        const Material mat = load_material(sc, sf.mat);
        p_rr = max3(v3(1.0f) * mat.albedo);
        can_live = 1u < rp.max_depth;
    }
    for (uint32_t g0 = 0; g0 < wp.S; g0 += 32u) {
        const uint32_t gn = wp.S - g0 < 32u ? wp.S - g0 : 32u;
        uint32_t amask = 0;
        if (can_live)
            for (uint32_t k = 0; k < gn; ++k)
                if (!(u01(sample_key(rp.fkey, pix, wp.s0 + g0 + k), SLOT_RR) > p_rr)) amask |= 1u << k;
        const uint32_t j0 = reserve_block_ranges(wp.cnt_next + C_ACTIVE, (uint32_t)__popc(amask));
        for (uint32_t k = 0; k < gn; ++k) {
            const uint32_t sl = g0 + k;
            bool alive = false, need_ray = false;
            PathState ps; ShadowOut so;
            const uint32_t skey = sample_key(rp.fkey, pix, wp.s0 + sl);
            if (valid) {
                path_begin(ps, o, d);                   // acc = 0: the firefly clamp of :441 leaves it unchanged
                logic_compute_k<TB && BVH>(sc, rp, cam_pos, ps, hit, found, skey, st, alive, need_ray, so);
            }
            uint32_t *const cs[1] = {wp.cnt + C_SHADOW};
            const bool ws[1] = {need_ray};
            uint32_t pos[1];
            reserve_block<1>(cs, ws, pos);
            if (valid) dense_store(wp, ps, so, alive, need_ray, j0 + (uint32_t)__popc(amask & ((1u << k) - 1u)), pos[0], slot, sl, skey);
        }
    }
#else
    for (uint32_t sl = 0; sl < wp.S; ++sl) {
        bool alive = false, need_ray = false;
        PathState ps; ShadowOut so;
        if (valid) {
            path_begin(ps, o, d);                   // acc = 0: the firefly clamp of :441 leaves it unchanged
            logic_compute<TB && BVH>(sc, rp, wp, cam_pos, ps, hit, found, pix, sl, st, alive, need_ray, so);
        }
        uint32_t *const cs[2] = {wp.cnt_next + C_ACTIVE, wp.cnt + C_SHADOW};
        const bool ws[2] = {alive, need_ray};
        uint32_t pos[2];
        reserve_block<2>(cs, ws, pos);
        if (valid) dense_store(wp, ps, so, alive, need_ray, pos[0], pos[1], slot, sl, sample_key(rp.fkey, pix, wp.s0 + sl));
    }
#endif
    wf_flush(st, rp.counters, STATS);
}

// depth 0 split in two (VKRT_GEN_PER_SAMPLE): the primary hit of every owned pixel, one thread per pixel slot ...
template <bool BVH, bool STATS>
__global__ void __launch_bounds__(256) k_wfd_primary(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                      const __grid_constant__ WaveParams wp)
{
    Stats st; stats_zero(st);
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t px = 0, py = 0;
    if (slot < wp.n_slots && slot_to_pixel_w(rp, slot, px, py)) {
        V3 o, d;
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        Hit hit{0.f, 0, 0};
        float cur = path_tmax(0);
        bool found = trace_tris<true>(sc, o, d, cur, hit);
        if (BVH) {
            const SBest b = bvh_query<false, STATS>(sc, o, d, 1e-3f, sphere_bound<true>(cur), st);
            if (b.idx >= 0) { cur = b.t; hit.kind = KIND_SPHERE; hit.index = (uint32_t)b.idx; found = true; }
        } else {
            for (uint32_t i = 0; i < sc.n_spheres; ++i) {                         // literal loop, Tracer.comp:398-412
                const float t = sphere_intersect(o, d, __ldg(sc.spheres + i));
                if ((t > 1e-3f) && (t < cur + 1e-3f)) { cur = t; hit.kind = KIND_SPHERE; hit.index = i; found = true; }
            }
        }
        found = trace_planes<true>(sc, o, d, cur, hit) || found;                 // :414-428
        const uint32_t id = found ? ((hit.kind << 28) | hit.index) : 0u;
        wp.prim[slot] = make_float2(cur, __uint_as_float(id));
        if (rp.hit_ids) rp.hit_ids[py * rp.width + px] = id;
    }
    wf_flush(st, rp.counters, STATS);
}
// ... and the shading of depth 0, one thread per (pixel slot, sample of the wave): thread = slot * S + sample
__global__ void __launch_bounds__(VKRT_SHADE_BLOCK, VKRT_LOGIC_MINBLOCKS) k_wfd_generate_ps(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                                                    const __grid_constant__ WaveParams wp)
{
    Stats st; stats_zero(st);
    const V3 cam_pos = v3(rp.fd.camera.pos.x, rp.fd.camera.pos.y, rp.fd.camera.pos.z);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t slot = tid / wp.S, sl = tid - slot * wp.S;
    uint32_t px = 0, py = 0;
    const bool valid = slot < wp.n_slots && slot_to_pixel_w(rp, slot, px, py);
    bool alive = false, need_ray = false;
    uint32_t skey = 0;
    PathState ps; ShadowOut so;
    if (valid) {
        V3 o, d;
        primary_ray(rp.fd, rp.width, rp.height, px, py, o, d);
        const float2 ph = __ldg(wp.prim + slot);
        const uint32_t id = __float_as_uint(ph.y);
        const Hit hit{ph.x, id >> 28, id & 0x0fffffffu};
        path_begin(ps, o, d);                   // acc = 0: the firefly clamp of :441 leaves it unchanged
        skey = sample_key(rp.fkey, py * rp.width + px, wp.s0 + sl);
        logic_compute_k(sc, rp, cam_pos, ps, hit, id != 0u, skey, st, alive, need_ray, so);
        ++st.closest; ++st.paths;
        if (sl != 0u || wp.prim_mode == 2u) ++st.shared;      // every sample but the frame's first reuses the pixel's one query
    }
    uint32_t *const cs[2] = {wp.cnt_next + C_ACTIVE, wp.cnt + C_SHADOW};
    const bool ws[2] = {alive, need_ray};
    uint32_t pos[2];
    reserve_block<2>(cs, ws, pos);
    if (valid) dense_store(wp, ps, so, alive, need_ray, pos[0], pos[1], slot, sl, skey);
    wf_flush(st, rp.counters, false);
}

// ---- reduce: per pixel, add the wave's samples in sample order -------------------------------------
__global__ void __launch_bounds__(256) k_wf_reduce(const __grid_constant__ RenderParams rp, const __grid_constant__ WaveParams wp,
                                                    float4 *__restrict__ frame_sum, uint32_t first_wave, uint32_t last_wave)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= wp.n_slots) return;
    uint32_t px, py;
    if (!slot_to_pixel_w(rp, slot, px, py)) return;
    V3 sum = first_wave ? v3(0.0f) : xyz(frame_sum[slot]);
    for (uint32_t s = 0; s < wp.S; ++s) sum = sum + xyz(wp.rad[(size_t)s * wp.n_slots + slot]);      // Tracer.comp:580
    if (!last_wave) { frame_sum[slot] = make_float4(sum.x, sum.y, sum.z, 0.f); return; }
    const uint32_t pix = py * rp.width + px;
    float4 a = make_float4(sum.x, sum.y, sum.z, (float)(rp.s_end - rp.s_begin));
    if (rp.accumulate) { const float4 o = rp.accum[pix]; a.x = o.x + a.x; a.y = o.y + a.y; a.z = o.z + a.z; a.w = o.w + a.w; }
    rp.accum[pix] = a;
}

// ---------------------------------------------------------------------------------------------------
cudaError_t wave_alloc(WaveBuffers &wb, size_t capacity)
{
    wb = WaveBuffers{};
    wb.capacity = capacity;
    cudaError_t e;
#define A(ptr, bytes) do { e = cudaMalloc((void **)&(ptr), (bytes)); if (e != cudaSuccess) { wave_free(wb); return e; } } while (0)
    // the path-record arrays depend on the pipeline the scene takes (lane_prepare)
    A(wb.sample_rad, capacity * sizeof(float4));
    A(wb.counts, (size_t)C_SETS * C_N * sizeof(uint32_t));
#undef A
    return cudaSuccess;
}

void wave_free(WaveBuffers &wb)
{
    cudaFree(wb.rec); cudaFree(wb.sample_rad);
    cudaFree(wb.queue[0]); cudaFree(wb.queue[1]); cudaFree(wb.queue_mat[0]); cudaFree(wb.queue_mat[1]); cudaFree(wb.counts);
    cudaFree(wb.shadow); cudaFree(wb.occ); cudaFree(wb.queue_shadow); cudaFree(wb.term); cudaFree(wb.shrec);
    cudaFree(wb.d_ray[0]); cudaFree(wb.d_ray[1]); cudaFree(wb.d_st[0]); cudaFree(wb.d_st[1]); cudaFree(wb.d_hit); cudaFree(wb.d_shr);
    for (cudaEvent_t e : wb.ev) cudaEventDestroy(e);
    wb = WaveBuffers{};
}

// lazily sizes the per-scene buffers of one lane: the dense arrays of the fused pipeline, or the indexed path records,
// queues and per-light shadow arrays of the four-kernel one
static cudaError_t lane_prepare(WaveBuffers &wb, uint32_t nl, bool dense, bool fused)
{
    cudaError_t e;
    wb.n_ev = 0;
#define A(ptr, bytes) do { if (!(ptr) && (e = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) return e; } while (0)
    if (dense) {
        for (int k = 0; k < 2; ++k) { A(wb.d_ray[k], wb.capacity * 2 * sizeof(float4)); A(wb.d_st[k], wb.capacity * 2 * sizeof(float4)); }
        A(wb.d_hit, wb.capacity * sizeof(float2));
        A(wb.d_shr, wb.capacity * 4 * sizeof(float4));
        return cudaSuccess;
    }
    A(wb.rec, wb.capacity * 4 * sizeof(float4));
    A(wb.queue[0], wb.capacity * sizeof(uint32_t));
    A(wb.queue[1], wb.capacity * sizeof(uint32_t));
    A(wb.queue_mat[0], wb.capacity * sizeof(uint32_t));
    A(wb.queue_mat[1], wb.capacity * sizeof(uint32_t));
    if (fused) A(wb.shrec, wb.capacity * 4 * sizeof(float4));
#undef A
    if (wb.shadow_lights < nl) {
        cudaFree(wb.shadow); cudaFree(wb.occ); cudaFree(wb.queue_shadow); cudaFree(wb.term);
        wb.shadow = nullptr; wb.occ = nullptr; wb.queue_shadow = nullptr; wb.term = nullptr;
        if ((e = cudaMalloc((void **)&wb.shadow, wb.capacity * nl * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&wb.term, wb.capacity * nl * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&wb.occ, wb.capacity * nl)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&wb.queue_shadow, wb.capacity * nl * sizeof(uint32_t))) != cudaSuccess) return e;
        wb.shadow_lights = nl;
    }
    return cudaSuccess;
}

cudaError_t wave_engine_init(WaveEngine &eng, size_t lane_capacity, uint32_t n_lanes)
{
    cudaError_t e;
    eng = WaveEngine{};
    eng.n_lanes = n_lanes;
    for (uint32_t l = 0; l < n_lanes; ++l) {
        if ((e = wave_alloc(eng.lane[l], lane_capacity)) != cudaSuccess) { wave_engine_free(eng); return e; }
        if (n_lanes > 1) {
            if ((e = cudaStreamCreateWithFlags(&eng.stream[l], cudaStreamNonBlocking)) != cudaSuccess) { wave_engine_free(eng); return e; }
            if ((e = cudaEventCreateWithFlags(&eng.ev_reduce[l], cudaEventDisableTiming)) != cudaSuccess) { wave_engine_free(eng); return e; }
        }
    }
    if (n_lanes > 1 && (e = cudaEventCreateWithFlags(&eng.ev_fork, cudaEventDisableTiming)) != cudaSuccess) { wave_engine_free(eng); return e; }
    if (n_lanes > 1 && (e = cudaEventCreateWithFlags(&eng.ev_prim, cudaEventDisableTiming)) != cudaSuccess) { wave_engine_free(eng); return e; }
    for (uint32_t l = 0; l < n_lanes && n_lanes > 1; ++l)
        if ((e = cudaEventCreateWithFlags(&eng.ev_stage[l], cudaEventDisableTiming)) != cudaSuccess) { wave_engine_free(eng); return e; }
    return cudaSuccess;
}

void wave_engine_free(WaveEngine &eng)
{
    for (uint32_t l = 0; l < 4; ++l) {
        wave_free(eng.lane[l]);
        if (eng.stream[l]) { cudaStreamSynchronize(eng.stream[l]); cudaStreamDestroy(eng.stream[l]); }
        if (eng.ev_reduce[l]) cudaEventDestroy(eng.ev_reduce[l]);
        if (eng.ev_stage[l]) cudaEventDestroy(eng.ev_stage[l]);
    }
    if (eng.ev_fork) cudaEventDestroy(eng.ev_fork);
    if (eng.ev_prim) cudaEventDestroy(eng.ev_prim);
    cudaFree(eng.frame_sum); cudaFree(eng.prim[0]); cudaFree(eng.prim[1]);
    eng = WaveEngine{};
}

// A frame is cut into waves of S samples per pixel.  With two lanes the waves alternate between two buffer
// sets on two streams, so the launch gaps and the drain tails of one wave's ~34 small kernels are filled by the
// other wave's kernels (this matters most when a GPU owns only 1/8 of the tiles); the per-pixel sums are still
// formed in sample order because the `reduce` launches are chained with events, wave after wave.
cudaError_t launch_path_wavefront(const DevScene &sc, const RenderParams &rp, WaveEngine &eng, bool bvh, bool stats, bool timing,
                                  int sm_count, cudaStream_t st, cudaEvent_t ev_consumed, cudaEvent_t ev_begin,
                                  cudaStream_t *tail, uint32_t *n_launches)
{
    cudaError_t e;
    uint32_t launches = 0;
    const uint32_t spp = rp.s_end - rp.s_begin;
    const uint32_t nl = sc.n_lights ? sc.n_lights : 1;
    const uint32_t n_lanes = eng.n_lanes;
    // samples per wave: as many as fit a lane, and at least `n_lanes` waves per frame when there are enough samples
    uint32_t S = (uint32_t)(eng.lane[0].capacity / rp.n_work);
    if (S == 0 || eng.lane[0].capacity >= ((size_t)1 << 28)) return cudaErrorMemoryAllocation;   // shadow items pack path << 4
    if (S > spp) S = spp;
    // (a lane large enough for the whole frame means ONE wave per frame: the lanes then alternate between consecutive frames)
    if (n_lanes > 1 && spp >= n_lanes && S < spp && S > (spp + n_lanes - 1) / n_lanes) S = (spp + n_lanes - 1) / n_lanes;
    const uint32_t n_waves = (spp + S - 1) / S;
    // scenes with at most one light take the fused pipeline (logic + mixed trace), the others the four-kernel one
    const bool fused = VKRT_FUSED != 0 && VKRT_LEAF_BATCH != 0 && sc.n_lights <= 1;
    const bool dense = fused && VKRT_DENSE != 0 && VKRT_FUSED_GENERATE != 0;
    for (uint32_t l = 0; l < n_lanes; ++l) if ((e = lane_prepare(eng.lane[l], nl, dense, fused && sc.n_lights == 1)) != cudaSuccess) return e;
    if (n_waves > 1 && !eng.frame_sum) {
        if ((e = cudaMalloc((void **)&eng.frame_sum, (size_t)rp.n_work * sizeof(float4))) != cudaSuccess) return e;
    }
    // the per-pixel primary hit travels from the frame's first wave to its later ones (two copies: consecutive frames overlap)
    const bool per_sample = dense && VKRT_GEN_PER_SAMPLE != 0;
    const bool share_prim = dense && (n_waves > 1 || per_sample);
    for (int k = 0; k < 2 && share_prim; ++k)
        if (!eng.prim[k] && (e = cudaMalloc((void **)&eng.prim[k], (size_t)rp.n_work * sizeof(float2))) != cudaSuccess) return e;
    if (share_prim) eng.prim_flip ^= 1u;

    typedef void (*trace_fn)(const DevScene, const RenderParams, const WaveParams, const uint32_t *, const uint32_t *, const uint32_t *,
                             const uint32_t *, uint32_t *, uint32_t);
    auto pick_trace = [&](int mode) -> trace_fn {
        if (mode == TRACE_SHADOW) return bvh ? (stats ? k_wf_trace<TRACE_SHADOW, true, true> : k_wf_trace<TRACE_SHADOW, true, false>)
                                             : (stats ? k_wf_trace<TRACE_SHADOW, false, true> : k_wf_trace<TRACE_SHADOW, false, false>);
        if (mode == TRACE_MIXED) return bvh ? (stats ? k_wf_trace<TRACE_MIXED, true, true> : k_wf_trace<TRACE_MIXED, true, false>)
                                            : (stats ? k_wf_trace<TRACE_MIXED, false, true> : k_wf_trace<TRACE_MIXED, false, false>);
        if (mode == TRACE_DENSE) {
            if (bvh && sc.tbvh) return stats ? k_wf_trace<TRACE_DENSE, true, true, true> : k_wf_trace<TRACE_DENSE, true, false, true>;
            return bvh ? (stats ? k_wf_trace<TRACE_DENSE, true, true, false> : k_wf_trace<TRACE_DENSE, true, false, false>)
                       : (stats ? k_wf_trace<TRACE_DENSE, false, true, false> : k_wf_trace<TRACE_DENSE, false, false, false>);
        }
        return bvh ? (stats ? k_wf_trace<TRACE_EXTEND, true, true> : k_wf_trace<TRACE_EXTEND, true, false>)
                   : (stats ? k_wf_trace<TRACE_EXTEND, false, true> : k_wf_trace<TRACE_EXTEND, false, false>);
    };
    trace_fn k_extend = pick_trace(dense ? TRACE_DENSE : (fused && sc.n_lights) ? TRACE_MIXED : TRACE_EXTEND),
             k_shadow = pick_trace(dense ? TRACE_DENSE : fused ? TRACE_MIXED : TRACE_SHADOW);
    int occ_e = 0, occ_s = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend, VKRT_TRACE_BLOCK, 0)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_shadow, VKRT_TRACE_BLOCK, 0)) != cudaSuccess) return e;
#if VKRT_TRACE_RESIDENT
    // leave room on every SM for the other lane's streaming kernels (complementary: they wait on DRAM, this is ALU work)
    if (occ_e > VKRT_TRACE_RESIDENT) occ_e = VKRT_TRACE_RESIDENT;
    if (occ_s > VKRT_TRACE_RESIDENT) occ_s = VKRT_TRACE_RESIDENT;
#endif
    const unsigned grid_e = (unsigned)(sm_count * (occ_e > 0 ? occ_e : 1)), grid_s = (unsigned)(sm_count * (occ_s > 0 ? occ_s : 1));
    const unsigned grid_shade = (unsigned)sm_count * (2048u / VKRT_SHADE_BLOCK);

    // with lanes the waves never wait for `st` as a whole (that would serialise consecutive frames): see ev_consumed
    const bool fork = n_lanes > 1;
    static const uint32_t stagger_env = []() { const char *v = std::getenv("VKRT_TUNE_STAGGER"); return v ? (uint32_t)std::atoi(v) : (uint32_t)VKRT_STAGGER; }();
    const uint32_t stagger = (dense && rp.max_depth > 2) ? (stagger_env < rp.max_depth - 1 ? stagger_env : rp.max_depth - 2) : 0u;
    cudaStream_t ls = st;
    for (uint32_t wv = 0; wv < n_waves; ++wv) {
        const uint32_t lane = fork ? (eng.wave_seq++ % n_lanes) : 0u;
        WaveBuffers &wb = eng.lane[lane];
        ls = fork ? eng.stream[lane] : st;
        if (wv == 0) {
            // the primary-hit AOV is written early in the frame: it must not overtake a pending read of the last one
            if (fork && rp.hit_ids && (e = cudaStreamWaitEvent(ls, ev_consumed, 0)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(ev_begin, ls)) != cudaSuccess) return e;
        }
        // VKRT_FLAG_LAUNCH_TIMING: every launch of the wave is bracketed by timing events (roofline timing,
        // vkrt_debug_dump_timeline); without the flag the hot path records none
        cudaError_t ev_err = cudaSuccess;
        auto ev_rec = [&]() {
            if (wb.n_ev == wb.ev.size()) { cudaEvent_t n; const cudaError_t ce = cudaEventCreate(&n); if (ce != cudaSuccess) { ev_err = ce; return; } wb.ev.push_back(n); }
            const cudaError_t ce = cudaEventRecord(wb.ev[wb.n_ev++], ls);
            if (ce != cudaSuccess) ev_err = ce;
        };
        auto ev_open = [&](uint8_t tag) { if (!timing) return; if (wb.ev_tag.size() <= wb.n_ev / 2) wb.ev_tag.resize(wb.n_ev / 2 + 1); wb.ev_tag[wb.n_ev / 2] = tag; ev_rec(); };
        auto ev_close = [&]() { if (timing && (wb.n_ev & 1u)) ev_rec(); };
        WaveParams wp{};
        wp.rec = wb.rec; wp.shrec = wb.shrec; wp.sh = wb.shadow; wp.term = wb.term; wp.rad = wb.sample_rad;
        wp.q_active[0] = wb.queue[0]; wp.q_active[1] = wb.queue[1]; wp.q_diel = wb.queue_mat[0]; wp.q_diff = wb.queue_mat[1];
        wp.q_shadow = wb.queue_shadow;
        wp.occ = wb.occ; wp.cnt = wb.counts; wp.cnt_next = wb.counts + C_N;
        wp.s0 = rp.s_begin + wv * S;
        wp.S = (wp.s0 + S <= rp.s_end) ? S : (rp.s_end - wp.s0);
        wp.n_slots = rp.n_work;
        wp.n_lights = sc.n_lights;
        // staggered lanes: this wave starts when the wave before it (on another lane) has reached depth `stagger`, so one
        // lane's thin, latency-bound deep bounces run next to the other lane's fat early ones instead of next to its own kind
        if (fork && stagger && eng.have_stage && (e = cudaStreamWaitEvent(ls, eng.ev_stage[eng.stage_lane], 0)) != cudaSuccess) return e;
        wp.prim = share_prim ? eng.prim[eng.prim_flip] : nullptr;
        wp.prim_mode = share_prim ? (wv == 0 ? 1u : 2u) : 0u;
        // the later waves run on other streams than the first: they wait for its generate kernel (ev_prim below)
        if (share_prim && wv > 0 && fork && (e = cudaStreamWaitEvent(ls, eng.ev_prim, 0)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(wb.counts, 0, (size_t)(rp.max_depth + 1) * C_N * sizeof(uint32_t), ls)) != cudaSuccess) return e;
        if (dense) {
            // depth 0 inside generate; per depth d >= 1: trace (rays of d + shadow rays of d - 1) -> logic; the last depth's
            // shadow rays at the end.  Arrays of depth d: parity d & 1.
            const bool tb = bvh && sc.tbvh != nullptr;       // scenes with a triangle tree run the kernels that carry its walk
            void (*k_gen)(const DevScene, const RenderParams, const WaveParams) =
                tb ? (stats ? k_wfd_generate_logic<true, true, true> : k_wfd_generate_logic<true, false, true>)
                   : bvh ? (stats ? k_wfd_generate_logic<true, true, false> : k_wfd_generate_logic<true, false, false>)
                         : (stats ? k_wfd_generate_logic<false, true, false> : k_wfd_generate_logic<false, false, false>);
            void (*k_logic)(const DevScene, const RenderParams, const WaveParams, const uint32_t *) = tb ? k_wfd_logic<true> : k_wfd_logic<false>;
            wp.x_hit = wb.d_hit; wp.x_shr = wb.d_shr;
            wp.cnt = wb.counts; wp.cnt_next = wb.counts + C_N;
            wp.n_ray = wb.d_ray[1]; wp.n_st = wb.d_st[1];
            if (per_sample) {
                if (wv == 0) {
                    void (*k_prim)(const DevScene, const RenderParams, const WaveParams) =
                        bvh ? (stats ? k_wfd_primary<true, true> : k_wfd_primary<true, false>)
                            : (stats ? k_wfd_primary<false, true> : k_wfd_primary<false, false>);
                    ev_open(0);
                    k_prim<<<(wp.n_slots + 255u) / 256u, 256, 0, ls>>>(sc, rp, wp); ++launches;
                    ev_close();
                }
                const unsigned long long n_thr = (unsigned long long)wp.n_slots * wp.S;
                ev_open(0);
                k_wfd_generate_ps<<<(unsigned)((n_thr + VKRT_SHADE_BLOCK - 1u) / VKRT_SHADE_BLOCK), VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp); ++launches;
                ev_close();
            } else {
            ev_open(0);
            k_gen<<<(wp.n_slots + VKRT_SHADE_BLOCK - 1u) / VKRT_SHADE_BLOCK, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp); ++launches;
            ev_close();
            }
            if (share_prim && wv == 0 && fork && (e = cudaEventRecord(eng.ev_prim, ls)) != cudaSuccess) return e;
            for (uint32_t depth = 1; depth <= rp.max_depth; ++depth) {
                const uint32_t cur = depth & 1u, nxt = cur ^ 1u;
                wp.cnt = wb.counts + (size_t)depth * C_N;
                wp.cnt_next = wp.cnt + C_N;
                wp.x_ray = wb.d_ray[cur]; wp.x_st = wb.d_st[cur]; wp.n_ray = wb.d_ray[nxt]; wp.n_st = wb.d_st[nxt];
                const bool last = depth == rp.max_depth;          // no rays of this depth exist: only the shadow rays of depth - 1
                if (last && !sc.n_lights) break;
                ev_open(last ? 3 : 1);
                k_extend<<<grid_e, VKRT_TRACE_BLOCK, 0, ls>>>(sc, rp, wp, nullptr, last ? wp.cnt + C_ZERO : wp.cnt + C_ACTIVE, nullptr,
                                                               wp.cnt - C_N + C_SHADOW, wp.cnt + C_HEAD_EXTEND, depth); ++launches;
                ev_close();
                if (last) break;
                ev_open(2);
                k_logic<<<grid_shade, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp, wp.cnt + C_ACTIVE); ++launches;
                ev_close();
                if (fork && stagger && (depth == stagger || (depth + 1 == rp.max_depth && depth < stagger))) {
                    if ((e = cudaEventRecord(eng.ev_stage[lane], ls)) != cudaSuccess) return e;
                    eng.stage_lane = lane; eng.have_stage = true;
                }
            }
        } else if (fused && VKRT_FUSED_GENERATE) {
            void (*k_gen)(const DevScene, const RenderParams, const WaveParams) =
                bvh ? (stats ? k_wf_generate_logic<true, true> : k_wf_generate_logic<true, false>)
                    : (stats ? k_wf_generate_logic<false, true> : k_wf_generate_logic<false, false>);
            wp.cnt = wb.counts; wp.cnt_next = wb.counts + C_N;
            ev_open(0);
            k_gen<<<(wp.n_slots + VKRT_SHADE_BLOCK - 1u) / VKRT_SHADE_BLOCK, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp); ++launches;
            ev_close();
        } else {
            void (*k_gen)(const DevScene, const RenderParams, const WaveParams) =
                bvh ? (stats ? k_wf_generate<true, true> : k_wf_generate<true, false>)
                    : (stats ? k_wf_generate<false, true> : k_wf_generate<false, false>);
            ev_open(0);
            k_gen<<<(wp.n_slots + 255u) / 256u, 256, 0, ls>>>(sc, rp, wp); ++launches;
            ev_close();
        }
        for (uint32_t depth = 0; depth < rp.max_depth && !dense; ++depth) {
            const uint32_t cur = depth & 1u, nxt = cur ^ 1u;
            wp.cnt = wb.counts + (size_t)depth * C_N;
            wp.cnt_next = wp.cnt + C_N;
            const uint32_t *n_active = wp.cnt + C_ACTIVE;
            if (fused) {
                // this depth's nearest-hit rays and the previous depth's shadow rays in one launch, then `logic`
                if (depth > 0) {
                    ev_open(1);
                    k_extend<<<grid_e, VKRT_TRACE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_active[cur], n_active, wp.q_shadow, wp.cnt - C_N + C_SHADOW,
                                                                   wp.cnt + C_HEAD_EXTEND, depth); ++launches;
                    ev_close();
                }
                if (depth > 0 || !VKRT_FUSED_GENERATE) {     // depth 0 was shaded by generate_logic
                    ev_open(2);
                    k_wf_logic<<<grid_shade, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_active[cur], n_active, nxt); ++launches;
                    ev_close();
                }
                if (depth + 1 == rp.max_depth && sc.n_lights) {      // the last depth's shadow rays (no nearest-hit items left)
                    ev_open(3);
                    k_shadow<<<grid_s, VKRT_TRACE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_active[nxt], wp.cnt_next + C_ZERO, wp.q_shadow, wp.cnt + C_SHADOW,
                                                                   wp.cnt_next + C_HEAD_EXTEND, depth + 1); ++launches;
                    ev_close();
                }
                continue;
            }
            if (depth > 0) {       // depth 0 was traced once per pixel by `generate`
                ev_open(1);
                k_extend<<<grid_e, VKRT_TRACE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_active[cur], n_active, nullptr, nullptr, wp.cnt + C_HEAD_EXTEND, depth); ++launches;
                ev_close();
            }
            ev_open(2);
            k_wf_classify<<<grid_shade, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_active[cur], n_active); ++launches;
            ev_close();
            if (sc.n_lights) {
                ev_open(3);
                k_shadow<<<grid_s, VKRT_TRACE_BLOCK, 0, ls>>>(sc, rp, wp, wp.q_shadow, wp.cnt + C_SHADOW, nullptr, nullptr, wp.cnt + C_HEAD_SHADOW, depth); ++launches;
                ev_close();
            }
            ev_open(4);
            k_wf_shade<<<grid_shade, VKRT_SHADE_BLOCK, 0, ls>>>(sc, rp, wp, nxt); ++launches;
            ev_close();
        }
        // the running per-pixel sum continues in wave order: wait for the previous wave's reduce (also across
        // frames: frame_sum is one buffer); the last reduce overwrites the accumulator, which the previous frame's
        // resolve / pack / read-back may still be reading
        if (fork && eng.have_prev_reduce && (e = cudaStreamWaitEvent(ls, eng.ev_reduce[eng.prev_reduce_lane], 0)) != cudaSuccess) return e;
        if (fork && wv + 1 == n_waves && (e = cudaStreamWaitEvent(ls, ev_consumed, 0)) != cudaSuccess) return e;
        ev_open(5);
        k_wf_reduce<<<(wp.n_slots + 255u) / 256u, 256, 0, ls>>>(rp, wp, eng.frame_sum, wv == 0, wv + 1 == n_waves); ++launches;
        ev_close();
        if (fork) {
            if ((e = cudaEventRecord(eng.ev_reduce[lane], ls)) != cudaSuccess) return e;
            eng.prev_reduce_lane = lane; eng.have_prev_reduce = true;
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (ev_err != cudaSuccess) return ev_err;
    }
    if (tail) *tail = ls;
    if (n_launches) *n_launches = launches;
    return cudaSuccess;
}

} // namespace vkrt
