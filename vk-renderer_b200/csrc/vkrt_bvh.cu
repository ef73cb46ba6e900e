// Device LBVH build over the sphere list: centroid bounds -> 30-bit Morton codes -> LSD radix
// sort (4 x 8 bit, stable) -> Karras 2012 hierarchy -> bottom-up refit with per-node arrival
// counters -> packed 64-byte nodes (vkrt_device.cuh "HBM layout").
//
// New functionality: the reference brute-forces every primitive (Tracer.comp:378-428); what it
// pins is only the nearest-hit semantics the tree has to reproduce (DESIGN.md "Rule S").
// The tree is a pure function of the sphere array (stable sort by (code, index), exact min/max),
// so every GPU that builds it gets the identical tree.
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

#ifndef VKRT_SAH_TREE
#define VKRT_SAH_TREE 1           // the wavefront's traversal nodes come from a binned-SAH tree (0: from the LBVH itself)
#endif
#ifndef VKRT_SAH_BLOCK_LEVELS
#define VKRT_SAH_BLOCK_LEVELS 1   // the top levels of its build give every node a block instead of a warp (0: warps only; same tree)
#endif
#ifndef VKRT_SAH_MIN_PRIMS
#define VKRT_SAH_MIN_PRIMS 2
#endif

namespace vkrt {

// The builder works on any primitive list that can name a centroid (Morton code) and a padded leaf box (refit):
// the spheres (rule S, vkrt_device.cuh) and the triangles of the SSBO (rule T).
struct PrimSpheres {
    const float4 *s;
    __device__ __forceinline__ void centroid(uint32_t i, float *c) const { const float4 v = s[i]; c[0] = v.x; c[1] = v.y; c[2] = v.z; }
    __device__ __forceinline__ void box(uint32_t i, float *lo, float *hi) const
    {
        const float4 v = s[i];
        const float rp = sphere_pad_radius(v.w);
        lo[0] = v.x - rp; lo[1] = v.y - rp; lo[2] = v.z - rp;
        hi[0] = v.x + rp; hi[1] = v.y + rp; hi[2] = v.z + rp;
    }
};
struct PrimTris {
    const float4 *t;      // 3 float4 per triangle (the reference's SSBO layout, Tracer.comp:127-132)
    __device__ __forceinline__ void box(uint32_t i, float *lo, float *hi) const
    {
        V3 l, h;
        tri_padded_box(xyz(t[3 * i]), xyz(t[3 * i + 1]), xyz(t[3 * i + 2]), l, h);
        lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
    }
    __device__ __forceinline__ void centroid(uint32_t i, float *c) const
    {
        float lo[3], hi[3];
        box(i, lo, hi);
        c[0] = 0.5f * (lo[0] + hi[0]); c[1] = 0.5f * (lo[1] + hi[1]); c[2] = 0.5f * (lo[2] + hi[2]);
    }
};

// float <-> order-preserving int, for atomicMin/atomicMax on floats
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_bounds_init(int *b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7fffffff;          // min
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0x80000000; // max
}

template <class P>
__global__ void __launch_bounds__(256) k_bounds(const P prim, uint32_t n, int *b)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float c[3];
        prim.centroid(i, c);
        lo[0] = fminf(lo[0], c[0]); lo[1] = fminf(lo[1], c[1]); lo[2] = fminf(lo[2], c[2]);
        hi[0] = fmaxf(hi[0], c[0]); hi[1] = fmaxf(hi[1], c[1]); hi[2] = fmaxf(hi[2], c[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int l = f2ord(lo[k]), h = f2ord(hi[k]);
        l = __reduce_min_sync(0xffffffffu, l);
        h = __reduce_max_sync(0xffffffffu, h);
        if ((threadIdx.x & 31) == 0) { atomicMin(b + k, l); atomicMax(b + 3 + k, h); }
    }
}

__device__ __forceinline__ uint32_t expand_bits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t quant10(float c, float cmin, float scale)
{
    float q = (c - cmin) * scale;
    q = fminf(fmaxf(q, 0.0f), 1023.0f);
    return (uint32_t)q;
}

template <class P>
__global__ void __launch_bounds__(256) k_morton(const P prim, uint32_t n, const int *__restrict__ b,
                                                 uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float cminx = ord2f(b[0]), cminy = ord2f(b[1]), cminz = ord2f(b[2]);
    const float ex = ord2f(b[3]) - cminx, ey = ord2f(b[4]) - cminy, ez = ord2f(b[5]) - cminz;
    const float sx = ex > 0.0f ? 1024.0f / ex : 0.0f, sy = ey > 0.0f ? 1024.0f / ey : 0.0f, sz = ez > 0.0f ? 1024.0f / ez : 0.0f;
    float c[3];
    prim.centroid(i, c);
    keys[i] = (expand_bits(quant10(c[0], cminx, sx)) << 2) | (expand_bits(quant10(c[1], cminy, sy)) << 1) |
              expand_bits(quant10(c[2], cminz, sz));
    vals[i] = i;
}

// ---- LSD radix sort, 8 bits per pass, stable --------------------------------------------------
enum { RS_THREADS = 256, RS_ROUNDS = 4, RS_TILE = RS_THREADS * RS_ROUNDS, RS_DIGITS = 256 };

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift,
                                                         uint32_t *__restrict__ hist, uint32_t n_blocks)
{
    __shared__ uint32_t h[RS_DIGITS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

// exclusive scan of hist[0 .. m) by one block (m = 256 * n_blocks); fine for a one-off build
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t *__restrict__ hist, uint32_t m)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < m ? hist[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t c = carry;
        if (i < m) hist[i] = c + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warp_off + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                            uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                            uint32_t n, int shift, const uint32_t *__restrict__ hist,
                                                            uint32_t n_blocks)
{
    __shared__ uint32_t warp_cnt[RS_THREADS / 32][RS_DIGITS];
    __shared__ uint32_t running[RS_DIGITS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    running[threadIdx.x] = hist[threadIdx.x * n_blocks + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keys_in[i] : 0xffffffffu;
        const uint32_t dig = valid ? ((key >> shift) & 255u) : 256u;     // 256 = "no digit"
        const unsigned peers = __match_any_sync(0xffffffffu, dig);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[warp][dig] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = running[dig] + rank_in_warp;
            for (unsigned w = 0; w < warp; ++w) off += warp_cnt[w][dig];
            keys_out[off] = key;
            vals_out[off] = vals_in[i];
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) tot += warp_cnt[w][threadIdx.x];
        running[threadIdx.x] += tot;
        __syncthreads();
    }
}

// ---- Karras 2012: one thread per inner node ---------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t *__restrict__ codes, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

// child encoding in the temporary arrays: >= 0 inner node, < 0 leaf ~sorted_position
__global__ void __launch_bounds__(256) k_karras(const uint32_t *__restrict__ codes, int n, int2 *__restrict__ children,
                                                 int *__restrict__ parent_inner, int *__restrict__ parent_leaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1)) > 0 ? 1 : -1;
    const int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    int2 c;
    if (first == gamma) { c.x = ~gamma; parent_leaf[gamma] = i; } else { c.x = gamma; parent_inner[gamma] = i; }
    if (last == gamma + 1) { c.y = ~(gamma + 1); parent_leaf[gamma + 1] = i; } else { c.y = gamma + 1; parent_inner[gamma + 1] = i; }
    children[i] = c;
    if (i == 0) parent_inner[0] = -1;
}

// ---- refit + pack: one thread per leaf walks up; the second arrival at a node owns it ---------
// child record (32 B): {lo.x lo.y lo.z hi.x | hi.y hi.z index kind}; kind 0 = inner node `index`,
// kind 1 = leaf = sphere `index`, whose record holds the sphere's own padded box
__device__ __forceinline__ void write_child(float4 *node, int k, bool leaf, int index, const float *lo, const float *hi)
{
    node[2 * k] = make_float4(lo[0], lo[1], lo[2], hi[0]);
    node[2 * k + 1] = make_float4(hi[1], hi[2], __int_as_float(index), __int_as_float(leaf ? 1 : 0));
}
template <class P>
__global__ void __launch_bounds__(256) k_refit(const P prim, const uint32_t *__restrict__ sorted_idx, int n,
                                                const int2 *__restrict__ children, const int *__restrict__ parent_inner,
                                                const int *__restrict__ parent_leaf, int *__restrict__ arrivals,
                                                float4 *__restrict__ box_lo, float4 *__restrict__ box_hi,
                                                float4 *__restrict__ nodes)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrivals + node, 1) == 0) return;      // the sibling subtree is not done yet
        __threadfence();
        const int2 c = children[node];
        float lo[2][3], hi[2][3];
        int idx[2];
        bool is_leaf[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ch = k == 0 ? c.x : c.y;
            is_leaf[k] = ch < 0;
            if (is_leaf[k]) {
                idx[k] = (int)sorted_idx[~ch];
                prim.box((uint32_t)idx[k], lo[k], hi[k]);
            } else {
                idx[k] = ch;
                const volatile float4 *pl = box_lo + ch, *ph = box_hi + ch;
                lo[k][0] = pl->x; lo[k][1] = pl->y; lo[k][2] = pl->z;
                hi[k][0] = ph->x; hi[k][1] = ph->y; hi[k][2] = ph->z;
            }
            write_child(nodes + 4 * (size_t)node, k, is_leaf[k], idx[k], lo[k], hi[k]);
        }
        box_lo[node] = make_float4(fminf(lo[0][0], lo[1][0]), fminf(lo[0][1], lo[1][1]), fminf(lo[0][2], lo[1][2]), 0.f);
        box_hi[node] = make_float4(fmaxf(hi[0][0], hi[1][0]), fmaxf(hi[0][1], hi[1][1]), fmaxf(hi[0][2], hi[1][2]), 0.f);
        node = parent_inner[node];
    }
}

// depth of the tree = the longest leaf-to-root chain.  The traversal stacks hold one entry per level at most (the
// near-first binary walk pushes the far child and descends), so this bounds them; it cannot exceed 64: the Karras
// hierarchy is the radix tree of the DISTINCT 64-bit keys (Morton code, index), every level fixes at least one more bit.
__global__ void __launch_bounds__(256) k_tree_depth(const int *__restrict__ parent_inner, const int *__restrict__ parent_leaf, int n, int *depth)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (leaf < n) for (int node = parent_leaf[leaf]; node >= 0; node = parent_inner[node]) ++d;
    d = __reduce_max_sync(0xffffffffu, d);
    if ((threadIdx.x & 31) == 0) atomicMax(depth, d);
}

template <class P>
__global__ void k_single_leaf(const P prim, float4 *__restrict__ nodes)
{
    float lo[3], hi[3];
    prim.box(0u, lo, hi);
    write_child(nodes, 0, true, 0, lo, hi);
    write_child(nodes, 1, true, 0, lo, hi);
}

// ---- collapse to 4-wide traversal nodes: node4[i] holds the (up to four) grandchildren of binary node i ------
// Every binary node gets one (indices stay binary node ids, no compaction pass); the traversal only ever reaches
// the ones on even levels below the root.  A leaf child keeps its own slot, empty slots hold an unreachable box.
// Halving the number of dependent node fetches per ray is what matters: the kernels are bound by the latency of
// that chain (a lone 300-step ray was the ~230 us floor of every trace launch).
__global__ void __launch_bounds__(256) k_collapse4(const float4 *__restrict__ nodes, uint32_t n_inner, float4 *__restrict__ nodes4)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
    float4 slot[8];
    int n = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 a = nodes[4 * (size_t)i + 2 * k], b = nodes[4 * (size_t)i + 2 * k + 1];
        if (__float_as_int(b.w) == 1) { slot[2 * n] = a; slot[2 * n + 1] = b; ++n; }
        else {
            const float4 *c = nodes + 4 * (size_t)__float_as_int(b.z);
            slot[2 * n] = c[0]; slot[2 * n + 1] = c[1]; ++n;
            slot[2 * n] = c[2]; slot[2 * n + 1] = c[3]; ++n;
        }
    }
    for (; n < 4; ++n) {
        slot[2 * n] = make_float4(1e30f, 1e30f, 1e30f, 1e30f);
        slot[2 * n + 1] = make_float4(1e30f, 1e30f, __int_as_float(0), __int_as_float(2));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) nodes4[8 * (size_t)i + k] = slot[k];
}

// ---- 32-byte traversal nodes on a 16-bit grid (vkrt_device.cuh "32-byte nodes") -------------------------
// grid[0..2] = s, grid[3..5] = b2; the coordinate a code q stands for is X(q) = (2^23 + q) * s + b2 in REAL
// arithmetic.  Codes are chosen with X(q_lo) <= lo and X(q_hi) >= hi, evaluated in double precision and then
// moved one more step outward (the double evaluation may round), by a monotone function of the coordinate, so
// that the coded boxes of the tree nest exactly like the exact boxes they were made from.
__global__ void k_qgrid(const float4 *__restrict__ nodes, float *__restrict__ grid)
{
    const int a = threadIdx.x;
    if (a >= 3) return;
    // scene box = union of the root's two child boxes: {lo.x lo.y lo.z hi.x | hi.y hi.z ...}
    const float4 a0 = nodes[0], b0 = nodes[1], a1 = nodes[2], b1 = nodes[3];
    const float lo0[3] = {a0.x, a0.y, a0.z}, hi0[3] = {a0.w, b0.x, b0.y}, lo1[3] = {a1.x, a1.y, a1.z}, hi1[3] = {a1.w, b1.x, b1.y};
    const float lo = fminf(lo0[a], lo1[a]), hi = fmaxf(hi0[a], hi1[a]);
    float s = (hi - lo) / 65024.0f;
    if (!(s > 1e-12f)) s = 1e-6f;
    const float base = lo - 128.0f * s;
    grid[a] = s;
    grid[3 + a] = __fmaf_rn(-8388608.0f, s, base);
}
__device__ __forceinline__ double q_coord(int q, float s, float b2) { return (double)(8388608 + q) * (double)s + (double)b2; }
__device__ __forceinline__ uint32_t quantize_lo(float x, float s, float b2)
{
    int q = (int)floor(((double)x - q_coord(0, s, b2)) / (double)s);
    q = min(max(q, 0), 65535);
    while (q > 0 && q_coord(q, s, b2) > (double)x) --q;
    while (q < 65535 && q_coord(q + 1, s, b2) <= (double)x) ++q;
    return (uint32_t)max(q - 1, 0);
}
__device__ __forceinline__ uint32_t quantize_hi(float x, float s, float b2)
{
    int q = (int)ceil(((double)x - q_coord(0, s, b2)) / (double)s);
    q = min(max(q, 0), 65535);
    while (q < 65535 && q_coord(q, s, b2) < (double)x) ++q;
    while (q > 0 && q_coord(q - 1, s, b2) >= (double)x) --q;
    return (uint32_t)min(q + 1, 65535);
}
__global__ void __launch_bounds__(256) k_quantize(const float4 *__restrict__ nodes, uint32_t n_inner, const float *__restrict__ grid,
                                                   uint4 *__restrict__ qnodes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 a = nodes[4 * (size_t)i + 2 * k], b = nodes[4 * (size_t)i + 2 * k + 1];
        const float lo[3] = {a.x, a.y, a.z}, hi[3] = {a.w, b.x, b.y};
        uint32_t w[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) w[c] = quantize_lo(lo[c], grid[c], grid[3 + c]) | (quantize_hi(hi[c], grid[c], grid[3 + c]) << 16);
        const int ref = __float_as_int(b.z) ^ -__float_as_int(b.w);      // inner node index, or ~sphere for a leaf
        qnodes[2 * (size_t)i + k] = make_uint4(w[0], w[1], w[2], (uint32_t)ref);
    }
}

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { err = _e; goto done; } } while (0)

// the hierarchy itself: `nodes` = (n > 1 ? n - 1 : 1) exact 64-byte nodes (caller-allocated), depth, launch count
template <class P>
static cudaError_t build_tree(const P prim, uint32_t n, float4 *nodes, int *depth_out, uint32_t *launches_out, cudaStream_t st)
{
    cudaError_t err = cudaSuccess;
    int *bounds = nullptr, *parent_inner = nullptr, *parent_leaf = nullptr, *arrivals = nullptr, *d_depth = nullptr;
    uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr}, *hist = nullptr;
    int2 *children = nullptr;
    float4 *box_lo = nullptr, *box_hi = nullptr;
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    uint32_t launches = 0;
    *depth_out = 1;
    if (n == 1) {
        k_single_leaf<P><<<1, 1, 0, st>>>(prim, nodes); ++launches;
        CK(cudaGetLastError());
    } else {
        CK(cudaMalloc(&bounds, 6 * sizeof(int)));
        for (int k = 0; k < 2; ++k) { CK(cudaMalloc(&keys[k], n * sizeof(uint32_t))); CK(cudaMalloc(&vals[k], n * sizeof(uint32_t))); }
        CK(cudaMalloc(&hist, (size_t)RS_DIGITS * n_blocks * sizeof(uint32_t)));
        CK(cudaMalloc(&children, (size_t)(n - 1) * sizeof(int2)));
        CK(cudaMalloc(&parent_inner, (size_t)(n - 1) * sizeof(int)));
        CK(cudaMalloc(&parent_leaf, (size_t)n * sizeof(int)));
        CK(cudaMalloc(&arrivals, (size_t)(n - 1) * sizeof(int)));
        CK(cudaMalloc(&box_lo, (size_t)(n - 1) * sizeof(float4)));
        CK(cudaMalloc(&box_hi, (size_t)(n - 1) * sizeof(float4)));
        CK(cudaMalloc(&d_depth, sizeof(int)));
        CK(cudaMemsetAsync(arrivals, 0, (size_t)(n - 1) * sizeof(int), st));
        CK(cudaMemsetAsync(d_depth, 0, sizeof(int), st));

        k_bounds_init<<<1, 32, 0, st>>>(bounds); ++launches;
        {
            unsigned g = (n + 255u) / 256u; if (g > 592u) g = 592u;   // 4 x 148 SMs
            k_bounds<P><<<g, 256, 0, st>>>(prim, n, bounds); ++launches;
        }
        k_morton<P><<<(n + 255u) / 256u, 256, 0, st>>>(prim, n, bounds, keys[0], vals[0]); ++launches;
        int cur = 0;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = pass * 8;
            k_rs_hist<<<n_blocks, RS_THREADS, 0, st>>>(keys[cur], n, shift, hist, n_blocks); ++launches;
            k_rs_scan<<<1, 1024, 0, st>>>(hist, RS_DIGITS * n_blocks); ++launches;
            k_rs_scatter<<<n_blocks, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, hist, n_blocks); ++launches;
            cur ^= 1;
        }
        k_karras<<<(n - 1 + 255u) / 256u, 256, 0, st>>>(keys[cur], (int)n, children, parent_inner, parent_leaf); ++launches;
        k_refit<P><<<(n + 255u) / 256u, 256, 0, st>>>(prim, vals[cur], (int)n, children, parent_inner, parent_leaf, arrivals,
                                                      box_lo, box_hi, nodes); ++launches;
        CK(cudaGetLastError());
        k_tree_depth<<<(n + 255u) / 256u, 256, 0, st>>>(parent_inner, parent_leaf, (int)n, d_depth); ++launches;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(depth_out, d_depth, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));       // the temporaries below are freed when this returns
    }
    *launches_out += launches;
done:
    cudaFree(bounds); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(vals[0]); cudaFree(vals[1]); cudaFree(hist);
    cudaFree(children); cudaFree(parent_inner); cudaFree(parent_leaf); cudaFree(arrivals); cudaFree(box_lo); cudaFree(box_hi);
    cudaFree(d_depth);
    return err;
}


// ---- the traversal tree: top-down binned SAH over the same primitives ---------------------------------------
// The LBVH above is a pure function of the Morton order; on the large scenes the wavefront's traversal kernel walks a
// tree of better quality instead: the same leaves (one primitive each, its own padded box), inner boxes = exact min / max
// unions -- so rule S / rule T return the identical answer (DESIGN.md "Rule S": ANY such hierarchy does) -- but the
// splits minimise the surface-area heuristic over SAH_BINS centroid bins per axis (CPU simulation on the cfg4 scene,
// tests/tools/trav_sim.py b16: 9-10 % fewer node visits, 15 % fewer stack pushes than the LBVH).
// Device build, level-synchronous and deterministic: per level one warp per node (1) bins the node's primitives
// (shared-memory atomicMin / atomicMax on order-preserving integers: order-independent) and picks the cheapest of the
// 3 x (SAH_BINS - 1) planes, (2) a one-block scan numbers the next level's nodes, (3) the warp partitions its index
// range stably and writes the node's two child records.  Node ids are breadth-first (a level's nodes are consecutive).
#ifndef VKRT_SAH_BINS
#define VKRT_SAH_BINS 16
#endif
enum { SAH_BINS = VKRT_SAH_BINS, SAH_WARPS = 8, SAH_MEDIAN_FROM_LEVEL = 48 };
struct SahSplit { int axis, bin; float clo, scale; uint32_t n_left; };      // axis < 0: split the index range in the middle

__device__ __forceinline__ float half_area(const float *lo, const float *hi)
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ int sah_bin(float c, float clo, float scale)
{
    const int b = (int)((c - clo) * scale);
    return min(max(b, 0), SAH_BINS - 1);
}

// One warp evaluates the 3 x (SAH_BINS - 1) planes of a node from its bins and writes the split (lane 0).
// candidate t = axis * (SAH_BINS - 1) + (s - 1): bins [0, s) go left, [s, SAH_BINS) right
__device__ __forceinline__ void sah_pick_plane(const int (*bins)[SAH_BINS][7], const float *clo, const float *scale, uint32_t c, int force_median,
                                               unsigned lane, SahSplit *out)
{
    const unsigned full = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    float best_cost = INF;
    uint32_t best_t = 0xffffffffu, best_nl = 0;
    for (uint32_t t = lane; t < 3u * (SAH_BINS - 1); t += 32u) {
        const int k = (int)(t / (SAH_BINS - 1)), sp = (int)(t % (SAH_BINS - 1)) + 1;
        if (!(scale[k] > 0.0f)) continue;
        float llo[3] = {INF, INF, INF}, lhi[3] = {-INF, -INF, -INF}, rlo[3] = {INF, INF, INF}, rhi[3] = {-INF, -INF, -INF};
        uint32_t nl = 0, nr = 0;
        for (int b = 0; b < SAH_BINS; ++b) {
            const int *bb = bins[k][b];
            if (bb[6] == 0) continue;
            if (b < sp) { nl += (uint32_t)bb[6]; for (int a = 0; a < 3; ++a) { llo[a] = fminf(llo[a], ord2f(bb[a])); lhi[a] = fmaxf(lhi[a], ord2f(bb[3 + a])); } }
            else        { nr += (uint32_t)bb[6]; for (int a = 0; a < 3; ++a) { rlo[a] = fminf(rlo[a], ord2f(bb[a])); rhi[a] = fmaxf(rhi[a], ord2f(bb[3 + a])); } }
        }
        if (nl == 0 || nr == 0) continue;
        const float cost = half_area(llo, lhi) * (float)nl + half_area(rlo, rhi) * (float)nr;
        if (cost < best_cost || (cost == best_cost && t < best_t)) { best_cost = cost; best_t = t; best_nl = nl; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float oc = __shfl_xor_sync(full, best_cost, o);
        const uint32_t ot = __shfl_xor_sync(full, best_t, o), on = __shfl_xor_sync(full, best_nl, o);
        if (ot != 0xffffffffu && (best_t == 0xffffffffu || oc < best_cost || (oc == best_cost && ot < best_t))) { best_cost = oc; best_t = ot; best_nl = on; }
    }
    if (lane == 0) {
        SahSplit r;
        if (best_t == 0xffffffffu || force_median) { r.axis = -1; r.bin = 0; r.clo = 0.f; r.scale = 0.f; r.n_left = c / 2u; }
        else {
            r.axis = (int)(best_t / (SAH_BINS - 1)); r.bin = (int)(best_t % (SAH_BINS - 1)) + 1;
            r.clo = r.axis == 0 ? clo[0] : r.axis == 1 ? clo[1] : clo[2];
            r.scale = r.axis == 0 ? scale[0] : r.axis == 1 ? scale[1] : scale[2];
            r.n_left = best_nl;
        }
        *out = r;
    }
}
// which side primitive p (position i of the node's range) goes to
template <class P>
__device__ __forceinline__ bool sah_goes_left(const P &prim, uint32_t p, uint32_t i, const SahSplit &sp)
{
    if (sp.axis < 0) return i < sp.n_left;
    float ce[3];
    prim.centroid(p, ce);
    const float cx = sp.axis == 0 ? ce[0] : sp.axis == 1 ? ce[1] : ce[2];
    return sah_bin(cx, sp.clo, sp.scale) < sp.bin;
}
// the node's two child records and the next level's descriptors of its inner children
__device__ __forceinline__ void sah_write_node(float4 *nodes, uint32_t id, uint32_t q, uint32_t next_base, uint32_t nl, uint32_t c, uint32_t f,
                                               const uint32_t *idx_out, float (*blo)[3], float (*bhi)[3], uint32_t *next_first, uint32_t *next_count)
{
    float4 *node = nodes + 4 * (size_t)id;
    const uint32_t cnt[2] = {nl, c - nl}, fst[2] = {f, f + nl};
    for (int s = 0; s < 2; ++s) {
        if (cnt[s] == 1u) write_child(node, s, true, (int)idx_out[fst[s]], blo[s], bhi[s]);
        else {
            write_child(node, s, false, (int)(next_base + q), blo[s], bhi[s]);
            next_first[q] = fst[s]; next_count[q] = cnt[s];
            ++q;
        }
    }
}

template <class P>
__global__ void __launch_bounds__(32 * SAH_WARPS) k_sah_split(const P prim, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ first,
                                                               const uint32_t *__restrict__ count, uint32_t m, int force_median,
                                                               SahSplit *__restrict__ split)
{
    __shared__ int s_bins[SAH_WARPS][3][SAH_BINS][7];       // per warp, axis, bin: box as 6 ordered ints, count
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x * SAH_WARPS + warp;
    if (j >= m) return;                                      // whole warps leave: only __syncwarp below
    const uint32_t f = first[j], c = count[j];
    const float INF = __int_as_float(0x7f800000);
    float clo[3] = {INF, INF, INF}, chi[3] = {-INF, -INF, -INF};
    for (uint32_t i = lane; i < c; i += 32u) {
        float ce[3];
        prim.centroid(idx[f + i], ce);
#pragma unroll
        for (int k = 0; k < 3; ++k) { clo[k] = fminf(clo[k], ce[k]); chi[k] = fmaxf(chi[k], ce[k]); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            clo[k] = fminf(clo[k], __shfl_xor_sync(full, clo[k], o));
            chi[k] = fmaxf(chi[k], __shfl_xor_sync(full, chi[k], o));
        }
    float scale[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) scale[k] = chi[k] > clo[k] ? (float)SAH_BINS / (chi[k] - clo[k]) : 0.0f;
    int (*bins)[SAH_BINS][7] = s_bins[warp];
    for (uint32_t t = lane; t < 3u * SAH_BINS; t += 32u) {
        int *b = bins[t / SAH_BINS][t % SAH_BINS];
        b[0] = b[1] = b[2] = f2ord(INF); b[3] = b[4] = b[5] = f2ord(-INF); b[6] = 0;
    }
    __syncwarp();
    for (uint32_t i = lane; i < c; i += 32u) {
        const uint32_t p = idx[f + i];
        float ce[3], lo[3], hi[3];
        prim.centroid(p, ce);
        prim.box(p, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!(scale[k] > 0.0f)) continue;
            int *b = bins[k][sah_bin(ce[k], clo[k], scale[k])];
#pragma unroll
            for (int a = 0; a < 3; ++a) { atomicMin(b + a, f2ord(lo[a])); atomicMax(b + 3 + a, f2ord(hi[a])); }
            atomicAdd(b + 6, 1);
        }
    }
    __syncwarp();
    sah_pick_plane(bins, clo, scale, c, force_median, lane, split + j);
}

__device__ __forceinline__ uint32_t sah_inner_children(const SahSplit &s, uint32_t c) { return (s.n_left >= 2u ? 1u : 0u) + (c - s.n_left >= 2u ? 1u : 0u); }
__global__ void __launch_bounds__(1024) k_sah_scan(const SahSplit *__restrict__ split, const uint32_t *__restrict__ count, uint32_t m,
                                                    uint32_t *__restrict__ off, uint32_t *__restrict__ m_next)
{
    __shared__ uint32_t s_sum[1024];
    const uint32_t chunk = (m + 1023u) / 1024u, b = threadIdx.x * chunk, e = min(b + chunk, m);
    uint32_t sum = 0;
    for (uint32_t j = b; j < e; ++j) sum += sah_inner_children(split[j], count[j]);
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? s_sum[threadIdx.x - o] : 0u;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_sum[threadIdx.x] - sum;
    for (uint32_t j = b; j < e; ++j) { off[j] = run; run += sah_inner_children(split[j], count[j]); }
    if (threadIdx.x == 1023u) *m_next = s_sum[1023];
}

template <class P>
__global__ void __launch_bounds__(32 * SAH_WARPS) k_sah_partition(const P prim, const uint32_t *__restrict__ idx_in, uint32_t *__restrict__ idx_out,
                                                                   const uint32_t *__restrict__ first, const uint32_t *__restrict__ count, uint32_t m,
                                                                   const SahSplit *__restrict__ split, const uint32_t *__restrict__ off,
                                                                   uint32_t base, uint32_t next_base, uint32_t *__restrict__ next_first,
                                                                   uint32_t *__restrict__ next_count, float4 *__restrict__ nodes)
{
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x * SAH_WARPS + warp;
    if (j >= m) return;
    const uint32_t f = first[j], c = count[j];
    const SahSplit sp = split[j];
    const uint32_t nl = sp.n_left;
    const float INF = __int_as_float(0x7f800000);
    float blo[2][3] = {{INF, INF, INF}, {INF, INF, INF}}, bhi[2][3] = {{-INF, -INF, -INF}, {-INF, -INF, -INF}};
    uint32_t wl = 0, wr = 0;
    for (uint32_t base_i = 0; base_i < c; base_i += 32u) {
        const uint32_t i = base_i + lane;
        const bool valid = i < c;
        uint32_t p = 0;
        bool left = false;
        if (valid) {
            p = idx_in[f + i];
            left = sah_goes_left(prim, p, i, sp);
            float lo[3], hi[3];
            prim.box(p, lo, hi);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (left) { blo[0][a] = fminf(blo[0][a], lo[a]); bhi[0][a] = fmaxf(bhi[0][a], hi[a]); }
                else      { blo[1][a] = fminf(blo[1][a], lo[a]); bhi[1][a] = fmaxf(bhi[1][a], hi[a]); }
            }
        }
        const unsigned ml = __ballot_sync(full, valid && left), mr = __ballot_sync(full, valid && !left);
        const unsigned lt = (1u << lane) - 1u;
        if (valid) idx_out[left ? f + wl + (uint32_t)__popc(ml & lt) : f + nl + wr + (uint32_t)__popc(mr & lt)] = p;
        wl += (uint32_t)__popc(ml); wr += (uint32_t)__popc(mr);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                blo[s][a] = fminf(blo[s][a], __shfl_xor_sync(full, blo[s][a], o));
                bhi[s][a] = fmaxf(bhi[s][a], __shfl_xor_sync(full, bhi[s][a], o));
            }
    __syncwarp();
    if (lane == 0) sah_write_node(nodes, base + j, off[j], next_base, nl, c, f, idx_out, blo, bhi, next_first, next_count);
}

// The same two steps with one BLOCK per node, for the few large nodes of the top levels (one warp per node would leave
// the GPU idle there).  Bins and boxes are order-independent, the partition is stable: the tree is bit-identical.
template <class P>
__global__ void __launch_bounds__(32 * SAH_WARPS) k_sah_split_block(const P prim, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ first,
                                                                     const uint32_t *__restrict__ count, uint32_t m, int force_median,
                                                                     SahSplit *__restrict__ split)
{
    __shared__ int s_bins[3][SAH_BINS][7];
    __shared__ float s_red[SAH_WARPS][6];
    __shared__ float s_c[6];
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x;
    const uint32_t f = first[j], c = count[j];
    const float INF = __int_as_float(0x7f800000);
    float clo[3] = {INF, INF, INF}, chi[3] = {-INF, -INF, -INF};
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        float ce[3];
        prim.centroid(idx[f + i], ce);
#pragma unroll
        for (int k = 0; k < 3; ++k) { clo[k] = fminf(clo[k], ce[k]); chi[k] = fmaxf(chi[k], ce[k]); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            clo[k] = fminf(clo[k], __shfl_xor_sync(full, clo[k], o));
            chi[k] = fmaxf(chi[k], __shfl_xor_sync(full, chi[k], o));
        }
    if (lane == 0) for (int k = 0; k < 3; ++k) { s_red[warp][k] = clo[k]; s_red[warp][3 + k] = chi[k]; }
    for (uint32_t t = threadIdx.x; t < 3u * SAH_BINS; t += blockDim.x) {
        int *b = s_bins[t / SAH_BINS][t % SAH_BINS];
        b[0] = b[1] = b[2] = f2ord(INF); b[3] = b[4] = b[5] = f2ord(-INF); b[6] = 0;
    }
    __syncthreads();
    if (threadIdx.x < 6u) {
        float v = s_red[0][threadIdx.x];
        for (int w = 1; w < SAH_WARPS; ++w) v = threadIdx.x < 3u ? fminf(v, s_red[w][threadIdx.x]) : fmaxf(v, s_red[w][threadIdx.x]);
        s_c[threadIdx.x] = v;
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { clo[k] = s_c[k]; chi[k] = s_c[3 + k]; scale[k] = chi[k] > clo[k] ? (float)SAH_BINS / (chi[k] - clo[k]) : 0.0f; }
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        const uint32_t p = idx[f + i];
        float ce[3], lo[3], hi[3];
        prim.centroid(p, ce);
        prim.box(p, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!(scale[k] > 0.0f)) continue;
            int *b = s_bins[k][sah_bin(ce[k], clo[k], scale[k])];
#pragma unroll
            for (int a = 0; a < 3; ++a) { atomicMin(b + a, f2ord(lo[a])); atomicMax(b + 3 + a, f2ord(hi[a])); }
            atomicAdd(b + 6, 1);
        }
    }
    __syncthreads();
    if (warp != 0) return;
    sah_pick_plane(s_bins, clo, scale, c, force_median, lane, split + j);
}

template <class P>
__global__ void __launch_bounds__(32 * SAH_WARPS) k_sah_partition_block(const P prim, const uint32_t *__restrict__ idx_in, uint32_t *__restrict__ idx_out,
                                                                         const uint32_t *__restrict__ first, const uint32_t *__restrict__ count, uint32_t m,
                                                                         const SahSplit *__restrict__ split, const uint32_t *__restrict__ off,
                                                                         uint32_t base, uint32_t next_base, uint32_t *__restrict__ next_first,
                                                                         uint32_t *__restrict__ next_count, float4 *__restrict__ nodes)
{
    __shared__ uint32_t s_cnt[SAH_WARPS][2];
    __shared__ float s_box[SAH_WARPS][12];
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x;
    const uint32_t f = first[j], c = count[j];
    const SahSplit sp = split[j];
    const uint32_t nl = sp.n_left;
    const float INF = __int_as_float(0x7f800000);
    float blo[2][3] = {{INF, INF, INF}, {INF, INF, INF}}, bhi[2][3] = {{-INF, -INF, -INF}, {-INF, -INF, -INF}};
    uint32_t wl = 0, wr = 0;
    for (uint32_t base_i = 0; base_i < c; base_i += blockDim.x) {
        const uint32_t i = base_i + threadIdx.x;
        const bool valid = i < c;
        uint32_t p = 0;
        bool left = false;
        if (valid) {
            p = idx_in[f + i];
            left = sah_goes_left(prim, p, i, sp);
            float lo[3], hi[3];
            prim.box(p, lo, hi);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (left) { blo[0][a] = fminf(blo[0][a], lo[a]); bhi[0][a] = fmaxf(bhi[0][a], hi[a]); }
                else      { blo[1][a] = fminf(blo[1][a], lo[a]); bhi[1][a] = fmaxf(bhi[1][a], hi[a]); }
            }
        }
        const unsigned ml = __ballot_sync(full, valid && left), mr = __ballot_sync(full, valid && !left);
        if (lane == 0) { s_cnt[warp][0] = (uint32_t)__popc(ml); s_cnt[warp][1] = (uint32_t)__popc(mr); }
        __syncthreads();
        uint32_t pl = 0, pr = 0, tl = 0, tr = 0;
#pragma unroll
        for (int w = 0; w < SAH_WARPS; ++w) {
            if (w < (int)warp) { pl += s_cnt[w][0]; pr += s_cnt[w][1]; }
            tl += s_cnt[w][0]; tr += s_cnt[w][1];
        }
        const unsigned lt = (1u << lane) - 1u;
        if (valid) idx_out[left ? f + wl + pl + (uint32_t)__popc(ml & lt) : f + nl + wr + pr + (uint32_t)__popc(mr & lt)] = p;
        wl += tl; wr += tr;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                blo[s][a] = fminf(blo[s][a], __shfl_xor_sync(full, blo[s][a], o));
                bhi[s][a] = fmaxf(bhi[s][a], __shfl_xor_sync(full, bhi[s][a], o));
            }
    if (lane == 0)
        for (int s = 0; s < 2; ++s)
            for (int a = 0; a < 3; ++a) { s_box[warp][s * 6 + a] = blo[s][a]; s_box[warp][s * 6 + 3 + a] = bhi[s][a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SAH_WARPS; ++w)
            for (int s = 0; s < 2; ++s)
                for (int a = 0; a < 3; ++a) { blo[s][a] = fminf(blo[s][a], s_box[w][s * 6 + a]); bhi[s][a] = fmaxf(bhi[s][a], s_box[w][s * 6 + 3 + a]); }
        sah_write_node(nodes, base + j, off[j], next_base, nl, c, f, idx_out, blo, bhi, next_first, next_count);
    }
}
__global__ void __launch_bounds__(256) k_iota(uint32_t *p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// `nodes`: n - 1 exact 64-byte nodes (caller-allocated), root = node 0.  n >= 2.
template <class P>
static cudaError_t build_sah_tree(const P prim, uint32_t n, float4 *nodes, int *depth_out, uint32_t *launches_out, cudaStream_t st)
{
    cudaError_t err = cudaSuccess;
    uint32_t *idx[2] = {nullptr, nullptr}, *first[2] = {nullptr, nullptr}, *count[2] = {nullptr, nullptr}, *off = nullptr, *d_mnext = nullptr;
    SahSplit *split = nullptr;
    uint32_t launches = 0, m = 1, base = 0;
    int level = 0, cur = 0;
    const uint32_t root[2] = {0u, n};
    for (int k = 0; k < 2; ++k) {
        CK(cudaMalloc(&idx[k], (size_t)n * sizeof(uint32_t)));
        CK(cudaMalloc(&first[k], (size_t)n * sizeof(uint32_t)));
        CK(cudaMalloc(&count[k], (size_t)n * sizeof(uint32_t)));
    }
    CK(cudaMalloc(&off, (size_t)n * sizeof(uint32_t)));
    CK(cudaMalloc(&split, (size_t)n * sizeof(SahSplit)));
    CK(cudaMalloc(&d_mnext, sizeof(uint32_t)));
    k_iota<<<(n + 255u) / 256u, 256, 0, st>>>(idx[0], n); ++launches;
    CK(cudaMemcpyAsync(first[0], &root[0], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(count[0], &root[1], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    while (m > 0) {
        const unsigned grid = (m + SAH_WARPS - 1u) / SAH_WARPS;
        uint32_t m_next = 0;
        // a level whose nodes hold 1024 primitives or more on average (the top levels) gives every node a block
        const bool by_block = VKRT_SAH_BLOCK_LEVELS && (uint64_t)m * 1024u <= n;
        const int fm = level >= SAH_MEDIAN_FROM_LEVEL ? 1 : 0;
        if (by_block) k_sah_split_block<P><<<m, 32 * SAH_WARPS, 0, st>>>(prim, idx[cur], first[cur], count[cur], m, fm, split);
        else k_sah_split<P><<<grid, 32 * SAH_WARPS, 0, st>>>(prim, idx[cur], first[cur], count[cur], m, fm, split);
        ++launches;
        k_sah_scan<<<1, 1024, 0, st>>>(split, count[cur], m, off, d_mnext); ++launches;
        if (by_block) k_sah_partition_block<P><<<m, 32 * SAH_WARPS, 0, st>>>(prim, idx[cur], idx[cur ^ 1], first[cur], count[cur], m, split, off, base, base + m,
                                                                             first[cur ^ 1], count[cur ^ 1], nodes);
        else k_sah_partition<P><<<grid, 32 * SAH_WARPS, 0, st>>>(prim, idx[cur], idx[cur ^ 1], first[cur], count[cur], m, split, off, base, base + m,
                                                                 first[cur ^ 1], count[cur ^ 1], nodes);
        ++launches;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&m_next, d_mnext, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        base += m; m = m_next; cur ^= 1; ++level;
        if (level > 120) { err = cudaErrorUnknown; goto done; }      // cannot happen: median splits from level 48 on
    }
    if (base != n - 1u) { err = cudaErrorUnknown; goto done; }
    *depth_out = level;
    *launches_out += launches;
done:
    for (int k = 0; k < 2; ++k) { cudaFree(idx[k]); cudaFree(first[k]); cudaFree(count[k]); }
    cudaFree(off); cudaFree(split); cudaFree(d_mnext);
    return err;
}

cudaError_t build_lbvh(const float4 *d_spheres, uint32_t n, BvhBuild &out, cudaStream_t st)
{
    cudaError_t err = cudaSuccess;
    if (out.nodes) { cudaFree(out.nodes); out.nodes = nullptr; }
    if (out.nodes4) { cudaFree(out.nodes4); out.nodes4 = nullptr; }
    if (out.qnodes) { cudaFree(out.qnodes); out.qnodes = nullptr; }
    if (out.tnodes) { cudaFree(out.tnodes); out.tnodes = nullptr; }
    float *grid = nullptr;
    out.n_nodes = 0; out.build_ms = 0.f; out.launches = 0; out.depth = 0; out.tdepth = 0; out.sah_ms = 0.f;
    if (n == 0) return cudaSuccess;

    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    const uint32_t n_inner = n > 1 ? n - 1 : 1;
    uint32_t launches = 0;

    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMalloc(&out.nodes, (size_t)n_inner * 64));
    CK(cudaMalloc(&out.nodes4, (size_t)n_inner * 128));
    CK(cudaMalloc(&out.qnodes, (size_t)n_inner * 32));
    CK(cudaMalloc(&grid, 6 * sizeof(float)));
    CK(cudaEventRecord(e0, st));
    CK(build_tree(PrimSpheres{d_spheres}, n, out.nodes, &out.depth, &launches, st));
    k_collapse4<<<(n_inner + 255u) / 256u, 256, 0, st>>>(out.nodes, n_inner, out.nodes4); ++launches;
    CK(cudaGetLastError());
    // the traversal tree of the 32-byte coded nodes: the SAH tree for scenes of VKRT_SAH_MIN_PRIMS primitives or more
    // (below that the LBVH itself; rule S's answer does not depend on the choice)
    {
        const float4 *src = out.nodes;
        out.tdepth = out.depth;
        if (VKRT_SAH_TREE && n >= (uint32_t)VKRT_SAH_MIN_PRIMS) {
            CK(cudaEventCreate(&e2));
            CK(cudaMalloc(&out.tnodes, (size_t)n_inner * 64));
            CK(cudaEventRecord(e2, st));
            CK(build_sah_tree(PrimSpheres{d_spheres}, n, out.tnodes, &out.tdepth, &launches, st));
            src = out.tnodes;
        }
        k_qgrid<<<1, 32, 0, st>>>(src, grid); ++launches;
        k_quantize<<<(n_inner + 255u) / 256u, 256, 0, st>>>(src, n_inner, grid, out.qnodes); ++launches;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, st));
    CK(cudaMemcpyAsync(out.qgrid, grid, 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&out.build_ms, e0, e1));
    if (e2) CK(cudaEventElapsedTime(&out.sah_ms, e2, e1));
    out.n_nodes = n_inner;
    out.launches = launches;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    cudaFree(grid);
    if (err != cudaSuccess) {
        cudaFree(out.nodes); cudaFree(out.nodes4); cudaFree(out.qnodes); cudaFree(out.tnodes);
        out.nodes = nullptr; out.nodes4 = nullptr; out.qnodes = nullptr; out.tnodes = nullptr;
    }
    return err;
}

// the triangles' own hierarchy (rule T): the same builder over the padded triangle boxes; exact 64-byte nodes only
cudaError_t build_tri_lbvh(const float4 *d_tris, uint32_t n, TriBvhBuild &out, cudaStream_t st)
{
    cudaError_t err = cudaSuccess;
    if (out.nodes) { cudaFree(out.nodes); out.nodes = nullptr; }
    out.n_nodes = 0; out.depth = 0; out.build_ms = 0.f; out.launches = 0;
    if (n == 0) return cudaSuccess;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const uint32_t n_inner = n > 1 ? n - 1 : 1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMalloc(&out.nodes, (size_t)n_inner * 64));
    CK(cudaEventRecord(e0, st));
    CK(build_tree(PrimTris{d_tris}, n, out.nodes, &out.depth, &out.launches, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&out.build_ms, e0, e1));
    out.n_nodes = n_inner;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (err != cudaSuccess) { cudaFree(out.nodes); out.nodes = nullptr; }
    return err;
}

} // namespace vkrt
