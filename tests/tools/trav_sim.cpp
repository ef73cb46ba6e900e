// Design experiment (test infrastructure, CPU only): counts node visits / box tests / leaf tests of candidate
// traversal schemes on the oracle's LBVH, for synthetic secondary and shadow rays of the cfg4 scene.
//   usage: trav_sim spheres.bin nodes.bin n_rays [sah | b<bins> | p<radius>]
// The optional fourth argument replaces the LBVH by another hierarchy over the same spheres before the walks are
// counted: `sah` = top-down full-sweep SAH, `p16` = PLOC (locally-ordered clustering along the Morton order, search
// radius 16).  Any hierarchy with exact union boxes gives rule S's answer; the question is how many visits it costs.
// spheres.bin: n x {cx,cy,cz,r} float32;  nodes.bin: (n-1) x 16 float32 (orc_scene_read_bvh layout)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

struct V3 { float x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3 norm(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct Sphere { float cx, cy, cz, r; };
struct Child { float lo[3], hi[3]; int index, kind; };
struct Node { Child c[2]; };

struct Box { float lo[3], hi[3]; };
static inline float area(const Box &b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
struct WChild { Box b; int ref; };          // ref >= 0 wide node, < 0 ~sphere
struct WNode { std::vector<WChild> c; };

static std::vector<Sphere> sph;
static std::vector<Node> nodes;

struct Ray { V3 o, d, inv, oinv; float B; bool any; };
static inline float sinv(float d) { return 1.0f / (std::fabs(d) > 1e-20f ? d : std::copysign(1e-20f, d)); }
static inline void setup(Ray &r) { r.inv = {sinv(r.d.x), sinv(r.d.y), sinv(r.d.z)}; r.oinv = {r.o.x * r.inv.x, r.o.y * r.inv.y, r.o.z * r.inv.z}; }
static inline bool slab(const Ray &r, const float *lo, const float *hi, float &tn, float &tf)
{
    const float t0x = lo[0] * r.inv.x - r.oinv.x, t1x = hi[0] * r.inv.x - r.oinv.x;
    const float t0y = lo[1] * r.inv.y - r.oinv.y, t1y = hi[1] * r.inv.y - r.oinv.y;
    const float t0z = lo[2] * r.inv.z - r.oinv.z, t1z = hi[2] * r.inv.z - r.oinv.z;
    tn = std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::min(t0z, t1z));
    tf = std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::max(t0z, t1z));
    return tn <= tf && tf >= 0.0f;
}
static inline float isect(const Ray &r, const Sphere &s)
{
    const V3 oc = r.o - V3{s.cx, s.cy, s.cz};
    const float b = 2.0f * dot(oc, r.d), c = dot(oc, oc) - s.r * s.r, h = b * b - 4.0f * c;
    if (h < 0.0f) return -1.0f;
    return (-b - std::sqrt(h)) * 0.5f;
}
struct Best { float t; int idx; };
static inline void leaf(const Ray &r, int si, Best &best, uint64_t &leaves)
{
    const Sphere &s = sph[si];
    const float rp = s.r * 1.001f + 0.001f;
    const float lo[3] = {s.cx - rp, s.cy - rp, s.cz - rp}, hi[3] = {s.cx + rp, s.cy + rp, s.cz + rp};
    float tn, tf;
    if (!slab(r, lo, hi, tn, tf) || !(tn <= best.t)) return;
    ++leaves;
    const float t = isect(r, s);
    if (!(t > 1e-3f) || !(tn <= t)) return;
    if (best.idx < 0) { if (t < r.B) { best.t = t; best.idx = si; } }
    else if (t < best.t || (t == best.t && si < best.idx)) { best.t = t; best.idx = si; }
}


// ---- tree-quality experiments: other hierarchies over the same spheres (exact padded leaf boxes, exact unions) ----
static Box leafbox(int si)
{
    const Sphere &s = sph[si];
    const float rp = s.r * 1.001f + 0.001f;
    return Box{{s.cx - rp, s.cy - rp, s.cz - rp}, {s.cx + rp, s.cy + rp, s.cz + rp}};
}
static Box uni(const Box &a, const Box &b)
{
    Box r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
    return r;
}
static Node make_node(const Box &b0, int r0, const Box &b1, int r1)      // ref: >= 0 inner node, < 0 ~sphere
{
    Node nd;
    memcpy(nd.c[0].lo, b0.lo, 12); memcpy(nd.c[0].hi, b0.hi, 12); nd.c[0].kind = r0 < 0; nd.c[0].index = r0 < 0 ? ~r0 : r0;
    memcpy(nd.c[1].lo, b1.lo, 12); memcpy(nd.c[1].hi, b1.hi, 12); nd.c[1].kind = r1 < 0; nd.c[1].index = r1 < 0 ? ~r1 : r1;
    return nd;
}
// (1) top-down, full-sweep surface-area heuristic over the three centroid orders
static std::vector<Node> sahn;
static int sah_build(std::vector<int> &ids, int b, int e, Box &out)
{
    if (e - b == 1) { out = leafbox(ids[b]); return ~ids[b]; }
    const int n = e - b;
    float bestc = 1e30f; int bax = 0, bsplit = b + n / 2;
    std::vector<float> la(n);
    auto by_axis = [&](int ax) {
        std::sort(ids.begin() + b, ids.begin() + e, [&](int x, int y) {
            const float cx = (&sph[x].cx)[ax], cy = (&sph[y].cx)[ax];
            return cx < cy || (cx == cy && x < y);
        });
    };
    for (int ax = 0; ax < 3; ++ax) {
        by_axis(ax);
        Box acc = leafbox(ids[b]);
        for (int i = 0; i < n - 1; ++i) { acc = uni(acc, leafbox(ids[b + i])); la[i] = area(acc) * (i + 1); }
        acc = leafbox(ids[e - 1]);
        for (int i = n - 1; i >= 1; --i) {
            acc = uni(acc, leafbox(ids[b + i]));
            const float c = la[i - 1] + area(acc) * (n - i);
            if (c < bestc) { bestc = c; bax = ax; bsplit = b + i; }
        }
    }
    by_axis(bax);
    const int id = (int)sahn.size();
    sahn.emplace_back();
    Box b0, b1;
    const int r0 = sah_build(ids, b, bsplit, b0), r1 = sah_build(ids, bsplit, e, b1);
    sahn[id] = make_node(b0, r0, b1, r1);
    out = uni(b0, b1);
    return id;
}
// (1b) top-down binned SAH (NB bins per axis over the centroid bounds), what a level-synchronous device builder does
static int binned_build(std::vector<int> &ids, int b, int e, Box &out, int NB)
{
    if (e - b == 1) { out = leafbox(ids[b]); return ~ids[b]; }
    const int n = e - b;
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = b; i < e; ++i) for (int k = 0; k < 3; ++k) { const float c = (&sph[ids[i]].cx)[k]; clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c); }
    float bestc = 1e30f; int bax = -1, bbin = 0;
    for (int ax = 0; ax < 3; ++ax) {
        if (!(chi[ax] > clo[ax])) continue;
        std::vector<Box> bb(NB, Box{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}});
        std::vector<int> bc(NB, 0);
        const float sc = NB / (chi[ax] - clo[ax]);
        for (int i = b; i < e; ++i) {
            int k = (int)(((&sph[ids[i]].cx)[ax] - clo[ax]) * sc); k = std::min(std::max(k, 0), NB - 1);
            bb[k] = uni(bb[k], leafbox(ids[i])); ++bc[k];
        }
        std::vector<float> la(NB); std::vector<int> lc(NB);
        Box acc{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}}; int cnt = 0;
        for (int k = 0; k < NB; ++k) { if (bc[k]) acc = uni(acc, bb[k]); cnt += bc[k]; la[k] = cnt ? area(acc) * cnt : 0.f; lc[k] = cnt; }
        acc = Box{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}}; cnt = 0;
        for (int k = NB - 1; k >= 1; --k) {
            if (bc[k]) acc = uni(acc, bb[k]); cnt += bc[k];
            if (cnt == 0 || lc[k - 1] == 0) continue;
            const float c = la[k - 1] + area(acc) * cnt;
            if (c < bestc) { bestc = c; bax = ax; bbin = k; }
        }
    }
    int mid;
    if (bax < 0) mid = b + n / 2;            // all centroids equal: split the index range
    else {
        const float sc = NB / (chi[bax] - clo[bax]);
        mid = (int)(std::stable_partition(ids.begin() + b, ids.begin() + e, [&](int x) {
            int k = (int)(((&sph[x].cx)[bax] - clo[bax]) * sc); k = std::min(std::max(k, 0), NB - 1); return k < bbin; }) - ids.begin());
    }
    const int id = (int)sahn.size();
    sahn.emplace_back();
    Box b0, b1;
    const int r0 = binned_build(ids, b, mid, b0, NB), r1 = binned_build(ids, mid, e, b1, NB);
    sahn[id] = make_node(b0, r0, b1, r1);
    out = uni(b0, b1);
    return id;
}
// (1c) tree rotations on top of any tree in `t` (Kensler 2008): swap a child with a grandchild under the sibling when
// that lowers the sum of box areas; a few bottom-up passes
static Box child_box(const Node &n, int j) { Box b; memcpy(b.lo, n.c[j].lo, 12); memcpy(b.hi, n.c[j].hi, 12); return b; }
static void set_child(Node &n, int j, const Box &b, int kind, int index) { memcpy(n.c[j].lo, b.lo, 12); memcpy(n.c[j].hi, b.hi, 12); n.c[j].kind = kind; n.c[j].index = index; }
static int rotate_pass(std::vector<Node> &t)
{
    int done = 0;
    // children have larger ids than parents in the trees built here (sahn: preorder), so a reverse sweep is bottom-up
    for (int v = (int)t.size() - 1; v >= 0; --v) {
        for (int side = 0; side < 2; ++side) {
            Node &n = t[v];
            if (n.c[side].kind) continue;                    // the child to open must be an inner node
            const int w = n.c[side].index;
            const int other = 1 - side;
            float best = area(child_box(n, side));           // only the opened child's box changes
            int pick = -1;
            for (int g = 0; g < 2; ++g) {                    // swap n.c[other] with t[w].c[g]
                const Box nb = uni(child_box(n, other), child_box(t[w], 1 - g));
                if (area(nb) < best) { best = area(nb); pick = g; }
            }
            if (pick < 0) continue;
            Child a = n.c[other], b = t[w].c[pick];
            n.c[other] = b; t[w].c[pick] = a;
            const Box nb = uni(child_box(t[w], 0), child_box(t[w], 1));
            set_child(n, side, nb, 0, w);
            ++done;
        }
    }
    return done;
}
// (2) PLOC: clusters in the LBVH's leaf (Morton) order; every cluster looks `R` neighbours to each side for the partner
// with the smallest merged box, mutual choices merge, repeat until one cluster is left
static void leaf_order(int n, std::vector<int> &out)
{
    for (int j = 0; j < 2; ++j) { const Child &c = nodes[n].c[j]; if (c.kind) out.push_back(c.index); else leaf_order(c.index, out); }
}
static void ploc_build(int R)
{
    std::vector<int> order;
    leaf_order(0, order);
    struct Cl { Box b; int ref; };
    std::vector<Cl> cur;
    for (int si : order) cur.push_back({leafbox(si), ~si});
    std::vector<Node> out;
    while (cur.size() > 1) {
        const int n = (int)cur.size();
        std::vector<int> nn(n);
        for (int i = 0; i < n; ++i) {
            float best = 1e30f; int bj = -1;
            for (int j = std::max(0, i - R); j <= std::min(n - 1, i + R); ++j) {
                if (j == i) continue;
                const float a = area(uni(cur[i].b, cur[j].b));
                if (a < best) { best = a; bj = j; }
            }
            nn[i] = bj;
        }
        std::vector<Cl> nxt;
        for (int i = 0; i < n; ++i) {
            const int j = nn[i];
            if (nn[j] != i) { nxt.push_back(cur[i]); continue; }
            if (i > j) continue;
            out.push_back(make_node(cur[i].b, cur[i].ref, cur[j].b, cur[j].ref));
            nxt.push_back({uni(cur[i].b, cur[j].b), (int)out.size() - 1});
        }
        cur.swap(nxt);
    }
    const int N = (int)out.size();                 // the root was created last: renumber so that it is node 0
    std::vector<Node> ren(N);
    for (int i = 0; i < N; ++i) {
        Node nd = out[i];
        for (int j = 0; j < 2; ++j) if (!nd.c[j].kind) nd.c[j].index = N - 1 - nd.c[j].index;
        ren[N - 1 - i] = nd;
    }
    nodes = ren;
}
static double sum_child_area(const std::vector<Node> &t)
{
    double s = 0;
    for (const Node &n : t)
        for (int j = 0; j < 2; ++j) { Box b; memcpy(b.lo, n.c[j].lo, 12); memcpy(b.hi, n.c[j].hi, 12); s += area(b); }
    return s;
}

struct Cnt { uint64_t visits = 0, boxes = 0, leaves = 0, stale = 0, pushes = 0, maxsp = 0, popcull = 0; };

// (a) the shipping binary walk; cull_pop: the entry's own tn travels with it and is re-checked at the pop
static Best trav_binary(const Ray &r, bool cull_pop, int trunc_bits, Cnt &c)
{
    Best best{r.B, -1};
    struct E { int ref; float tn; };
    E stack[256]; int sp = 0;
    int node = 0;
    for (;;) {
        if (node < 0) {
            leaf(r, ~node, best, c.leaves);
            if (r.any && best.idx >= 0) return best;
        } else {
            const Node &n = nodes[node];
            ++c.visits; c.boxes += 2;
            float tn0, tn1, tf;
            const bool h0 = slab(r, n.c[0].lo, n.c[0].hi, tn0, tf) && tn0 <= best.t;
            const bool h1 = slab(r, n.c[1].lo, n.c[1].hi, tn1, tf) && tn1 <= best.t;
            const int c0 = n.c[0].kind ? ~n.c[0].index : n.c[0].index, c1 = n.c[1].kind ? ~n.c[1].index : n.c[1].index;
            if (!h0 && !h1) ++c.stale;
            const bool take1 = h1 && (!h0 || tn1 < tn0);
            if (h0 && h1) { stack[sp++] = {take1 ? c0 : c1, take1 ? tn0 : tn1}; ++c.pushes; c.maxsp = std::max<uint64_t>(c.maxsp, sp); }
            if (h0 || h1) { node = take1 ? c1 : c0; continue; }
        }
        for (;;) {
            if (sp == 0) return best;
            const E e = stack[--sp];
            if (cull_pop) {
                float tn = e.tn;
                if (trunc_bits) { uint32_t u; memcpy(&u, &tn, 4); if (tn > 0) { u &= ~((1u << (23 - trunc_bits)) - 1u); memcpy(&tn, &u, 4); } else tn = 0.f; }
                if (tn > best.t) { ++c.popcull; continue; }
            }
            node = e.ref; break;
        }
    }
}

// (c) k-wide collapse of the binary tree: open the inner child of largest area until k slots are used
static std::vector<WNode> wide;
static int build_wide(int bnode, int k)
{
    const int id = (int)wide.size();
    wide.emplace_back();
    std::vector<std::pair<Box, int>> slots;     // ref: >= 0 binary inner node, < 0 ~sphere
    for (int j = 0; j < 2; ++j) {
        const Child &ch = nodes[bnode].c[j];
        Box b; memcpy(b.lo, ch.lo, 12); memcpy(b.hi, ch.hi, 12);
        slots.push_back({b, ch.kind ? ~ch.index : ch.index});
    }
    while ((int)slots.size() < k) {
        int pick = -1; float best = -1.f;
        for (int j = 0; j < (int)slots.size(); ++j) if (slots[j].second >= 0 && area(slots[j].first) > best) { best = area(slots[j].first); pick = j; }
        if (pick < 0) break;
        const int bn = slots[pick].second;
        slots.erase(slots.begin() + pick);
        for (int j = 0; j < 2; ++j) {
            const Child &ch = nodes[bn].c[j];
            Box b; memcpy(b.lo, ch.lo, 12); memcpy(b.hi, ch.hi, 12);
            slots.push_back({b, ch.kind ? ~ch.index : ch.index});
        }
    }
    std::vector<WChild> cs;
    for (auto &s : slots) cs.push_back({s.first, s.second});
    for (auto &ch : cs) if (ch.ref >= 0) ch.ref = build_wide(ch.ref, k);
    wide[id].c = cs;
    return id;
}
// mode 0: hits sorted by tn, nearest first; mode 1: stored order (no sort); cull_pop as above
static Best trav_wide(const Ray &r, int mode, bool cull_pop, Cnt &c)
{
    Best best{r.B, -1};
    struct E { int ref; float tn; };
    E stack[512]; int sp = 0;
    int node = 0;
    for (;;) {
        if (node < 0) {
            leaf(r, ~node, best, c.leaves);
            if (r.any && best.idx >= 0) return best;
        } else {
            const WNode &n = wide[node];
            ++c.visits; c.boxes += n.c.size();
            E hit[16]; int nh = 0;
            for (const WChild &ch : n.c) {
                float tn, tf;
                if (slab(r, ch.b.lo, ch.b.hi, tn, tf) && tn <= best.t) hit[nh++] = {ch.ref, tn};
            }
            if (nh == 0) ++c.stale;
            if (mode == 0) std::sort(hit, hit + nh, [](const E &a, const E &b) { return a.tn < b.tn; });
            for (int j = nh - 1; j >= 1; --j) { stack[sp++] = hit[j]; ++c.pushes; }
            c.maxsp = std::max<uint64_t>(c.maxsp, sp);
            if (nh) { node = hit[0].ref; continue; }
        }
        for (;;) {
            if (sp == 0) return best;
            const E e = stack[--sp];
            if (cull_pop && e.tn > best.t) { ++c.popcull; continue; }
            node = e.ref; break;
        }
    }
}

int main(int argc, char **argv)
{
    if (argc < 4) return 1;
    {
        FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); const long n = ftell(f) / 16; fseek(f, 0, SEEK_SET);
        sph.resize(n); if (fread(sph.data(), 16, n, f) != (size_t)n) return 2; fclose(f);
    }
    {
        FILE *f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); const long n = ftell(f) / 64; fseek(f, 0, SEEK_SET);
        std::vector<float> raw(n * 16); if (fread(raw.data(), 64, n, f) != (size_t)n) return 2; fclose(f);
        nodes.resize(n);
        for (long i = 0; i < n; ++i)
            for (int k = 0; k < 2; ++k) {
                const float *p = raw.data() + i * 16 + k * 8;
                Child &c = nodes[i].c[k];
                c.lo[0] = p[0]; c.lo[1] = p[1]; c.lo[2] = p[2]; c.hi[0] = p[3]; c.hi[1] = p[4]; c.hi[2] = p[5];
                memcpy(&c.index, p + 6, 4); memcpy(&c.kind, p + 7, 4);
            }
    }
    const int n_rays = atoi(argv[3]);
    printf("LBVH: %zu nodes, sum of child box areas %.4g\n", nodes.size(), sum_child_area(nodes));
    if (argc > 4 && argv[4][0] == 'p') {
        ploc_build(atoi(argv[4] + 1));
        printf("PLOC radius %d: %zu nodes, sum of child box areas %.4g\n", atoi(argv[4] + 1), nodes.size(), sum_child_area(nodes));
    } else if (argc > 4 && argv[4][0] == 'b') {
        std::vector<int> ids(sph.size());
        for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
        Box rb;
        binned_build(ids, 0, (int)ids.size(), rb, atoi(argv[4] + 1));
        nodes = sahn;
        printf("binned SAH, %d bins: %zu nodes, sum of child box areas %.4g\n", atoi(argv[4] + 1), nodes.size(), sum_child_area(nodes));
        if (argc > 5) for (int pass = 0; pass < atoi(argv[5]); ++pass) { const int r = rotate_pass(nodes); printf("rotation pass %d: %d rotations, sum of child box areas %.4g\n", pass, r, sum_child_area(nodes)); }
    } else if (argc > 4) {
        std::vector<int> ids(sph.size());
        for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
        Box rb;
        sah_build(ids, 0, (int)ids.size(), rb);
        nodes = sahn;
        printf("sweep SAH: %zu nodes, sum of child box areas %.4g\n", nodes.size(), sum_child_area(nodes));
    }
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<Ray> rays;
    const V3 light{0.f, 96.f, 0.f};
    for (int i = 0; i < n_rays; ++i) {
        // origin: a point on a random sphere (70 %) or on a wall / floor (30 %); bounce = cosine lobe about the normal
        V3 P, N;
        if (U(rng) < 0.7f) {
            const Sphere &s = sph[1 + (size_t)(U(rng) * (sph.size() - 1)) % (sph.size() - 1)];
            const float z = 2.f * U(rng) - 1.f, ph = 6.2831853f * U(rng), rr = std::sqrt(std::max(0.f, 1.f - z * z));
            N = {rr * std::cos(ph), z, rr * std::sin(ph)};
            P = V3{s.cx, s.cy, s.cz} + N * s.r;
        } else {
            const int w = (int)(U(rng) * 5.f) % 5;
            const float a = -60.f + 120.f * U(rng), b = -60.f + 120.f * U(rng), h = 128.f * U(rng);
            if (w == 0) { P = {a, 0.f, b}; N = {0, 1, 0}; }
            else if (w == 1) { P = {a, 128.f, b}; N = {0, -1, 0}; }
            else if (w == 2) { P = {-64.f, h, b}; N = {1, 0, 0}; }
            else if (w == 3) { P = {a, h, 64.f}; N = {0, 0, -1}; }
            else { P = {64.f, h, b}; N = {-1, 0, 0}; }
        }
        Ray r;
        r.o = P;
        if (i & 1) {      // shadow ray towards the light sphere
            const V3 L = light - P;
            const float dist = std::sqrt(dot(L, L));
            r.d = norm(L + V3{U(rng) - .5f, U(rng) - .5f, U(rng) - .5f} * 8.f);
            if (dot(r.d, N) <= 0.f) { --i; continue; }      // zero-term rays are not traced
            r.B = dist - 12.f + 1e-3f; r.any = true;
        } else {
            const V3 w = N, u = norm(cross(std::fabs(w.x) > .5f ? V3{0, 1, 0} : V3{1, 0, 0}, w)), v = cross(w, u);
            const float r2 = U(rng), ph = 6.2831853f * U(rng), s = std::sqrt(r2), cz = std::sqrt(1.f - r2);
            r.d = norm(u * (s * std::cos(ph)) + v * (s * std::sin(ph)) + w * cz);
            const int depth = 1 + (int)(U(rng) * U(rng) * 7.f);
            r.B = 3000.f / ((depth + 1.f) * (depth + 1.f)) + 1e-3f; r.any = false;
        }
        setup(r);
        rays.push_back(r);
    }
    auto report = [&](const char *name, const Cnt &c, int kind) {
        const double n = rays.size() / 2.0;
        printf("%-34s %s  visits %6.2f  boxes %6.2f  leaves %5.2f  stale %5.2f  pushes %5.2f  popcull %5.2f  maxsp %d\n", name,
               kind ? "shadow " : "nearest", c.visits / n, c.boxes / n, c.leaves / n, c.stale / n, c.pushes / n, c.popcull / n, (int)c.maxsp);
    };
    std::vector<Best> ref(rays.size());
    for (int kind = 0; kind < 2; ++kind) {
        Cnt c;
        for (size_t i = kind; i < rays.size(); i += 2) ref[i] = trav_binary(rays[i], false, 0, c);
        report("binary (shipping)", c, kind);
    }
    for (int kind = 0; kind < 2; ++kind) {
        Cnt c;
        for (size_t i = kind; i < rays.size(); i += 2) { const Best b = trav_binary(rays[i], true, 0, c); if (!rays[i].any && (b.idx != ref[i].idx)) printf("MISMATCH\n"); }
        report("binary + tn on stack", c, kind);
    }
    for (int kind = 0; kind < 2; ++kind) {
        Cnt c;
        for (size_t i = kind; i < rays.size(); i += 2) trav_binary(rays[i], true, 5, c);
        report("binary + 5-bit tn on stack", c, kind);
    }
    for (int k : {4, 8}) {
        wide.clear();
        build_wide(0, k);
        double fill = 0; for (auto &w : wide) fill += w.c.size();
        printf("-- %d-wide: %zu nodes, mean fill %.2f\n", k, wide.size(), fill / wide.size());
        for (int mode = 0; mode < 2; ++mode)
            for (int cp = 0; cp < 2; ++cp)
                for (int kind = 0; kind < 2; ++kind) {
                    Cnt c;
                    for (size_t i = kind; i < rays.size(); i += 2) { const Best b = trav_wide(rays[i], mode, cp, c); if (!rays[i].any && (b.idx != ref[i].idx)) printf("MISMATCH\n"); }
                    char nm[64]; snprintf(nm, sizeof nm, "%d-wide %s%s", k, mode ? "unsorted" : "sorted", cp ? " + tn on stack" : "");
                    report(nm, c, kind);
                }
    }
    return 0;
}
