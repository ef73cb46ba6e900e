"""GPU parity tests proper: libvkrt_cuda (through the C ABI) against the CPU oracle on identical
scene, camera and RNG seed.  Bar: primary hit ids bit-exact; linear radiance bit-exact where the
arithmetic contract promises it (this is stricter than BASELINE.json's 1e-3 / 99.9 % / RMSE 1e-4,
which is asserted as well)."""
import numpy as np
import pytest

from helpers import apply_scene, bits_equal, mismatch_report, radiance_parity

pytestmark = pytest.mark.gpu


def _render_gpu(V, scene, w, h, fd, spp, depth, integrator, bvh, seed=7, frame_index=0, variant=0, flags=0, default=None):
    r = V.Renderer(w, h, spp=spp, max_depth=depth, integrator=integrator, variant=variant,
                   flags=V.FLAG_HIT_IDS | flags)
    if default is not None:
        r.use_default_scene(default)
    else:
        r.set_scene(scene)
    if bvh:
        r.build_bvh()
    r.set_seed(seed)
    r.set_frame_index(frame_index)
    r.draw(fd)
    out = (r.read_accum(), r.read_hit_ids(), r.read_rgba8(), r.counters())
    return r, out


def test_whitted_default_scene_bit_exact(vk, oracle):
    """BASELINE config 1: Raytracer.comp scene, 640x480, whitted; fully deterministic."""
    V = vk
    w, h = 640, 480
    fd = V.default_frame_data(aspect_ratio=w / h)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, None, w, h, fd, 1, 2, V.INTEGRATOR_WHITTED, False, default=V.SCENE_RAYTRACER)
    sc = oracle.Scene().use_default(oracle.SCENE_RAYTRACER)
    oacc, oids, orgba, ocnt = sc.render(fd, w, h, spp=1, max_depth=2, integrator=oracle.WHITTED)
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert np.array_equal(rgba, orgba)
    assert (cnt.closest_rays, cnt.shadow_rays) == (ocnt.closest_rays, ocnt.shadow_rays)
    r.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_path_default_scene_bit_exact(vk, oracle, variant):
    """Tracer.comp scene at the reference's SAMPLES = DEPTH = 4, literal primitive order."""
    V = vk
    w, h = 320, 240
    fd = V.default_frame_data(aspect_ratio=1024 / 768, seed=0.25)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, None, w, h, fd, 4, 4, V.INTEGRATOR_PATH, False, variant=variant, default=V.SCENE_TRACER)
    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    oacc, oids, orgba, ocnt = sc.render(fd, w, h, spp=4, max_depth=4, integrator=oracle.PATH, seed=7)
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    frac, rmse = radiance_parity(acc, oacc)
    assert frac >= 0.999 and rmse <= 1e-4
    assert np.array_equal(rgba, orgba)
    assert (cnt.closest_rays, cnt.shadow_rays, cnt.paths) == (ocnt.closest_rays, ocnt.shadow_rays, ocnt.paths)
    r.close()


def test_path_config2_full_size_subrect(vk, oracle):
    """BASELINE config 2 at full size (1920x1080, 16 spp, depth 8) on the GPU; the oracle renders two
    sub-rectangles of the same frame and those must match bit for bit."""
    V = vk
    w, h = 1920, 1080
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, None, w, h, fd, 16, 8, V.INTEGRATOR_PATH, False, default=V.SCENE_TRACER)
    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    for rect in ((0, 0, 64, 48), (900, 500, 1000, 560), (1856, 1040, 1920, 1080)):
        oacc, oids, _, _ = sc.render(fd, w, h, spp=16, max_depth=8, integrator=oracle.PATH, seed=7, rect=rect, want_rgba=False)
        x0, y0, x1, y1 = rect
        assert np.array_equal(ids[y0:y1, x0:x1], oids[y0:y1, x0:x1])
        assert bits_equal(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1]), mismatch_report(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1])
    # size-independent properties at full size
    assert np.all(acc[..., 3] == 16.0)
    assert cnt.paths == w * h * 16
    assert cnt.closest_rays >= cnt.paths and cnt.shadow_rays <= cnt.closest_rays
    r.close()


def test_path_config3_full_size_subrect(vk, oracle):
    """BASELINE configs[2] at its full size (1,024 spheres, 3840x2160, 64 spp, depth 8, device LBVH, wavefront): the
    oracle renders two windows of the same frame and those must match the GPU frame bit for bit."""
    V = vk
    w, h = 3840, 2160
    scene = V.scenes.random_spheres(1024)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    r = V.Renderer(w, h, spp=64, max_depth=8, variant=V.VARIANT_WAVEFRONT, flags=V.FLAG_HIT_IDS | V.FLAG_NO_RESOLVE)
    r.set_scene(scene); r.build_bvh(); r.set_seed(7)
    r.draw(fd)
    acc, ids, cnt = r.read_accum(), r.read_hit_ids(), r.counters()
    r.close()
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    for rect in ((1900, 1060, 1948, 1092), (3808, 2136, 3840, 2160)):
        oacc, oids, _, _ = sc.render(fd, w, h, spp=64, max_depth=8, sphere_mode=oracle.S_BVH, seed=7, rect=rect, want_rgba=False)
        x0, y0, x1, y1 = rect
        assert np.array_equal(ids[y0:y1, x0:x1], oids[y0:y1, x0:x1])
        assert bits_equal(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1]), mismatch_report(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1])
    assert np.all(acc[..., 3] == 64.0) and cnt.paths == w * h * 64


def test_path_config5_tile_shard_of_the_full_frame(vk, oracle):
    """BASELINE configs[4] (100k spheres, 7680x4320, 256 spp, depth 8): one interleaved tile shard (1 of 256, the way a
    rank of the multi-GPU run holds it) of the full-size frame on the GPU; the oracle renders one of its 32x32 tiles
    near the image centre and a 128-sample half of it (the sample shard of configs[4]) -- bit-identical."""
    V = vk
    w, h = 7680, 4320
    tiles_x = w // 32
    tile = (h // 64) * tiles_x + tiles_x // 2          # a tile in the middle of the frame
    count = 256
    rank = tile % count
    tx, ty = tile % tiles_x, tile // tiles_x
    rect = (tx * 32, ty * 32, tx * 32 + 32, ty * 32 + 32)
    scene = V.scenes.grid_spheres()
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    for samp, srange in (((0, 1), (0, 256)), ((1, 2), (128, 256))):
        r = V.Renderer(w, h, spp=256, max_depth=8, variant=V.VARIANT_WAVEFRONT, flags=V.FLAG_NO_RESOLVE,
                       tile_shard=(rank, count), sample_shard=samp)
        r.set_scene(scene); r.build_bvh(); r.set_seed(2026)
        r.draw(fd)
        acc = r.read_accum()
        r.close()
        oacc, _, _, _ = sc.render(fd, w, h, spp=256, max_depth=8, sphere_mode=oracle.S_BVH, seed=2026, rect=rect, samples=srange,
                                  want_ids=False, want_rgba=False)
        x0, y0, x1, y1 = rect
        assert np.all(acc[y0:y1, x0:x1, 3] == float(srange[1] - srange[0]))
        assert bits_equal(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1]), mismatch_report(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1])


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1024])
def test_bvh_tree_matches_oracle_tree(vk, oracle, n):
    V = vk
    scene = V.scenes.random_spheres(n)
    scene.spheres, scene.sphere_mat = scene.spheres[1:], scene.sphere_mat[1:]     # drop the light: exactly n spheres
    r = V.Renderer(64, 64)
    r.set_scene(scene)
    info = r.build_bvh()
    assert info.n_nodes == max(n - 1, 1)
    gpu_nodes = r.bvh_nodes()
    sc = apply_scene(oracle, scene).build_bvh()
    assert bits_equal(gpu_nodes, sc.bvh_nodes())
    r.close()


@pytest.mark.parametrize("n", [1, 2, 17, 1024, 20000])
def test_bvh_quantised_nodes_enclose_exact_nodes(vk, n):
    """The 32-byte traversal nodes (16-bit codes) must enclose the exact child boxes in REAL arithmetic -- checked
    here with exact integer/rational arithmetic on the float bits -- and nest like them (parent codes enclose the
    children's codes).  That is what keeps the LBVH traversal identical to rule S's linear scan."""
    from fractions import Fraction
    V = vk
    scene = V.scenes.random_spheres(n) if n <= 1024 else V.scenes.grid_spheres(nx=30, ny=25, nz=27)
    if n <= 1024:
        scene.spheres, scene.sphere_mat = scene.spheres[1:], scene.sphere_mat[1:]
    r = V.Renderer(64, 64)
    r.set_scene(scene)
    r.build_bvh()
    exact = r.bvh_traversal_nodes().reshape(-1, 2, 8)       # [node][child] {lo.xyz hi.x | hi.yz index kind}: the tree the codes are made from
    qn, grid = r.bvh_qnodes()
    qn = qn.reshape(-1, 2, 4)
    r.close()
    lo = exact[:, :, 0:3].astype(np.float64)
    hi = np.stack([exact[:, :, 3], exact[:, :, 4], exact[:, :, 5]], axis=-1).astype(np.float64)
    idx = np.ascontiguousarray(exact[:, :, 6]).view(np.int32)
    kind = np.ascontiguousarray(exact[:, :, 7]).view(np.int32)
    qlo = (qn[:, :, 0:3] & 0xffff).astype(np.int64)
    qhi = (qn[:, :, 0:3] >> 16).astype(np.int64)
    ref = np.ascontiguousarray(qn[:, :, 3]).view(np.int32)
    assert np.array_equal(ref, np.where(kind != 0, ~idx, idx))
    assert np.all(qlo <= qhi)
    # float64 holds (2^23 + q) * s exactly (24 + 24 bits); adding the offset may round, so compare with a margin of
    # half a grid step first (vectorised) and exactly (Fraction) on a sample
    s, b2 = grid[0:3].astype(np.float64), grid[3:6].astype(np.float64)
    xlo, xhi = (8388608 + qlo) * s + b2, (8388608 + qhi) * s + b2
    assert np.all(xlo <= lo - 0.5 * s) and np.all(xhi >= hi + 0.5 * s)
    assert np.all(lo - xlo < 3.0 * s) and np.all(xhi - hi < 3.0 * s)          # ... and they are tight
    rng = np.random.default_rng(5)
    for node in rng.integers(0, qn.shape[0], size=min(200, qn.shape[0])):
        for k in range(2):
            for a in range(3):
                X = lambda q: Fraction(8388608 + int(q)) * Fraction(float(grid[a])) + Fraction(float(grid[3 + a]))
                assert X(qlo[node, k, a]) <= Fraction(float(exact[node, k, a]))
                assert X(qhi[node, k, a]) >= Fraction(float(hi[node, k, a]))
    # nesting: an inner child's coded box encloses the coded boxes of that child's own two children
    inner = kind == 0
    ci = idx[inner]
    plo, phi = qlo[inner], qhi[inner]
    assert np.all(plo[:, None, :] <= qlo[ci]) and np.all(phi[:, None, :] >= qhi[ci])


@pytest.mark.parametrize("case", ["2", "3", "17", "1024", "grid20k", "coincident"])
def test_traversal_tree_is_a_hierarchy_of_exact_unions_over_the_same_leaves(vk, case):
    """The wavefront's traversal nodes come from a tree of its own (top-down binned SAH, built on the device).  Rule S only
    needs: every sphere is exactly one leaf whose box is the sphere's own padded box, and every inner box is
    the exact min / max union of its two children.  Checked on the bits; also: breadth-first node ids, the reported depth,
    and that the build is deterministic."""
    V = vk
    if case == "grid20k":
        scene = V.scenes.grid_spheres(nx=30, ny=25, nz=27)
    elif case == "coincident":
        scene = V.scenes.random_spheres(64)
        scene.spheres = np.repeat(scene.spheres[:1], 5000, axis=0).copy()
        scene.spheres[:, 3] = np.linspace(0.5, 3.0, 5000, dtype=np.float32)
        scene.sphere_mat = np.full(5000, 8, dtype=np.uint32)
    else:
        scene = V.scenes.random_spheres(int(case))
        scene.spheres, scene.sphere_mat = scene.spheres[1:], scene.sphere_mat[1:]
    n = scene.spheres.shape[0]
    r = V.Renderer(64, 64)
    r.set_scene(scene)
    info = r.build_bvh()
    t = r.bvh_traversal_nodes()
    r.build_bvh()
    assert bits_equal(t, r.bvh_traversal_nodes())                    # deterministic
    r.close()
    assert info.traversal_is_sah == 1 and t.shape[0] == n - 1
    t = t.reshape(-1, 2, 8)
    lo = t[:, :, 0:3]
    hi = np.stack([t[:, :, 3], t[:, :, 4], t[:, :, 5]], axis=-1)
    idx = np.ascontiguousarray(t[:, :, 6]).view(np.int32)
    kind = np.ascontiguousarray(t[:, :, 7]).view(np.int32)
    assert set(np.unique(kind)) <= {0, 1}
    # leaves: every sphere once, with its own padded box (the same float32 operations as the device)
    leaf = kind == 1
    assert np.array_equal(np.sort(idx[leaf]), np.arange(n))
    sp = scene.spheres.astype(np.float32)
    rp = (sp[:, 3] * np.float32(1.001) + np.float32(0.001)).astype(np.float32)
    assert bits_equal(lo[leaf], (sp[idx[leaf], 0:3] - rp[idx[leaf], None]).astype(np.float32))
    assert bits_equal(hi[leaf], (sp[idx[leaf], 0:3] + rp[idx[leaf], None]).astype(np.float32))
    # inner children: every node but the root once; box = exact union of that node's two child boxes
    inner = kind == 0
    ci = idx[inner]
    assert np.array_equal(np.sort(ci), np.arange(1, n - 1))
    assert bits_equal(lo[inner], lo[ci].min(axis=1)) and bits_equal(hi[inner], hi[ci].max(axis=1))
    # breadth-first ids: a child's id is larger than its parent's; depth = longest chain of inner nodes
    parent = np.repeat(np.arange(n - 1), 2).reshape(-1, 2)[inner]
    assert np.all(ci > parent)
    depth = np.ones(n - 1, dtype=np.int64)
    order = np.argsort(ci)                                            # children in id order: parents are final before them
    for c_, p_ in zip(ci[order], parent[order]):
        depth[c_] = depth[p_] + 1
    assert depth.max() == info.traversal_depth <= 126


def test_bvh_duplicate_centres(vk, oracle):
    """Many identical Morton codes (coincident centres) exercise the index tie-break of the hierarchy."""
    V = vk
    scene = V.scenes.random_spheres(64)
    scene.spheres = np.repeat(scene.spheres[:8], 8, axis=0).copy()
    scene.spheres[:, 3] = np.linspace(0.5, 3.0, 64, dtype=np.float32)
    scene.sphere_mat = np.arange(8, 8 + 64, dtype=np.uint32)
    r = V.Renderer(64, 64)
    r.set_scene(scene)
    r.build_bvh()
    sc = apply_scene(oracle, scene).build_bvh()
    assert bits_equal(r.bvh_nodes(), sc.bvh_nodes())
    r.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_degenerate_scene_100k_coincident_centres(vk, oracle, variant):
    """Worst case for the hierarchy: 100,000 spheres around ONE centre (every Morton code equal, every box overlaps every
    other).  The Karras tree is then the radix tree of the indices: its depth stays far below the 128-entry traversal
    stacks (vkrt_bvh_info.depth <= 64 for ANY scene, enforced by vkrt_build_bvh), every ray that meets the cluster
    visits all of it, and the answer is still rule S's, bit for bit."""
    V = vk
    n = 100000
    scene = V.scenes.random_spheres(8)
    scene.spheres = np.zeros((n + 1, 4), dtype=np.float32)
    scene.spheres[0] = (0.0, 96.0, 0.0, 12.0)                                       # the light keeps its place
    scene.spheres[1:, :3] = (10.0, 40.0, 20.0)
    scene.spheres[1:, 3] = 8.0 + np.arange(n, dtype=np.float32) * np.float32(1e-5)
    scene.sphere_mat = np.concatenate([np.array([7], dtype=np.uint32), 8 + (np.arange(n, dtype=np.uint32) % 8)])
    w, h = 48, 32
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.4)
    r = V.Renderer(w, h, spp=2, max_depth=3, variant=variant, flags=V.FLAG_HIT_IDS)
    r.set_scene(scene)
    info = r.build_bvh()
    assert 17 <= info.depth <= 64, info.depth
    r.set_seed(3)
    r.draw(fd)
    acc, ids = r.read_accum(), r.read_hit_ids()
    r.close()
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    oacc, oids, _, _ = sc.render(fd, w, h, spp=2, max_depth=3, sphere_mode=oracle.S_BVH, seed=3)
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)


@pytest.mark.parametrize("variant", [0, 1])
def test_path_random_spheres_bvh_bit_exact(vk, oracle, variant):
    """BASELINE config 3's scene (1,024 random spheres, device LBVH) at a size the oracle finishes in
    seconds: the GPU's LBVH traversal must equal both the oracle's own BVH traversal and the oracle's
    LINEAR scan under rule S; the count of primary rays on which the literal chain rule of
    Tracer.comp:398-412 disagrees with rule S (the "epsilon band of grazing ties") is reported."""
    V = vk
    w, h = 320, 180
    scene = V.scenes.random_spheres(1024)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.75)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, scene, w, h, fd, 4, 8, V.INTEGRATOR_PATH, True, flags=V.FLAG_STATS, variant=variant)
    sc = apply_scene(oracle, scene).build_bvh()
    oacc, oids, orgba, ocnt = sc.render(fd, w, h, spp=4, max_depth=8, integrator=oracle.PATH, sphere_mode=oracle.S_BVH, seed=7)
    lacc, lids, _, lcnt = sc.render(fd, w, h, spp=4, max_depth=8, integrator=oracle.PATH, sphere_mode=oracle.S_LINEAR, seed=7)
    assert np.array_equal(ids, oids) and np.array_equal(ids, lids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert bits_equal(acc, lacc), mismatch_report(acc, lacc)
    assert np.array_equal(rgba, orgba)
    assert (cnt.closest_rays, cnt.shadow_rays) == (ocnt.closest_rays, ocnt.shadow_rays)
    band = ocnt.literal_vs_s_mismatch
    print("tie band: %d of %d primary rays differ between the literal chain rule and rule S" % (band, w * h))
    assert band <= w * h * 1e-3
    # literal-order brute force on the GPU agrees with the oracle's literal loop, too
    r.clear_bvh(); r.reset_counters(); r.set_frame_index(0)
    r.draw(fd)
    bacc, bids = r.read_accum(), r.read_hit_ids()
    oacc2, oids2, _, _ = sc.render(fd, w, h, spp=4, max_depth=8, integrator=oracle.PATH, sphere_mode=oracle.LITERAL, seed=7)
    assert np.array_equal(bids, oids2)
    assert bits_equal(bacc, oacc2), mismatch_report(bacc, oacc2)
    assert int((bids != ids).sum()) == band
    r.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_path_grid_100k_bvh_bit_exact(vk, oracle, variant):
    """BASELINE config 4's scene (100,000 procedural spheres) against the oracle's BVH traversal."""
    V = vk
    w, h = 192, 108
    scene = V.scenes.grid_spheres()
    assert scene.spheres.shape[0] == 100001
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.125)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, scene, w, h, fd, 4, 8, V.INTEGRATOR_PATH, True, variant=variant)
    assert r.bvh_info().n_nodes == 100000
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    oacc, oids, orgba, ocnt = sc.render(fd, w, h, spp=4, max_depth=8, integrator=oracle.PATH, sphere_mode=oracle.S_BVH, seed=7)
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert (cnt.closest_rays, cnt.shadow_rays) == (ocnt.closest_rays, ocnt.shadow_rays)
    r.close()


def test_whitted_with_bvh(vk, oracle):
    V = vk
    w, h = 256, 192
    scene = V.scenes.random_spheres(300)
    scene.materials[8::3, 7] = 1.0          # make a third of the spheres reflective
    fd = V.default_frame_data(aspect_ratio=w / h)
    r, (acc, ids, rgba, cnt) = _render_gpu(V, scene, w, h, fd, 1, 2, V.INTEGRATOR_WHITTED, True)
    sc = apply_scene(oracle, scene).build_bvh()
    oacc, oids, orgba, ocnt = sc.render(fd, w, h, spp=1, max_depth=2, integrator=oracle.WHITTED, sphere_mode=oracle.S_LINEAR)
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert np.array_equal(rgba, orgba)
    r.close()


def test_progressive_accumulation(vk, oracle):
    """Three accumulated frames == the oracle accumulating the same three frame keys."""
    V = vk
    w, h = 160, 120
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    r = V.Renderer(w, h, spp=4, max_depth=4, flags=V.FLAG_PROGRESSIVE)
    r.use_default_scene(V.SCENE_TRACER)
    r.set_seed(3)
    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    oacc = None
    for f in range(3):
        r.draw(fd)
        oacc, _, orgba, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=3, frame_index=f, accum=oacc)
    acc = r.read_accum()
    assert np.all(acc[..., 3] == 12.0)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert np.array_equal(r.read_rgba8(), orgba)
    r.reset_accum()
    r.set_frame_index(0)
    r.draw(fd)
    first, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=3, frame_index=0)
    assert bits_equal(r.read_accum(), first)
    r.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_progressive_64_frames_config2_full_size(vk, oracle, variant):
    """BASELINE configs[1] as bench.py's `cfg2` runs it: Raytracer.comp's geometry (+ the light sphere) through the path
    integrator at 1920x1080, 16 spp, depth 8, **64 progressively accumulated frames** (frame index = seed of the frame).
    The oracle accumulates the same 64 frames over two windows of the full-size image: bit-identical sums."""
    V = vk
    w, h, frames = 1920, 1080, 64
    scene = V.scenes.raytracer_default(with_emitter=True)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    r = V.Renderer(w, h, spp=16, max_depth=8, variant=variant, flags=V.FLAG_PROGRESSIVE | V.FLAG_NO_RESOLVE)
    r.set_scene(scene)
    r.set_seed(2026)
    for f in range(frames):                        # enqueued back to back: two frames in flight
        r.set_frame_index(f)
        r.draw(fd)
    acc = r.read_accum()
    r.close()
    assert np.all(acc[..., 3] == 16.0 * frames)
    sc = apply_scene(oracle, scene, fast=True)
    for rect in ((940, 520, 972, 544), (0, 0, 24, 16)):
        oacc = None
        for f in range(frames):
            oacc, _, _, _ = sc.render(fd, w, h, spp=16, max_depth=8, seed=2026, frame_index=f, rect=rect, accum=oacc,
                                      want_ids=False, want_rgba=False)
        x0, y0, x1, y1 = rect
        assert bits_equal(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1]), mismatch_report(acc[y0:y1, x0:x1], oacc[y0:y1, x0:x1])


def test_wavefront_multi_wave_equals_megakernel(vk):
    """A frame of more than 16 spp is cut into waves of <= 16 samples (53 spp: 16 + 16 + 16 + 5): the running per-pixel sum
    must still be formed in sample order, i.e. be bit-identical to the megakernel, the later waves reuse the first wave's
    primary hits; also one wave per frame (13 of 40 spp), tile + sample shards, progressive accumulation, three frames
    back to back (the lanes alternate between frames)."""
    V = vk
    w, h = 96, 72
    scene = V.scenes.random_spheres(200)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.9)
    for spp, tile, samp, n_expect in ((40, (1, 2), (1, 3), 40 * 2 // 3 - 40 // 3), (53, (0, 1), (0, 1), 53), (35, (1, 3), (0, 1), 35)):
        outs = []
        for variant in (0, 1):
            r = V.Renderer(w, h, spp=spp, max_depth=5, variant=variant, flags=V.FLAG_PROGRESSIVE | V.FLAG_HIT_IDS,
                           tile_shard=tile, sample_shard=samp)
            r.set_scene(scene); r.build_bvh(); r.set_seed(5)
            r.draw(fd); r.draw(fd); r.draw(fd)
            c = r.counters()
            outs.append((r.read_accum(), r.read_hit_ids(), (c.closest_rays, c.shadow_rays, c.paths)))
            r.close()
        assert bits_equal(outs[0][0], outs[1][0]), mismatch_report(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1], outs[1][1])
        assert outs[0][2] == outs[1][2]
        assert outs[0][0][..., 3].max() == 3 * n_expect


def test_pixel_major_groups_of_more_than_32_samples(vk, monkeypatch):
    """Depth 0 hands every pixel one contiguous range of the depth-1 arrays for its surviving samples, 32 samples at a
    time (a 32-bit survivor mask).  Waves normally hold <= 16 samples; the tuning knob VKRT_TUNE_WAVE_SPP makes one wave
    hold 45, i.e. a full group and a ragged one: still bit-identical to the megakernel, same ray counts."""
    V = vk
    w, h = 96, 72
    scene = V.scenes.random_spheres(200)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.3)
    outs = []
    for variant, wave in ((0, None), (1, "45"), (1, "33")):
        if wave:
            monkeypatch.setenv("VKRT_TUNE_WAVE_SPP", wave)
        else:
            monkeypatch.delenv("VKRT_TUNE_WAVE_SPP", raising=False)
        r = V.Renderer(w, h, spp=45, max_depth=6, variant=variant, flags=V.FLAG_HIT_IDS)
        r.set_scene(scene); r.build_bvh(); r.set_seed(9)
        r.draw(fd); r.draw(fd)
        c = r.counters()
        outs.append((r.read_accum(), r.read_hit_ids(), (c.closest_rays, c.shadow_rays, c.paths)))
        r.close()
    for o in outs[1:]:
        assert bits_equal(outs[0][0], o[0]), mismatch_report(outs[0][0], o[0])
        assert np.array_equal(outs[0][1], o[1])
        assert outs[0][2] == o[2]


def test_serial_waves_flag_is_bit_identical(vk):
    """VKRT_FLAG_SERIAL_WAVES (bench.py's measurement aid) only changes where the waves are enqueued."""
    V = vk
    w, h = 200, 120
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.3)
    scene = V.scenes.random_spheres(512)
    out = []
    for flags in (V.FLAG_LAUNCH_TIMING, V.FLAG_SERIAL_WAVES | V.FLAG_LAUNCH_TIMING, 0):
        r = V.Renderer(w, h, spp=16, max_depth=6, variant=V.VARIANT_WAVEFRONT, flags=flags)
        r.set_scene(scene)
        r.build_bvh()
        r.set_seed(5)
        for i in range(2):
            r.set_frame_index(i)
            r.draw(fd)
        if flags:
            out.append((r.read_accum(), r.read_rgba8(), r.last_frame_traversal_timing()))
        else:                        # without VKRT_FLAG_LAUNCH_TIMING the hot path records no per-launch events
            with pytest.raises(V.VkrtError):
                r.last_frame_traversal_timing()
            out.append((r.read_accum(), r.read_rgba8(), None))
        r.close()
    assert bits_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert bits_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
    assert out[0][2][1] == out[1][2][1] and out[1][2][0] > 0.0          # the same traversal launches, timed


def test_tile_and_sample_shards_recombine(vk, oracle):
    """tile shards are disjoint pixels -> recombination is bit-identical to one GPU;
    sample shards change the summation order -> equal to the oracle's per-range sums added in order."""
    V = vk
    w, h = 200, 150      # not a multiple of the 32-px tile
    scene = V.scenes.random_spheres(256)
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.3)

    def make(tile=(0, 1), samp=(0, 1)):
        r = V.Renderer(w, h, spp=8, max_depth=6, tile_shard=tile, sample_shard=samp, flags=V.FLAG_NO_RESOLVE)
        r.set_scene(scene); r.build_bvh(); r.set_seed(11)
        return r

    full = make(); full.draw(fd); ref = full.read_accum()
    gather = make()
    for rank in range(3):
        part = make(tile=(rank, 3)); part.draw(fd)
        ptr, n = part.pack_shard()
        part.wait_idle()
        gather.unpack_shard(ptr, rank, 3, add=False)
        gather.wait_idle()
        part.close()
    assert bits_equal(gather.read_accum(), ref), mismatch_report(gather.read_accum(), ref)

    sc = apply_scene(oracle, scene).build_bvh()
    gather.reset_accum()
    expect = None
    for srank in range(2):
        part = make(samp=(srank, 2)); part.draw(fd)
        ptr, n = part.pack_shard(); part.wait_idle()
        gather.unpack_shard(ptr, 0, 1, add=(srank > 0)); gather.wait_idle()
        part.close()
        expect, _, _, _ = sc.render(fd, w, h, spp=8, max_depth=6, sphere_mode=oracle.S_BVH, seed=11,
                                    samples=(srank * 4, srank * 4 + 4), accum=expect, want_ids=False, want_rgba=False)
    got = gather.read_accum()
    assert bits_equal(got, expect), mismatch_report(got, expect)
    frac, rmse = radiance_parity(got, ref)
    assert frac >= 0.999 and rmse <= 1e-4
    full.close(); gather.close()


def test_error_behaviour(vk):
    V = vk
    with pytest.raises(V.VkrtError) as e:
        V.Renderer(0, 10)
    assert e.value.code == V._lib.BAD_ARG and "[app] - err ::" in str(e.value)
    with pytest.raises(V.VkrtError):
        V.Renderer(16, 16, device_id=99)
    r = V.Renderer(16, 16)
    with pytest.raises(V.VkrtError):
        r.read_hit_ids()                      # created without FLAG_HIT_IDS
    with pytest.raises(V.VkrtError):
        r.set_sampling(0, 4)
    scene = V.scenes.tracer_default()
    scene.sphere_mat = scene.sphere_mat + 100
    r.set_scene(scene)
    with pytest.raises(V.VkrtError):
        r.draw(V.default_frame_data())
    r.close()


def test_graphics_device_mirror(vk):
    """The reference's call pattern (Main.cpp:105-196): Construct -> Draw per frame -> WaitIdle -> Destruct."""
    V = vk
    dev = V.GraphicsDevice()
    info = V.GraphicsDevice.CreateInfo(None, 3, 2, 256, False)
    assert dev.Construct(info) == V.GraphicsDevice.Error.SUCCESS
    cam = V.default_camera()
    fd = V.default_frame_data(camera=cam)
    for _ in range(3):
        cam.move_forward(1.0)
        fd.camera = cam.data
        dev.Draw(fd)
    dev.WaitIdle()
    img = dev.renderer.read_rgba8()
    assert img.shape == (256, 256, 4) and img[..., 3].min() == 255 and img[..., :3].max() > 0
    assert dev.Destruct() == V.GraphicsDevice.Error.SUCCESS
    bad = V.GraphicsDevice.CreateInfo(None, 3, 2, 256, False, device_id=77)
    assert V.GraphicsDevice().Construct(bad) == V.GraphicsDevice.Error.NO_SUITABLE_GPU


@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_matches_golden_frames(vk, variant):
    """The committed fixtures (tests/golden/frames.npz, written by the oracle in the build container)."""
    import os
    from golden.make_golden import FRAMES, scene_for
    V = vk
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames.npz"))
    for name, (scn, w, h, spp, depth, integ, mode, seed, fseed) in FRAMES.items():
        fd = gold[name + ".frame_data"].tobytes()
        r = V.Renderer(w, h, spp=spp, max_depth=depth, integrator=integ, variant=variant, flags=V.FLAG_HIT_IDS)
        r.set_scene(scene_for(scn))
        if mode == 2:
            r.build_bvh()
        r.set_seed(seed)
        r.draw(fd)
        c = r.counters()
        assert bits_equal(r.read_accum(), gold[name + ".accum"]), name
        assert np.array_equal(r.read_hit_ids(), gold[name + ".ids"]), name
        assert np.array_equal(r.read_rgba8(), gold[name + ".rgba"]), name
        assert [c.closest_rays, c.shadow_rays, c.paths] == gold[name + ".counts"].tolist(), name
        r.close()


def test_pack_layout_matches_host_statement(vk):
    """vkrt_pack_shard's slot order == sharding.owned_pixels (what the CPU gloo test relies on)."""
    import torch
    from vk_renderer_b200.sharding import owned_pixels
    V = vk
    w, h = 200, 150
    fd = V.default_frame_data(aspect_ratio=w / h)
    for rank, count in ((0, 1), (1, 3)):
        r = V.Renderer(w, h, spp=1, max_depth=1, tile_shard=(rank, count), flags=V.FLAG_NO_RESOLVE)
        r.use_default_scene(V.SCENE_TRACER)
        r.draw(fd)
        acc = r.read_accum().reshape(-1, 4)
        n = r.shard_floats(0)
        buf = torch.zeros(n, dtype=torch.float32, device="cuda")
        r.pack_shard_into(buf.data_ptr(), n)
        r.wait_idle()
        torch.cuda.synchronize()
        packed = buf.cpu().numpy().reshape(-1, 4)
        pix = owned_pixels(w, h, rank, count)
        assert bits_equal(packed[: pix.shape[0]][pix >= 0], acc[pix[pix >= 0]])
        assert not packed[: pix.shape[0]][pix < 0].any()
        r.close()


def test_primary_hit_ids_config4_full_size(vk):
    """BASELINE config 4 at its full size: the primary nearest-hit primitive id of every one of the 2,073,600 pixels is
    bit-exact against the oracle's (tests/golden/cfg4_primary_ids.npz, made by tests/golden/make_cfg4_ids.py); the band of
    grazing ties in which the reference's literal in-order loop would name another sphere is 2 pixels (reported)."""
    import os
    from conftest import ROOT
    V = vk
    z = np.load(os.path.join(ROOT, "tests", "golden", "cfg4_primary_ids.npz"))
    w, h = 1920, 1080
    scene = V.scenes.grid_spheres()
    assert scene.digest() == str(z["scene_sha"][0]) and int(z["tie_band"][0]) == 2
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    for variant in (0, 1):
        r = V.Renderer(w, h, spp=1, max_depth=1, variant=variant, flags=V.FLAG_HIT_IDS | V.FLAG_NO_RESOLVE)
        r.set_scene(scene); r.build_bvh(); r.set_seed(2026)
        r.draw(fd)
        ids = r.read_hit_ids()
        r.close()
        assert np.array_equal(ids, z["ids"]), "variant %d: %d primary hit ids differ" % (variant, int((ids != z["ids"]).sum()))


def test_full_size_properties_config4(vk):
    """Size-independent checks at BASELINE config 4's full size (100k spheres, 1920x1080, 16 spp, depth 8):
    the two kernel variants agree bit for bit, a second draw of the same frame index is idempotent, and
    tile shards partition the ray counts exactly."""
    V = vk
    w, h = 1920, 1080
    scene = V.scenes.grid_spheres()
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    outs = []
    for variant in (0, 1):
        r = V.Renderer(w, h, spp=16, max_depth=8, variant=variant, flags=V.FLAG_NO_RESOLVE)
        r.set_scene(scene); r.build_bvh(); r.set_seed(2026)
        r.draw(fd)
        a = r.read_accum()
        c = r.counters()
        r.set_frame_index(0); r.reset_counters()
        r.draw(fd)
        assert bits_equal(r.read_accum(), a)
        outs.append((a, (c.closest_rays, c.shadow_rays, c.paths)))
        r.close()
    assert bits_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert outs[0][1][2] == w * h * 16 and np.all(outs[0][0][..., 3] == 16.0)
    assert np.isfinite(outs[0][0]).all() or np.isnan(outs[0][0]).sum() < 100
    total = np.zeros(3, dtype=np.int64)
    for rank in range(2):
        r = V.Renderer(w, h, spp=16, max_depth=8, variant=1, tile_shard=(rank, 2), flags=V.FLAG_NO_RESOLVE)
        r.set_scene(scene); r.build_bvh(); r.set_seed(2026)
        r.draw(fd)
        c = r.counters()
        total += np.array([c.closest_rays, c.shadow_rays, c.paths])
        part = r.read_accum()
        own = part[..., 3] > 0
        assert bits_equal(part[own], outs[0][0][own])
        r.close()
    assert tuple(total.tolist()) == outs[0][1]


def test_cpp_headless_loop_matches_python_mirror(vk, tmp_path):
    """The C++ GraphicsDevice drop-in driven by the headless Main.cpp loop produces the same image, byte for
    byte, as the Python mirror driven with the same camera script and the same libc rand() seeds."""
    import ctypes
    import os
    import subprocess
    import importlib.util
    V = vk
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_vkrt_build", os.path.join(root, "vk-renderer_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    exe = b.build_host()
    out = tmp_path / "frame.ppm"
    frames, res = 5, 256
    r = subprocess.run([exe, "--frames", str(frames), "--res", str(res), "--seed", "42", "--out", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "%d frames in" % frames in r.stdout
    data = open(out, "rb").read()
    header = b"P6\n%d %d\n255\n" % (res, res)
    assert data.startswith(header)
    cpp = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(res, res, 3)[::-1]       # back to bottom-up rows

    dev = V.GraphicsDevice()
    assert dev.Construct(V.GraphicsDevice.CreateInfo(None, 3, 2, res, False)) == V.GraphicsDevice.Error.SUCCESS
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(42)
    cam = V.default_camera()
    fd = V.default_frame_data(camera=cam)
    dev.Draw(fd)
    for f in range(1, frames):
        if f % 3 == 0: cam.move_forward(0.5)
        elif f % 3 == 1: cam.move_right(0.25)
        else: cam.move_up(0.125)
        fd.camera = cam.data
        dev.Draw(fd)
    dev.WaitIdle()
    py = dev.renderer.read_rgba8()[..., :3]
    dev.Destruct()
    assert np.array_equal(cpp, py)


def test_present_filter_matches_oracle(vk, oracle):
    """Fullscreen.frag (variance-gated NSEW blur, bilinear clamp-to-border sampler) on two
    consecutive traced frames, resampled to the 1024x768 window of Main.cpp:92 -- byte-exact against the oracle,
    including the reference's quirk that binding 0/1 are fixed image slots."""
    V = vk
    res = 256
    r = V.Renderer(res, res, spp=4, max_depth=4, frames_in_flight=2)
    r.use_default_scene(V.SCENE_TRACER)
    r.set_seed(9)
    cam = V.default_camera()
    slots = {}
    for f in range(3):
        cam.move_right(2.0)
        r.draw(V.default_frame_data(camera=cam, seed=0.1 * f))
        slots[f % 2] = r.read_rgba8()                # draw f writes image slot f % 2 like state.currentFrame (:1341)
        if f >= 1:
            got = r.present(1024, 768)
            want = oracle.present(slots[0], slots[1], 1024, 768)
            assert np.array_equal(got, want), "frame %d: %d bytes differ" % (f, int((got != want).sum()))
            assert got[..., 3].min() == 255
    # the gate really switches: with a moving camera both branches of :28 occur
    a, b = slots[0].astype(np.float32) / 255, slots[1].astype(np.float32) / 255
    var = ((a - b)[..., :3] ** 2).sum(-1)
    assert (var > 0.0005).any() and (var <= 0.0005).any()
    r.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_edge_cases(vk, oracle, variant):
    """Ragged / degenerate inputs: 1x1 and odd-sized images, spp = 1 (single lane), depth 1, a scene with no
    light, no spheres at all, an empty scene, triangles together with the LBVH, many lights."""
    V = vk
    fd = V.default_frame_data(aspect_ratio=1.5, seed=0.4)

    def check(scene, w, h, spp, depth, bvh, name):
        r = V.Renderer(w, h, spp=spp, max_depth=depth, variant=variant, flags=V.FLAG_HIT_IDS)
        r.set_scene(scene)
        if bvh:
            r.build_bvh()
        r.set_seed(4)
        r.draw(fd)
        acc, ids, rgba, c = r.read_accum(), r.read_hit_ids(), r.read_rgba8(), r.counters()
        r.close()
        sc = apply_scene(oracle, scene)
        if bvh:
            sc.build_bvh()
        oacc, oids, orgba, oc = sc.render(fd, w, h, spp=spp, max_depth=depth, sphere_mode=oracle.S_BVH if bvh else oracle.LITERAL, seed=4)
        assert np.array_equal(ids, oids), name
        assert bits_equal(acc, oacc), name + ": " + mismatch_report(acc, oacc)
        assert np.array_equal(rgba, orgba), name
        assert (c.closest_rays, c.shadow_rays, c.paths) == (oc.closest_rays, oc.shadow_rays, oc.paths), name

    base = V.scenes.random_spheres(50)
    check(base, 1, 1, 1, 1, True, "1x1")
    check(base, 37, 23, 1, 8, True, "odd size, spp 1")
    check(base, 33, 65, 3, 1, True, "depth 1, spp 3")
    check(base, 70, 40, 5, 3, False, "literal loop on a synthetic scene")

    dark = V.scenes.random_spheres(50)
    dark.materials[7, 4:7] = 0.0                                   # the light no longer emits: no shadow rays at all
    check(dark, 40, 30, 2, 4, True, "no light")

    planes_only = V.scenes.tracer_default()
    planes_only.spheres = np.zeros((0, 4), np.float32); planes_only.sphere_mat = np.zeros(0, np.uint32)
    check(planes_only, 40, 30, 2, 4, True, "no spheres, BVH over nothing")
    check(planes_only, 40, 30, 2, 4, False, "no spheres")

    empty = V.scenes.tracer_default()
    empty.spheres = np.zeros((0, 4), np.float32); empty.sphere_mat = np.zeros(0, np.uint32)
    empty.planes = np.zeros((0, 4), np.float32); empty.plane_mat = np.zeros(0, np.uint32)
    empty.triangles = np.zeros((0, 12), np.float32)
    check(empty, 16, 16, 2, 4, False, "empty scene")

    tris = V.scenes.random_spheres(80)
    rng = np.random.default_rng(3)
    t = np.zeros((6, 12), np.float32)
    for k in range(6):
        c = rng.uniform([-40, 10, -40], [40, 100, 40])
        for v in range(3):
            t[k, 4 * v:4 * v + 3] = c + rng.uniform(-25, 25, 3)
    tris.triangles = t
    tris.tri_mat = 5
    check(tris, 64, 48, 2, 5, True, "triangles + LBVH")

    lights = V.scenes.random_spheres(60)
    lights.materials[8:20, 4:7] = 40.0                             # 12 more emissive spheres: 13 lights
    check(lights, 48, 36, 2, 3, True, "13 lights")
    two = V.scenes.random_spheres(60)
    two.materials[8, 4:7] = 40.0                                   # 2 lights: the first scene past the fused (<= 1 light) pipeline
    check(two, 48, 36, 3, 4, True, "2 lights")
    check(two, 48, 36, 3, 4, False, "2 lights, literal loop")
    too_many = V.scenes.random_spheres(60)
    too_many.materials[8:30, 4:7] = 40.0
    r = V.Renderer(16, 16, variant=variant)
    r.set_scene(too_many)
    with pytest.raises(V.VkrtError):
        r.draw(fd)
    r.close()
