"""The oracle pinned against THE REFERENCE ITSELF: the compiled shaders the engine loads (ref:
Assets/Compiled/Raytracer.comp.spv, Tracer.comp.spv, Fullscreen.frag.spv; Source/GraphicsDevice.cpp:1086-1091) were run
on the CPU by oracle/spirv_interp.py and their outputs committed as tests/golden/spirv_vectors.npz
(tests/golden/make_spirv_vectors.py).  The interpreter executes every float op as one binary32 operation without
contraction and takes sin / cos / pow from libm, the oracle follows the repository's own arithmetic contract (a few
fused multiply-adds, own polynomials: DESIGN.md section 2) -- two legal executions of the same shader, so float results
are compared with the tolerances stated below and discrete results (8-bit image, primitive ids, hit / miss) exactly.

Tracer.comp's float-hash rand() is a stated deviation (DESIGN.md "RNG"): while the reference binary ran, each of its
seven rand() call sites was answered with the repository's integer RNG for that site's dimension, so these vectors also
pin the call-site -> dimension mapping, the sample and depth counting and the order of the draws.

Where /root/reference is mounted the fixtures are additionally re-derived live for a sample of pixels."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "spirv_vectors.npz")
SPV = "/root/reference/Assets/Compiled"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _unorm8(a):
    """The rgba8 store conversion on an array (oracle/spirv_interp.unorm8 / Tracer.comp:592's imageStore)."""
    a = np.where(np.isnan(a), 0.0, a).astype(np.float32)
    return np.floor(np.clip(a, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.int32)


def _frame_data(vk, aspect, seed, gold=None):
    """Default camera / light (Source/Main.cpp:134-141), or the fixture's second view when `gold` is given."""
    fd = vk.default_frame_data(aspect_ratio=float(aspect), seed=float(seed))
    if gold is not None:
        for name, v in zip(("pos", "dir", "right", "up"), gold["view2_camera"]):
            a = getattr(fd.camera, name)
            a.x, a.y, a.z = float(v[0]), float(v[1]), float(v[2])
        fd.light_pos.x, fd.light_pos.y, fd.light_pos.z = [float(x) for x in gold["view2_light"]]
    return fd


def _check_whitted(acc, rgba, gold, key="whitted"):
    wt = gold[key + "_texels"]
    d = np.abs(wt[..., :3] - acc[..., :3]).max(-1)
    assert d.max() <= 1e-3 and (d <= 1e-5).mean() >= 0.99, (d.max(), (d <= 1e-5).mean())   # measured: 4.4e-4, 99.7 %
    assert np.array_equal(_unorm8(wt[..., :3]), rgba[..., :3].astype(np.int32))            # the image the engine shows
    assert (wt[..., 3] == 1.0).all()


def _check_path(acc, ids, rgba, gold, key="path"):
    pt, pr = gold[key + "_texels"], gold[key + "_radiance_sum"]
    # primary nearest-hit primitive ids: bit-exact (BASELINE.json north_star; no grazing tie in these views)
    assert np.array_equal(ids, gold[key + "_primary_id"])
    # the 8-bit image the engine presents: identical
    assert np.array_equal(_unorm8(pt[..., :3]), rgba[..., :3].astype(np.int32))
    # accumulated linear radiance: north_star's bar (1e-3 relative for >= 99.9 % of the pixels, RMSE <= 1e-4)
    rel = np.abs(pr - acc[..., :3]) / np.maximum(np.abs(acc[..., :3]), 1e-3)
    assert (rel <= 1e-3).all(-1).mean() >= 0.999, (rel <= 1e-3).all(-1).mean()               # measured: 99.93 %
    assert rel.max() <= 5e-3
    spp = acc[..., 3:4]
    assert (spp == 4.0).all()
    assert np.sqrt(np.mean((pr / spp - acc[..., :3] / spp) ** 2)) <= 1e-4                  # measured: 7.8e-5


def test_oracle_whitted_matches_the_reference_spirv(vk, oracle, gold):
    """Raytracer.comp.spv main() -- BASELINE configs[0]'s shader -- against the oracle's whitted integrator."""
    h, w = gold["whitted_texels"].shape[:2]
    fd = _frame_data(vk, w / h, 0.0)
    acc, _, rgba, _ = oracle.Scene().use_default(oracle.SCENE_RAYTRACER).render(fd, w, h, spp=1, max_depth=2, integrator=oracle.WHITTED)
    _check_whitted(acc, rgba, gold)


def test_oracle_path_matches_the_reference_spirv(vk, oracle, gold):
    """Tracer.comp.spv main() (the shader the engine dispatches) against the oracle's path integrator."""
    h, w = gold["path_texels"].shape[:2]
    fd = _frame_data(vk, gold["path_aspect"][0], gold["path_frame_seed"][0])
    acc, ids, rgba, _ = oracle.Scene().use_default(oracle.SCENE_TRACER).render(
        fd, w, h, spp=4, max_depth=4, integrator=oracle.PATH, seed=int(gold["path_seed"][0]))
    _check_path(acc, ids, rgba, gold)
    t_ref = gold["path_primary_t"]
    assert (t_ref[gold["path_primary_id"] != 0] > 1e-3).all() and (t_ref[gold["path_primary_id"] == 0] == 0).all()


def test_oracle_matches_the_reference_spirv_from_a_second_view(vk, oracle, gold):
    """Both compute shaders from a camera that sees the host's triangle front-on (Tracer.comp:378-396, mirror) and the
    room's open side (misses, Tracer.comp:445), another light position, seed and frame index."""
    ids = gold["path2_primary_id"]
    kinds = set((ids >> 28).ravel().tolist())
    assert kinds == {0, 1, 2, 3}, kinds                                    # miss, triangle, spheres, planes all occur
    h, w = gold["whitted2_texels"].shape[:2]
    fd = _frame_data(vk, w / h, 0.0, gold)
    acc, _, rgba, _ = oracle.Scene().use_default(oracle.SCENE_RAYTRACER).render(fd, w, h, spp=1, max_depth=2, integrator=oracle.WHITTED)
    _check_whitted(acc, rgba, gold, "whitted2")
    h, w = gold["path2_texels"].shape[:2]
    fd = _frame_data(vk, gold["path2_aspect"][0], gold["path2_frame_seed"][0], gold)
    acc, ids, rgba, _ = oracle.Scene().use_default(oracle.SCENE_TRACER).render(
        fd, w, h, spp=4, max_depth=4, integrator=oracle.PATH, seed=int(gold["path2_seed"][0]), frame_index=int(gold["path2_frame_index"][0]))
    _check_path(acc, ids, rgba, gold, "path2")


def _check_config0(acc, rgba):
    """BASELINE.json configs[0] at its full size: Raytracer.comp.spv main() over 640x480 (tests/golden/spirv_config0.npz)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "spirv_config0.npz"))
    u8, fl = z["whitted_full_rgb8"].astype(np.int32), z["whitted_full_float_4x4"]
    assert u8.shape == (480, 640, 3) and rgba.shape[:2] == (480, 640)
    d = np.abs(u8 - rgba[..., :3].astype(np.int32)).max(-1)
    # measured: 307,170 of 307,200 pixels identical, the other 30 off by one LSB (a colour within 1e-5 of a rounding boundary)
    assert d.max() <= 1 and (d == 0).mean() >= 0.9995, (d.max(), (d == 0).mean())
    df = np.abs(fl - acc[::4, ::4, :3]).max(-1)
    assert df.max() <= 2e-3 and (df <= 1e-5).mean() >= 0.99, (df.max(), (df <= 1e-5).mean())   # measured: 8.5e-4, 99.7 %


def test_oracle_config0_matches_the_reference_spirv(vk, oracle):
    fd = _frame_data(vk, 640 / 480, 0.0)
    acc, _, rgba, _ = oracle.Scene().use_default(oracle.SCENE_RAYTRACER).render(fd, 640, 480, spp=1, max_depth=2, integrator=oracle.WHITTED)
    _check_config0(acc, rgba)


@pytest.mark.gpu
def test_gpu_config0_matches_the_reference_spirv(vk):
    r = vk.Renderer(640, 480, spp=1, max_depth=2, integrator=vk.INTEGRATOR_WHITTED)
    r.use_default_scene(vk.SCENE_RAYTRACER)
    r.draw(_frame_data(vk, 640 / 480, 0.0))
    acc, rgba = r.read_accum(), r.read_rgba8()
    r.close()
    _check_config0(acc, rgba)


def _scene3(vk, gold):
    """The custom 4-sphere / 5-plane scene of the path3 vectors as a vk_renderer_b200.scenes.Scene (vkrt_material rows)."""
    g = gold["scene3_materials"]
    mats = np.zeros((len(g), 12), np.float32)
    mats[:, :8] = g[:, :8]
    mats.view(np.uint32)[:, 8] = g[:, 8].astype(np.uint32)
    tri = np.zeros((1, 12), np.float32)
    tri[0, 0:3], tri[0, 4:7], tri[0, 8:11] = (10, 10, 0), (0, 20, 0), (-10, 10, 0)      # Source/GraphicsDevice.cpp:798-803
    return vk.scenes.Scene("scene3", mats, gold["scene3_spheres"], gold["scene3_sphere_mat"], gold["scene3_planes"],
                           gold["scene3_plane_mat"], tri, 9)


def _check_path3(acc, rgba, gold):
    """Two emissive spheres, rough dielectric and metal: the image is more sensitive than the default scene's, a pixel
    or two of the 1,728 take another decision in one sample path (measured: 1 pixel differs in 8 bit, 2 beyond 1e-3)."""
    pt, pr = gold["path3_texels"], gold["path3_radiance_sum"]
    d = np.abs(_unorm8(pt[..., :3]) - rgba[..., :3].astype(np.int32)).max(-1)
    assert (d == 0).mean() >= 0.997, (d == 0).mean()
    rel = np.abs(pr - acc[..., :3]) / np.maximum(np.abs(acc[..., :3]), 1e-3)
    worst = rel.max(-1)
    assert (worst <= 1e-3).mean() >= 0.996, (worst <= 1e-3).mean()
    assert np.percentile(worst, 99) <= 1e-4                                    # and the bulk agrees to 1e-4


def test_oracle_two_light_scene_matches_the_reference_spirv(vk, oracle, gold):
    """The reference binary with other scene constants (oracle/spirv_interp.set_tracer_scene): the light loop with two
    emissive spheres (Tracer.comp:458-503; RNG dimensions 3 + 2l / 4 + 2l, light terms summed in sphere order)."""
    from helpers import apply_scene
    scene = _scene3(vk, gold)
    h, w = gold["path3_texels"].shape[:2]
    fd = _frame_data(vk, gold["path3_aspect"][0], gold["path3_frame_seed"][0])
    acc, _, rgba, cnt = apply_scene(oracle, scene).render(fd, w, h, spp=4, max_depth=4, integrator=oracle.PATH,
                                                          seed=int(gold["path3_seed"][0]), frame_index=int(gold["path3_frame_index"][0]))
    assert cnt.shadow_rays > 1.5 * cnt.closest_rays * 0.9          # two shadow rays per diffuse hit
    _check_path3(acc, rgba, gold)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_two_light_scene_matches_the_reference_spirv(vk, gold, variant):
    """Two lights take the wavefront's four-kernel pipeline (extend / classify / shadow / shade)."""
    h, w = gold["path3_texels"].shape[:2]
    r = vk.Renderer(w, h, spp=4, max_depth=4, variant=variant)
    r.set_scene(_scene3(vk, gold))
    r.set_seed(int(gold["path3_seed"][0]))
    r.set_frame_index(int(gold["path3_frame_index"][0]))
    r.draw(_frame_data(vk, gold["path3_aspect"][0], gold["path3_frame_seed"][0]))
    acc, rgba = r.read_accum(), r.read_rgba8()
    r.close()
    _check_path3(acc, rgba, gold)


def _scene4(vk, gold):
    g = gold["scene4_materials"]
    mats = np.zeros((len(g), 12), np.float32)
    mats[:, :8] = g[:, :8]
    mats.view(np.uint32)[:, 8] = g[:, 8].astype(np.uint32)
    tri = np.zeros((1, 12), np.float32)
    tri[0, 0:3], tri[0, 4:7], tri[0, 8:11] = (10, 10, 0), (0, 20, 0), (-10, 10, 0)
    return vk.scenes.Scene("scene4", mats, gold["scene4_spheres"], gold["scene4_sphere_mat"], gold["scene3_planes"],
                           gold["scene3_plane_mat"], tri, 9)


def _check_path4(rgba, gold):
    d = np.abs(_unorm8(gold["path4_texels"][..., :3]) - rgba[..., :3].astype(np.int32)).max(-1)
    assert d.max() <= 1 and (d == 0).mean() >= 0.995, (d.max(), (d == 0).mean())       # measured: 99.75 % identical, rest 1 LSB


def test_oracle_literal_tie_rule_matches_the_reference_spirv(vk, oracle, gold):
    """Ties: coincident spheres and a sphere 5e-4 behind another's front.  The reference's in-order loop lets the LATER
    sphere win (t < cur + EPSILON, Tracer.comp:402); the oracle's literal mode -- what every scene without a BVH uses --
    must do the same, and rule S (BVH scenes only) is seen to differ here, which is why it is a stated rule."""
    from helpers import apply_scene
    scene = _scene4(vk, gold)
    h, w = gold["path4_texels"].shape[:2]
    fd = _frame_data(vk, w / h, gold["path4_frame_seed"][0])
    kw = dict(spp=4, max_depth=4, integrator=oracle.PATH, seed=int(gold["path4_seed"][0]))
    _, ids, rgba, _ = apply_scene(oracle, scene).render(fd, w, h, sphere_mode=oracle.LITERAL, **kw)
    _check_path4(rgba, gold)
    assert ((ids >> 28) == 2).any() and ((ids & 0x0FFFFFFF)[(ids >> 28) == 2] == 2).any()      # the later coincident sphere wins
    _, ids_s, rgba_s, _ = apply_scene(oracle, scene).render(fd, w, h, sphere_mode=oracle.S_LINEAR, **kw)
    assert ((ids_s & 0x0FFFFFFF)[(ids_s >> 28) == 2] == 0).any() and not np.array_equal(rgba_s, rgba)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_literal_tie_rule_matches_the_reference_spirv(vk, gold, variant):
    h, w = gold["path4_texels"].shape[:2]
    r = vk.Renderer(w, h, spp=4, max_depth=4, variant=variant)
    r.set_scene(_scene4(vk, gold))
    r.set_seed(int(gold["path4_seed"][0]))
    r.draw(_frame_data(vk, w / h, gold["path4_frame_seed"][0]))
    rgba = r.read_rgba8()
    r.close()
    _check_path4(rgba, gold)


def test_integer_rng_reproduces_the_float_hash_distribution(vk, oracle, gold):
    """The one stated deviation, checked as a distribution: Tracer.comp.spv ran with its OWN float-hash rand()
    (nothing substituted) for 48 radiance() calls per pixel; the oracle with the integer RNG, over 12 independent seeds,
    gives the spread an unbiased renderer shows at that sample count, and the float-hash image has to sit inside it --
    as a whole and per image quarter.  (Column 0 is excluded: the shader's rand() is NaN there.)"""
    fh = gold["floathash_mean"].astype(np.float64)
    h, w = fh.shape[:2]
    n = int(gold["floathash_spp"][0])
    fd = _frame_data(vk, w / h, gold["floathash_frame_seed"][0])
    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    runs = []
    for seed in range(12):
        acc, _, _, _ = sc.render(fd, w, h, spp=n, max_depth=4, integrator=oracle.PATH, seed=1000 + seed)
        runs.append(acc[..., :3].astype(np.float64) / n)
    runs = np.array(runs)
    regions = [(slice(0, h), slice(1, w)), (slice(0, h // 2), slice(1, w // 2)), (slice(0, h // 2), slice(w // 2, w)),
               (slice(h // 2, h), slice(1, w // 2)), (slice(h // 2, h), slice(w // 2, w))]
    for ys, xs in regions:
        mine = runs[:, ys, xs].mean(axis=(1, 2))                    # (12, 3): per seed, per channel
        mu, sd = mine.mean(0), mine.std(0, ddof=1)
        z = (fh[ys, xs].mean(axis=(0, 1)) - mu) / sd
        assert (sd / mu < 0.06).all()                               # the yardstick itself is tight (1-3 %)
        # measured: whole image 0.0 .. 0.8 sigma; quarters within 1.2 sigma except the top-left one, where the float hash
        # renders 5 % brighter (2.2 .. 3.2 sigma) -- cos-hash streams of neighbouring pixels are correlated, it is a weak
        # generator, which is why it was replaced
        assert (np.abs(z) < 4.5).all(), (z, mu, sd)


def test_oracle_present_filter_matches_the_reference_spirv(oracle, gold):
    """Fullscreen.frag.spv (variance gate, four taps, v flip) against the oracle's present filter."""
    a, b, pc = gold["present_binding0"], gold["present_binding1"], gold["present_color"]
    got = oracle.present(a, b, pc.shape[1], pc.shape[0])
    want = _unorm8(pc[..., :3])
    diff = np.abs(want - got[..., :3].astype(np.int32)).max(-1)
    assert diff.max() <= 1 and (diff == 0).mean() >= 0.999, (diff.max(), (diff == 0).mean())   # measured: 1 pixel of 3072 off by one
    assert (pc[..., 3] == 1.0).all() and (got[..., 3] == 255).all()
    # both branches of Fullscreen.frag:28 occur in the fixture
    fa, fb = a.astype(np.float32) / 255, b.astype(np.float32) / 255
    var = ((fa - fb)[..., :3] ** 2).sum(-1)
    assert (var > 0.0005).any() and (var <= 0.0005).any()


def test_oracle_intersections_match_the_reference_spirv(oracle, gold):
    """calc_sphere_intersect / calc_plane_intersect / calc_tri_intersect of Tracer.comp.spv (ref: Tracer.comp:314-372)
    called directly: the same hit / miss decision for every ray, t within a few ulp."""
    L = oracle.lib()
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    n = len(gold["kat_o"])
    mine = {k: np.zeros(n, np.float32) for k in ("sphere", "plane", "tri")}
    for i in range(n):
        o, d = f3(gold["kat_o"][i]), f3(gold["kat_d"][i])
        s = oracle.Sphere(*[float(x) for x in gold["kat_sphere"][i]])
        p = oracle.Plane(*[float(x) for x in gold["kat_plane"][i]])
        tri = np.zeros(12, np.float32)
        tri[0:3], tri[4:7], tri[8:11] = gold["kat_tri"][i]
        mine["sphere"][i] = L.orc_sphere_intersect(o, d, C.byref(s))
        mine["plane"][i] = L.orc_plane_intersect_tracer(o, d, C.byref(p))
        mine["tri"][i] = L.orc_tri_intersect(o, d, tri.ctypes.data_as(C.c_void_p), C.c_float(1e-3))
    for name, tol, min_hits in (("sphere", 2e-5, 100), ("plane", 1e-4, 100), ("tri", 2e-5, 40)):
        ref, got = gold["kat_t_" + name], mine[name]
        hit_r, hit_g = ref > 1e-3, got > 1e-3                       # what trace_ray accepts (NaN compares false)
        assert hit_r.sum() >= min_hits and (~hit_r).sum() >= 30, (name, hit_r.sum())
        assert np.array_equal(hit_r, hit_g), "%s: %d hit / miss decisions differ" % (name, (hit_r != hit_g).sum())
        rel = np.abs(got[hit_r] - ref[hit_r]) / np.abs(ref[hit_r])
        assert rel.max() <= tol, (name, rel.max())
        # misses are reported the same way: -1 (sphere root < 0 is returned as is), 0 for a receding plane
        same = (got == ref) | (np.isnan(got) & np.isnan(ref)) | (~hit_r & (np.abs(got - ref) <= tol * np.maximum(np.abs(ref), 1.0)))
        assert same[~hit_r].all(), name


@pytest.mark.skipif(not os.path.isdir(SPV), reason="the reference tree is not mounted here")
def test_fixtures_are_what_the_reference_binaries_compute(gold):
    """Provenance: re-runs the reference binaries for a sample of pixels and finds the committed values bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_spirv_vectors as G
    # whitted: two rows
    part = G.whitted_rows([10, 71])
    for (x, y), t in part.items():
        assert np.array_equal(np.asarray(t, np.float32).view(np.uint32), gold["whitted_texels"][y, x].view(np.uint32))
    # path: a diagonal of pixels
    h, w = gold["path_texels"].shape[:2]
    part = G.path_pixels([(x, (x * 3) % h) for x in range(0, w, 3)])
    for (x, y), (t, r, hid, ht) in part.items():
        assert np.array_equal(np.asarray(t, np.float32).view(np.uint32), gold["path_texels"][y, x].view(np.uint32))
        assert np.array_equal(np.asarray(r, np.float32).view(np.uint32), gold["path_radiance_sum"][y, x].view(np.uint32))
        assert hid == gold["path_primary_id"][y, x] and np.float32(ht) == gold["path_primary_t"][y, x]
    # the second view
    h, w = gold["path2_texels"].shape[:2]
    part = G.path_pixels([(x, (x * 5) % h) for x in range(0, w, 4)], True)
    for (x, y), (t, r, hid, ht) in part.items():
        assert np.array_equal(np.asarray(t, np.float32).view(np.uint32), gold["path2_texels"][y, x].view(np.uint32))
        assert np.array_equal(np.asarray(r, np.float32).view(np.uint32), gold["path2_radiance_sum"][y, x].view(np.uint32))
        assert hid == gold["path2_primary_id"][y, x]
    for (x, y), t in G.whitted_rows([17], True).items():
        assert np.array_equal(np.asarray(t, np.float32).view(np.uint32), gold["whitted2_texels"][y, x].view(np.uint32))


@pytest.mark.skipif(not os.path.isdir(SPV), reason="the reference tree is not mounted here")
def test_random_poses_live(oracle):
    """tests/tools/spirv_fuzz.py on 16 random cameras / lights / aspect ratios / seeds (24 random pixels each, both shaders): ids
    identical, whitted within one 8-bit step, at most a stray path pixel off (profiles/r01_spirv_fuzz.txt: 2 of 15,360)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
    import spirv_fuzz as F
    out = [F.work(k) for k in range(200, 216)]
    assert sum(o["id_bad"] for o in out) == 0 and sum(o["w_bad"] for o in out) == 0
    assert sum(o["p_bad"] for o in out) <= 2 and max(o["w_max"] for o in out) <= 1e-3


@pytest.mark.skipif(not os.path.isdir(SPV), reason="the reference tree is not mounted here")
def test_fullscreen_triangle_uv_convention():
    """Fullscreen.vert.spv: the three vertices carry uv (0,0), (2,0), (0,2) and position uv * 2 - 1, so the interpolated
    uv at the centre of framebuffer pixel (x, y) is ((x + 0.5) / W, (y + 0.5) / H), row 0 on top -- what the oracle's
    present filter and the Fullscreen.frag.spv vectors assume (Fullscreen.frag:16 then flips v)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import spirv_interp as S
    m = S.Module(os.path.join(SPV, "Fullscreen.vert.spv"))
    mc = S.Machine(m)
    vi, uv = mc.global_by_builtin(42), mc.global_cell("uv_coords")
    per_vertex = [g for g, (sc, tid, _) in m.globals.items() if sc == 3 and m.types[tid][0] == "struct"][0]
    got = []
    for i in range(3):
        vi[0] = i
        mc.run(m.entry)
        got.append((tuple(uv[0]), tuple(mc.g[per_vertex][0][0])))
    assert [g[0] for g in got] == [(0.0, 0.0), (2.0, 0.0), (0.0, 2.0)]
    assert [g[1] for g in got] == [(-1.0, -1.0, 0.0, 1.0), (3.0, -1.0, 0.0, 1.0), (-1.0, 3.0, 0.0, 1.0)]
    # position = uv * 2 - 1 is affine, so uv interpolates to (ndc + 1) / 2 = ((x + 0.5) / W, (y + 0.5) / H) at a pixel centre


def test_interpreter_covers_exactly_the_shipped_instruction_set():
    """oracle/spirv_interp.py is not a general SPIR-V implementation: it must refuse what it does not model."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import spirv_interp as S
    assert S.f32(0.1) == np.float32(0.1) and S.f32(1e39) == float("inf")
    assert S._fdiv(1.0, 0.0) == float("inf") and S._fdiv(-1.0, 0.0) == float("-inf") and np.isnan(S._fdiv(0.0, 0.0))
    assert np.isnan(S._pow(-1.0, 2.0)) and S._pow(0.0, 5.0) == 0.0 and S._sqrt(4.0) == 2.0
    assert S.unorm8(0.5) == 128 and S.unorm8(float("nan")) == 0 and S.unorm8(2.0) == 255
    for a, b in ((3.0, 7.0), (1e-30, 3.0), (16777216.0, 1.0), (0.1, 0.2)):
        assert S._fadd(S.f32(a), S.f32(b)) == np.float32(a) + np.float32(b)
        assert S._fmul(S.f32(a), S.f32(b)) == np.float32(a) * np.float32(b)
        assert S._fdiv(S.f32(a), S.f32(b)) == np.float32(a) / np.float32(b)


# ---- the CUDA path against the same vectors (transitively implied by the bit-exact GPU == oracle tests; stated) ----
@pytest.mark.gpu
def test_gpu_whitted_matches_the_reference_spirv(vk, gold):
    h, w = gold["whitted_texels"].shape[:2]
    r = vk.Renderer(w, h, spp=1, max_depth=2, integrator=vk.INTEGRATOR_WHITTED)
    r.use_default_scene(vk.SCENE_RAYTRACER)
    r.draw(_frame_data(vk, w / h, 0.0))
    acc, rgba = r.read_accum(), r.read_rgba8()
    r.close()
    _check_whitted(acc, rgba, gold)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_path_matches_the_reference_spirv(vk, gold, variant):
    h, w = gold["path_texels"].shape[:2]
    r = vk.Renderer(w, h, spp=4, max_depth=4, variant=variant, flags=vk.FLAG_HIT_IDS)
    r.use_default_scene(vk.SCENE_TRACER)
    r.set_seed(int(gold["path_seed"][0]))
    r.draw(_frame_data(vk, gold["path_aspect"][0], gold["path_frame_seed"][0]))
    acc, ids, rgba = r.read_accum(), r.read_hit_ids(), r.read_rgba8()
    r.close()
    _check_path(acc, ids, rgba, gold)


@pytest.mark.gpu
def test_gpu_matches_the_reference_spirv_from_a_second_view(vk, gold):
    h, w = gold["whitted2_texels"].shape[:2]
    r = vk.Renderer(w, h, spp=1, max_depth=2, integrator=vk.INTEGRATOR_WHITTED)
    r.use_default_scene(vk.SCENE_RAYTRACER)
    r.draw(_frame_data(vk, w / h, 0.0, gold))
    acc, rgba = r.read_accum(), r.read_rgba8()
    r.close()
    _check_whitted(acc, rgba, gold, "whitted2")
    h, w = gold["path2_texels"].shape[:2]
    for variant in (0, 1):
        r = vk.Renderer(w, h, spp=4, max_depth=4, variant=variant, flags=vk.FLAG_HIT_IDS)
        r.use_default_scene(vk.SCENE_TRACER)
        r.set_seed(int(gold["path2_seed"][0]))
        r.set_frame_index(int(gold["path2_frame_index"][0]))
        r.draw(_frame_data(vk, gold["path2_aspect"][0], gold["path2_frame_seed"][0], gold))
        acc, ids, rgba = r.read_accum(), r.read_hit_ids(), r.read_rgba8()
        r.close()
        _check_path(acc, ids, rgba, gold, "path2")
