"""CPU tests of the oracle (the CPU restatement of Assets/Tracer.comp / Raytracer.comp): known-answer
tests per routine incl. the NaN/inf edge cases of SURVEY.md 8a, the survey's sanity values, struct
layouts, and the golden fixtures."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from helpers import apply_scene, bits_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def f3(*v):
    return (C.c_float * 3)(*v)


def ulps(a, b):
    a = np.float32(a).view(np.int32).astype(np.int64)
    b = np.float32(b).view(np.int32).astype(np.int64)
    return abs(int(a) - int(b))


# ---- vkrt-f32 transcendental routines against libm ---------------------------------------------
def test_sin_cos_against_libm(oracle):
    l = oracle.lib()
    xs = np.linspace(0.0, 2.0 * math.pi, 4001, dtype=np.float32)       # the shader's arguments: 2*PI*rand()
    s = np.array([l.orc_sin(float(x)) for x in xs])
    c = np.array([l.orc_cos(float(x)) for x in xs])
    assert np.max(np.abs(s - np.sin(xs.astype(np.float64)))) < 2e-7
    assert np.max(np.abs(c - np.cos(xs.astype(np.float64)))) < 2e-7
    for x in (-7.5, 100.25, 1000.0):                                   # outside the primary range too
        assert abs(l.orc_sin(x) - math.sin(np.float32(x))) < 1e-5
        assert abs(l.orc_cos(x) - math.cos(np.float32(x))) < 1e-5


def test_exp2_log2_pow_against_libm(oracle):
    l = oracle.lib()
    for x in np.linspace(-125, 127, 2001, dtype=np.float32):
        assert abs(l.orc_exp2(float(x)) / 2.0 ** float(x) - 1.0) < 3e-7
    for x in np.logspace(-37, 38, 2001).astype(np.float32):
        ref = math.log2(float(x))
        assert abs(l.orc_log2(float(x)) - ref) <= 3e-7 * max(1.0, abs(ref))
    # the call sites: pow(1 - cos, 5) (:271,276), pow(x, 1/2.2) (:588), pow(d + 1, 2) (:444), pow(x, 16) (Raytracer :315)
    for x in np.linspace(0.0, 1.0, 501, dtype=np.float32)[1:]:
        assert abs(l.orc_pow(float(x), 5.0) / float(x) ** 5 - 1.0) < 5e-6
        assert abs(l.orc_pow(float(x), 1.0 / 2.2) / float(x) ** (1.0 / 2.2) - 1.0) < 2e-6
        assert abs(l.orc_pow(float(x), 16.0) - float(x) ** 16) < 2e-5
    assert l.orc_pow(0.0, 5.0) == 0.0 and l.orc_pow(0.0, 1.0 / 2.2) == 0.0
    assert l.orc_pow(1.0, 2.0) == 1.0 and l.orc_pow(2.0, 2.0) == 4.0 and l.orc_pow(4.0, 2.0) == 16.0
    assert math.isnan(l.orc_pow(-1e-7, 5.0))                      # GLSL: undefined for x < 0; here NaN
    assert l.orc_exp2(200.0) == math.inf and l.orc_exp2(-200.0) == 0.0
    assert l.orc_log2(0.0) == -math.inf and math.isnan(l.orc_log2(-1.0))
    assert abs(l.orc_log2(1e-40) - math.log2(np.float32(1e-40))) < 1e-4   # subnormal input


def test_pcg_rng_known_answers(oracle):
    l = oracle.lib()

    def pcg(v):
        state = (v * 747796405 + 2891336453) & 0xFFFFFFFF
        word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
        return ((word >> 22) ^ word) & 0xFFFFFFFF

    for v in (0, 1, 2, 12345, 0xFFFFFFFF, 0x9E3779B9):
        assert l.orc_pcg_hash(v) == pcg(v)
    key = l.orc_frame_key(2026, 0.5, 3)
    a = pcg((3 + 0x9E3779B9) & 0xFFFFFFFF)
    b = pcg(np.float32(0.5).view(np.uint32).item() ^ a)
    c = pcg(0 ^ b)
    assert key == pcg(2026 ^ c)
    us = np.array([l.orc_rand_u01(key, 77, s, d) for s in range(64) for d in range(64)])
    assert us.min() >= 0.0 and us.max() < 1.0 and abs(us.mean() - 0.5) < 0.02
    assert len(np.unique(us)) > 4000
    assert l.orc_rand_u01(key, 77, 1, 2) != l.orc_rand_u01(key, 77, 2, 1) != l.orc_rand_u01(key, 78, 1, 2)


# ---- intersection routines ------------------------------------------------------------------------
def test_sphere_intersect_kats(oracle):
    l = oracle.lib()
    s = oracle.Sphere(0, 0, 10, 2)
    assert l.orc_sphere_intersect(f3(0, 0, 0), f3(0, 0, 1), C.byref(s)) == 8.0           # near root only (:326)
    assert l.orc_sphere_intersect(f3(0, 5, 0), f3(0, 0, 1), C.byref(s)) == -1.0          # h < 0
    assert l.orc_sphere_intersect(f3(0, 0, 10), f3(0, 0, 1), C.byref(s)) == -2.0         # origin inside: negative near root
    assert l.orc_sphere_intersect(f3(0, 0, 20), f3(0, 0, 1), C.byref(s)) == -12.0        # behind
    assert l.orc_sphere_intersect(f3(0, 2, 0), f3(0, 0, 1), C.byref(s)) == 10.0          # tangent: h == 0
    assert math.isnan(l.orc_sphere_intersect(f3(0, 0, 0), f3(math.nan, 0, 1), C.byref(s)))


def test_plane_intersect_kats(oracle):
    l = oracle.lib()
    floor = oracle.Plane(0, 1, 0, 0)
    for fn in (l.orc_plane_intersect_tracer, l.orc_plane_intersect_raytracer):
        assert fn(f3(0, 5, 0), f3(0, -1, 0), C.byref(floor)) == 5.0
        assert fn(f3(0, 5, 0), f3(0, 1, 0), C.byref(floor)) == 0.0                        # behind -> max(dist, 0) = 0
    # parallel ray: Tracer multiplies by when_neq(d, 0) (:337) -> 0 * inf = NaN or 0; either way it fails t > EPSILON
    t = l.orc_plane_intersect_tracer(f3(0, 5, 0), f3(1, 0, 0), C.byref(floor))
    assert math.isnan(t) or t == 0.0
    assert not (t > 1e-3)
    t0 = l.orc_plane_intersect_tracer(f3(0, 0, 0), f3(1, 0, 0), C.byref(floor))           # 0/0
    assert not (t0 > 1e-3)
    assert l.orc_plane_intersect_raytracer(f3(0, 5, 0), f3(1, 0, 0), C.byref(floor)) == 0.0   # literal 0.0 (Raytracer :184-187)
    wall = oracle.Plane(1, 0, 0, 64)
    assert l.orc_plane_intersect_tracer(f3(0, 0, 0), f3(-1, 0, 0), C.byref(wall)) == 64.0


def test_triangle_intersect_kats(oracle):
    l = oracle.lib()
    tri = (C.c_float * 12)(10, 10, 0, 0, 0, 20, 0, 0, -10, 10, 0, 0)        # the host's triangle (GraphicsDevice.cpp:798-803)
    front = l.orc_tri_intersect(f3(0, 12, 10), f3(0, 0, -1), tri, 1e-3)
    back = l.orc_tri_intersect(f3(0, 12, -10), f3(0, 0, 1), tri, 1e-3)
    # exactly one side survives the `determinant < EPSILON` back-face cull (:348)
    assert sorted([front, back])[0] == -1.0 and sorted([front, back])[1] == 10.0
    hit_o, hit_d = (f3(0, 12, 10), f3(0, 0, -1)) if front > 0 else (f3(0, 12, -10), f3(0, 0, 1))
    assert l.orc_tri_intersect(f3(50, 12, hit_o[2]), hit_d, tri, 1e-3) == -1.0          # u out of range
    assert l.orc_tri_intersect(f3(0, 25, hit_o[2]), hit_d, tri, 1e-3) == -1.0           # v / u+v out of range
    assert l.orc_tri_intersect(f3(0, 12, hit_o[2]), f3(1, 0, 0), tri, 1e-3) == -1.0     # parallel: det = 0 < EPSILON


def test_slab_test_is_monotone_under_inclusion(oracle):
    """Rule S relies on: a box that contains another is hit whenever the inner one is, with tn' <= tn."""
    l = oracle.lib()
    rng = np.random.default_rng(5)
    tn, tf, tn2, tf2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    hits = 0
    for _ in range(20000):
        lo = rng.uniform(-50, 50, 3).astype(np.float32)
        hi = (lo + rng.uniform(0.01, 20, 3)).astype(np.float32)
        lo2 = (lo - rng.uniform(0, 5, 3)).astype(np.float32)
        hi2 = (hi + rng.uniform(0, 5, 3)).astype(np.float32)
        o = rng.uniform(-80, 80, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if rng.random() < 0.7:                                          # aim at (or just past) the inner box
            d = (lo + (hi - lo) * rng.uniform(-0.1, 1.1, 3) - o).astype(np.float32)
        if rng.random() < 0.2:
            d[rng.integers(0, 3)] = 0.0                                 # axis-parallel rays (safe_inv path)
        d /= np.linalg.norm(d)
        h1 = l.orc_slab(f3(*o), f3(*d), f3(*lo), f3(*hi), C.byref(tn), C.byref(tf))
        h2 = l.orc_slab(f3(*o), f3(*d), f3(*lo2), f3(*hi2), C.byref(tn2), C.byref(tf2))
        if h1:
            hits += 1
            assert h2 and tn2.value <= tn.value and tf2.value >= tf.value
        assert math.isfinite(tn.value) and math.isfinite(tf.value)
    assert hits > 500


# ---- layouts and the survey's sanity values ---------------------------------------------------------
def test_struct_layouts(vk):
    L = vk._lib
    assert C.sizeof(L.CameraData) == 64 and C.sizeof(L.FrameData) == 96 and C.sizeof(L.Triangle) == 48
    assert (L.FrameData.aspect_ratio.offset, L.FrameData.seed.offset, L.FrameData.light_pos.offset,
            L.FrameData.camera.offset) == (0, 4, 16, 32)
    assert (L.CameraData.pos.offset, L.CameraData.dir.offset, L.CameraData.right.offset, L.CameraData.up.offset) == (0, 16, 32, 48)
    assert (L.Triangle.v0.offset, L.Triangle.v1.offset, L.Triangle.v2.offset) == (0, 16, 32)


def test_reference_struct_layouts_if_ref_built(oracle):
    rl = oracle.ref_camera_lib()
    if rl is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    assert rl.ref_sizeof_frame_data() == 96 and rl.ref_sizeof_triangle() == 48
    assert (rl.ref_offsetof_seed(), rl.ref_offsetof_light_pos(), rl.ref_offsetof_camera()) == (4, 16, 32)


def test_default_camera_matches_reference_vector(vk):
    """tests/golden/camera_poses.json was produced by the reference's own Camera.cpp."""
    gold = json.load(open(os.path.join(GOLD, "camera_poses.json")))
    fd = vk.default_frame_data()
    got = np.frombuffer(bytes(fd), dtype=np.float32)[8:24].reshape(4, 4)[:, :3]
    assert bits_equal(got, np.array(gold["default_view"]["camera"], dtype=np.float32))
    assert [fd.light_pos.x, fd.light_pos.y, fd.light_pos.z] == gold["default_view"]["light_pos"]
    # SURVEY.md section 4 prints the same basis to 9 digits
    assert np.allclose(got[1], [-0.455874413, 0.0775888711, 0.886655807], atol=1e-9)
    assert np.allclose(got[2], [0.889336705, 0.0, 0.4572528], atol=1e-8)


def test_camera_mirror_matches_reference_poses(vk):
    gold = json.load(open(os.path.join(GOLD, "camera_poses.json")))
    ops = ["move_forward", "move_backward", "move_left", "move_right", "move_up", "move_down"]
    for p in gold["poses"]:
        cam = vk.Camera()
        cam.pos = np.array(p["pos"], dtype=np.float32)
        cam.pitch, cam.yaw = np.float32(p["pitch"]), np.float32(p["yaw"])
        cam.update()
        got = np.stack([cam.pos, cam.dir, cam.right, cam.up])
        assert bits_equal(got, np.array(p["camera"], dtype=np.float32))
        getattr(cam, ops[p["move_op"]])(p["move_speed"])
        got = np.stack([cam.pos, cam.dir, cam.right, cam.up])
        assert bits_equal(got, np.array(p["camera_after_move"], dtype=np.float32))


def test_camera_mirror_matches_live_reference_if_built(vk, oracle):
    rl = oracle.ref_camera_lib()
    if rl is None:
        pytest.skip("oracle/_ref not built")
    ref = np.zeros(24, dtype=np.float32)
    rl.ref_default_frame_data(ref.ctypes.data_as(C.c_void_p))
    mine = np.frombuffer(bytes(vk.default_frame_data()), dtype=np.float32)
    assert bits_equal(mine[8:24].reshape(4, 4)[:, :3], ref[8:24].reshape(4, 4)[:, :3])
    assert bits_equal(mine[4:7], ref[4:7])


def _histogram(ids):
    sub = ids[::8, ::8]
    vals, counts = np.unique(sub, return_counts=True)
    return {int(v): 100.0 * c / sub.size for v, c in zip(vals, counts)}


def test_primary_hit_histograms_match_survey(vk, oracle):
    """default camera, aspect 1024/768, every 8th pixel of 1024x1024."""
    fd = vk.default_frame_data(aspect_ratio=1024.0 / 768.0)
    sc = oracle.Scene().use_default(oracle.SCENE_RAYTRACER)
    _, ids, _, _ = sc.render(fd, 1024, 1024, spp=1, max_depth=2, integrator=oracle.WHITTED, want_rgba=False)
    h = _histogram(ids)
    P, S = 3 << 28, 2 << 28
    expect = {P | 0: 29.3, P | 1: 8.0, P | 2: 19.0, P | 3: 33.5, P | 4: 8.8, S | 0: 0.13, S | 1: 1.26}
    for k, v in expect.items():
        assert abs(h.get(k, 0.0) - v) < 0.06, (hex(k), h.get(k), v)
    assert (1 << 28) not in h and 0 not in h                       # triangle culled from this view, no misses
    assert ids[512, 512] == (P | 2)                                # centre pixel: plane 2

    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    _, ids, _, _ = sc.render(fd, 1024, 1024, spp=1, max_depth=1, integrator=oracle.PATH, want_rgba=False)
    h = _histogram(ids)
    expect = {P | 0: 26.6, P | 1: 7.1, P | 2: 31.7, P | 3: 15.9, P | 4: 8.6, S | 0: 3.4, S | 1: 1.25, S | 2: 3.0, S | 3: 2.5}
    for k, v in expect.items():
        assert abs(h.get(k, 0.0) - v) < 0.06, (hex(k), h.get(k), v)
    assert ids[512, 512] == (S | 2)                                # centre pixel: sphere 2 (mirror)
    o, d = f3(0, 0, 0), f3(0, 0, 0)
    oracle.lib().orc_primary_ray(np.frombuffer(bytes(fd), dtype=np.uint8).ctypes.data_as(C.c_void_p), 1024, 1024, 512, 512, o, d)
    i, t = sc.query_spheres(oracle.LITERAL, list(o), list(d), 3000.0 + 1e-3)
    assert i == 2 and abs(t - 126.482) < 2e-3


def test_oracle_matches_golden_frames(vk, oracle):
    from golden.make_golden import FRAMES, scene_for
    gold = np.load(os.path.join(GOLD, "frames.npz"))
    for name, (scn, w, h, spp, depth, integ, mode, seed, fseed) in FRAMES.items():
        sc = apply_scene(oracle, scene_for(scn))
        if mode == oracle.S_BVH:
            sc.build_bvh()
        fd = gold[name + ".frame_data"].tobytes()
        assert fd == bytes(vk.default_frame_data(aspect_ratio=w / h, seed=fseed))
        acc, ids, rgba, cnt = sc.render(fd, w, h, spp=spp, max_depth=depth, integrator=integ, sphere_mode=mode, seed=seed)
        assert bits_equal(acc, gold[name + ".accum"]), name
        assert np.array_equal(ids, gold[name + ".ids"]) and np.array_equal(rgba, gold[name + ".rgba"])
        assert [cnt.closest_rays, cnt.shadow_rays, cnt.paths] == gold[name + ".counts"].tolist()
        # the -O3 timing build of the oracle gives the same bits
        fast = apply_scene(oracle, scene_for(scn), fast=True)
        if mode == oracle.S_BVH:
            fast.build_bvh()
        facc, _, _, _ = fast.render(fd, w, h, spp=spp, max_depth=depth, integrator=integ, sphere_mode=mode, seed=seed)
        assert bits_equal(facc, acc), name


def test_oracle_semantic_details(vk, oracle):
    """Appendix A pitfalls that are easy to get wrong."""
    fd = vk.default_frame_data(aspect_ratio=4.0 / 3.0, seed=0.5)
    sc = oracle.Scene().use_default(oracle.SCENE_TRACER)
    w, h = 64, 48
    a1, ids, _, c1 = sc.render(fd, w, h, spp=4, max_depth=4, seed=1)
    # (5) every sample of a pixel shares one primary ray: the first closest-hit count is spp per pixel
    assert c1.paths == w * h * 4 and c1.closest_rays >= c1.paths
    # RNG is counter based: thread count and rectangles cannot change the image
    a2, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=1, n_threads=1)
    assert bits_equal(a1, a2)
    a3, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=1, rect=(10, 5, 30, 25))
    assert bits_equal(a3[5:25, 10:30], a1[5:25, 10:30]) and not a3[0:5].any()
    # sample ranges add up in order
    lo, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=1, samples=(0, 2), want_ids=False)
    both, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=1, samples=(2, 4), accum=lo, want_ids=False)
    assert np.all(both[..., 3] == 4.0) and np.allclose(both, a1, rtol=1e-5, atol=1e-6)
    # a different seed / frame index / FrameData.seed gives a different image
    for kw in ({"seed": 2}, {"frame_index": 1}):
        b, _, _, _ = sc.render(fd, w, h, spp=4, max_depth=4, **{"seed": 1, **kw})
        assert not bits_equal(a1, b)
    fd2 = vk.default_frame_data(aspect_ratio=4.0 / 3.0, seed=0.25)
    b, _, _, _ = sc.render(fd2, w, h, spp=4, max_depth=4, seed=1)
    assert not bits_equal(a1, b)
    # depth 1: no bounce light, radiance is direct light + emission only, and no NaNs on this view
    d1, _, _, cnt = sc.render(fd, w, h, spp=1, max_depth=1, seed=1)
    assert cnt.closest_rays == w * h and np.isfinite(d1).all()
    # resolve: Reinhard + gamma + dither keeps every byte in range and alpha = 255
    _, _, rgba, _ = sc.render(fd, w, h, spp=4, max_depth=4, seed=1)
    assert rgba[..., 3].min() == 255 and rgba[..., :3].max() > 100


def test_whitted_miss_keeps_retracing(vk, oracle):
    """Raytracer.comp:384-394: on a miss `bounce_count` is not advanced, so the loop re-traces the ray."""
    sc = oracle.Scene()
    sc.set_materials(vk.scenes.raytracer_default().materials)
    sc.set_spheres(np.zeros((0, 4), np.float32), np.zeros(0, np.uint32))
    sc.set_planes(np.zeros((0, 4), np.float32), np.zeros(0, np.uint32))
    sc.set_triangles(np.zeros((0, 12), np.float32), 0)
    fd = vk.default_frame_data()
    acc, ids, rgba, cnt = sc.render(fd, 8, 8, spp=1, max_depth=2, integrator=oracle.WHITTED)
    assert cnt.closest_rays == 8 * 8 * 3 and cnt.shadow_rays == 0 and not acc[..., :3].any() and not ids.any()
