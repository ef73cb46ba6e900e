#!/usr/bin/env python
"""Golden vectors produced by RUNNING THE REFERENCE'S OWN COMPILED SHADERS (ref: Assets/Compiled/*.spv, the
binaries the engine loads at Source/GraphicsDevice.cpp:1086-1091) through oracle/spirv_interp.py.

    python tests/golden/make_spirv_vectors.py          # needs /root/reference; writes tests/golden/spirv_vectors.npz

Contents (all float32 unless noted):
  whitted_*   Raytracer.comp.spv main() over a 128x96 image, default camera / light (Source/Main.cpp:134-141):
              the vec4 handed to imageStore for every pixel.
  path_*      Tracer.comp.spv main() over a 64x48 image (SAMPLES = DEPTH = 4 as compiled in), seed 0.25, with the
              shader's rand() replaced call site by call site by the repository's integer RNG (TracerRng): the
              imageStore vec4, the linear radiance sum (the shader's `accum` ahead of Tracer.comp:585), and the
              primary nearest hit (trace_ray() called on main()'s own Ray: primitive id and t).
  whitted2_* / path2_*  the same two shaders from a second camera (view2_camera, view2_light) that sees the host's triangle
              front-on and the open side of the room (misses); path2 at 48x36, seed 0.75, frame index 3.
  path3_* / scene3_*  Tracer.comp.spv with other scene CONSTANTS (set_tracer_scene: data only): two emissive spheres of
              different colours, a rough dielectric, a rough metal, metallic walls; 48x36, seed 0.5, frame index 1.
              scene3_materials rows = albedo[3], roughness, emissive[3], metalness, type (the vkrt_material order).
  path4_* / scene4_*  the same with TIES: two coincident spheres and one 5e-4 behind another's front -- the later sphere wins
              (t < cur + EPSILON, Tracer.comp:402); 40x30.  Planes as in scene 3.
  floathash_* Tracer.comp.spv with its own float-hash rand() left in place: mean radiance per pixel over 48 radiance()
              calls on a 24x18 image -- the distribution the integer RNG has to reproduce (not the values).
  present_*   Fullscreen.frag.spv over a 64x48 framebuffer sampling two 32x32 rgba8 images.
  kat_*       calc_sphere_intersect / calc_plane_intersect / calc_tri_intersect of Tracer.comp.spv called directly
              on 512 seeded rays each.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
SPV = "/root/reference/Assets/Compiled"
OUT = os.path.join(ROOT, "tests", "golden", "spirv_vectors.npz")

WHITTED_WH = (128, 96)
PATH_WH, PATH_SEED, PATH_FSEED, PATH_ASPECT = (64, 48), 7, 0.25, 1024.0 / 768.0
PRESENT_TEX, PRESENT_WH = 32, (64, 48)
HOST_TRIANGLE = [[(10.0, 10.0, 0.0), (0.0, 20.0, 0.0), (-10.0, 10.0, 0.0)]]      # Source/GraphicsDevice.cpp:798-803


# view 2: from inside the room towards -z, where the reference's room has no wall (Tracer.comp:204-211): the host's
# triangle front-on (Tracer.comp:348 culls its back face), three spheres, and rays that leave the scene (misses)
VIEW2 = dict(pos=(4.0, 30.0, 58.0), dir=(-0.05, -0.18, -1.0), light=(10.0, 70.0, 20.0))
WHITTED2_WH = (64, 48)
PATH2_WH, PATH2_SEED, PATH2_FSEED, PATH2_ASPECT, PATH2_FRAME = (48, 36), 11, 0.75, 4.0 / 3.0, 3


def view2_camera():
    """pos / dir / right / up of VIEW2 as float32 (an orthonormal basis; the shaders only combine them linearly)."""
    d = np.array(VIEW2["dir"], np.float64)
    d /= np.sqrt((d ** 2).sum())
    r = np.cross(d, np.array([0.0, 1.0, 0.0]))
    r /= np.sqrt((r ** 2).sum())
    u = np.cross(r, d)
    return [np.array(VIEW2["pos"], np.float32), d.astype(np.float32), r.astype(np.float32), u.astype(np.float32)]


# scene 3: the reference binary with other scene constants (oracle/spirv_interp.set_tracer_scene): TWO emissive spheres of
# different colours (the per-light loop and its RNG dimensions, Tracer.comp:458-503), a rough dielectric, a rough metal,
# metallic / coloured walls.  material = [albedo, emissive, roughness, metalness, type]
SCENE3_MATS = [
    [[1.0, 1.0, 1.0], [0.0, 0.0, 0.0], 0.58, 0.0, 1],          # 0 dielectric, eta 0.58
    [[1.0, 1.0, 1.0], [128.0, 128.0, 128.0], 0.6, 0.0, 0],     # 1 white light
    [[0.9, 0.8, 0.3], [0.0, 0.0, 0.0], 0.15, 0.9, 0],          # 2 rough gold
    [[0.8, 0.3, 0.2], [60.0, 20.0, 5.0], 0.5, 0.0, 0],         # 3 orange light
    [[0.7, 0.7, 0.8], [0.0, 0.0, 0.0], 0.2, 0.3, 0],           # 4 glossy floor
    [[0.9, 0.9, 0.9], [0.0, 0.0, 0.0], 0.8, 0.0, 0],           # 5 chalk ceiling
    [[0.2, 0.6, 0.7], [0.0, 0.0, 0.0], 0.35, 0.1, 0],          # 6 teal wall
    [[0.75, 0.25, 0.25], [0.0, 0.0, 0.0], 0.4, 0.0, 0],        # 7 matte red (the reference's)
    [[0.3, 0.3, 0.3], [0.0, 0.0, 0.0], 0.05, 1.0, 0],          # 8 dark mirror wall
    [[1.0, 0.5, 0.5], [0.0, 0.0, 0.0], 0.0, 1.0, 0],           # 9 `mirror`: what trace_ray gives every triangle (Tracer.comp:386)
]
SCENE3_SPHERES = [(0, (30.0, 18.0, 6.0), 18.0), (1, (0.0, 96.0, 0.0), 12.0), (2, (-28.0, 20.0, 18.0), 20.0), (3, (-30.0, 9.0, -40.0), 9.0)]
SCENE3_PLANES = [(4, (0.0, 1.0, 0.0), 0.0), (5, (0.0, -1.0, 0.0), 128.0), (6, (1.0, 0.0, 0.0), 64.0), (7, (0.0, 0.0, -1.0), 64.0),
                 (8, (-1.0, 0.0, 0.0), 64.0)]
PATH3_WH, PATH3_SEED, PATH3_FSEED, PATH3_ASPECT, PATH3_FRAME = (48, 36), 23, 0.5, 4.0 / 3.0, 1
# scene 4: ties.  trace_ray accepts a sphere when t < cur + EPSILON (Tracer.comp:402): a LATER sphere within 1e-3 overrides a
# nearer one.  Spheres 0 and 2 coincide (the later, gold one must win), sphere 3 sits 5e-4 behind sphere 1's front (it wins
# too, although it is farther); the default scene's materials otherwise.
SCENE4_SPHERES = [(7, (20.0, 20.0, 0.0), 14.0), (6, (-20.0, 20.0, 0.0), 14.0), (2, (20.0, 20.0, 0.0), 14.0), (3, (-20.0, 20.0, -0.0005), 14.0)]
SCENE4_MATS = SCENE3_MATS[:6] + [[[0.25, 0.75, 0.25], [0.0, 0.0, 0.0], 0.4, 0.0, 0], [[0.25, 0.25, 0.75], [0.0, 0.0, 0.0], 0.4, 0.0, 0]] + SCENE3_MATS[8:]
SCENE4_SPHERES = [(7, (20.0, 20.0, 0.0), 14.0), (6, (-20.0, 20.0, 0.0), 14.0), (2, (20.0, 20.0, 0.0), 14.0), (1, (-20.0, 20.0, -0.0005), 14.0)]
PATH4_WH, PATH4_SEED, PATH4_FSEED = (40, 30), 31, 0.125


def frame_data(aspect, seed, view2=False):
    import vk_renderer_b200.device as D          # host-side mirror only (Main.cpp:134-141 defaults); no CUDA involved
    fd = D.default_frame_data(aspect_ratio=aspect, seed=seed)
    if view2:
        cam = view2_camera()
        for name, v in zip(("pos", "dir", "right", "up"), cam):
            a = getattr(fd.camera, name)
            a.x, a.y, a.z = float(v[0]), float(v[1]), float(v[2])
        fd.light_pos.x, fd.light_pos.y, fd.light_pos.z = VIEW2["light"]
    return fd


def set_fd(S, mc, fd):
    g = lambda a: (a.x, a.y, a.z)
    S.set_frame_data(mc, fd.aspect_ratio, fd.seed, g(fd.light_pos), g(fd.camera.pos), g(fd.camera.dir), g(fd.camera.right), g(fd.camera.up))


def whitted_rows(rows, view2=False):
    import spirv_interp as S
    w, h = WHITTED2_WH if view2 else WHITTED_WH
    m = S.Module(os.path.join(SPV, "Raytracer.comp.spv"))
    mc = S.Machine(m)
    set_fd(S, mc, frame_data(w / h, 0.0, view2))
    tex = S.run_compute(mc, w, h, [(x, y) for y in rows for x in range(w)], triangles=HOST_TRIANGLE if m_has_ssbo(m) else None)
    return {k: v for k, v in tex.items()}


FULL_WH = (640, 480)                      # BASELINE.json configs[0]: Raytracer.comp, 640x480, 1 spp
OUT_FULL = os.path.join(ROOT, "tests", "golden", "spirv_config0.npz")


def whitted_full_rows(rows):
    """Raytracer.comp.spv main() over rows of the full configs[0] frame -> {(x, y): [r, g, b] as unorm8 + floats}."""
    import spirv_interp as S
    w, h = FULL_WH
    m = S.Module(os.path.join(SPV, "Raytracer.comp.spv"))
    mc = S.Machine(m)
    set_fd(S, mc, frame_data(w / h, 0.0))
    tex = S.run_compute(mc, w, h, [(x, y) for y in rows for x in range(w)])
    return {k: v[:3] for k, v in tex.items()}


def whitted2_rows(rows):
    return whitted_rows(rows, True)


def m_has_ssbo(m):
    return any(33 in m.decor.get(g, {}) and m.decor[g][33][0] == 1 for g in m.globals)


def path_pixels(pixels, view2=False, scene3=False, scene4=False):
    """-> {(x, y): (imageStore vec4, radiance sum, primary hit id, primary t)} for Tracer.comp.spv with the substituted RNG."""
    import oracle as O
    import spirv_interp as S
    w, h = PATH4_WH if scene4 else (PATH3_WH if scene3 else (PATH2_WH if view2 else PATH_WH))
    L = O.lib()
    if scene4:
        fd = frame_data(w / h, PATH4_FSEED)
        fkey = L.orc_frame_key(PATH4_SEED, fd.seed, 0)
    elif scene3:
        fd = frame_data(PATH3_ASPECT, PATH3_FSEED)
        fkey = L.orc_frame_key(PATH3_SEED, fd.seed, PATH3_FRAME)
    else:
        fd = frame_data(PATH2_ASPECT, PATH2_FSEED, True) if view2 else frame_data(PATH_ASPECT, PATH_FSEED)
        fkey = L.orc_frame_key(PATH2_SEED, fd.seed, PATH2_FRAME) if view2 else L.orc_frame_key(PATH_SEED, fd.seed, 0)
    m = S.Module(os.path.join(SPV, "Tracer.comp.spv"))
    if scene3:
        S.set_tracer_scene(m, [[SCENE3_MATS[i], list(p), r] for i, p, r in SCENE3_SPHERES],
                           [[SCENE3_MATS[i], list(n), l] for i, n, l in SCENE3_PLANES])
    if scene4:
        S.set_tracer_scene(m, [[SCENE4_MATS[i], list(p), r] for i, p, r in SCENE4_SPHERES],
                           [[SCENE4_MATS[i], list(n), l] for i, n, l in SCENE3_PLANES])
    rng = S.TracerRng(m, lambda pixel, sample, dim: L.orc_rand_u01(fkey, pixel, sample, dim))
    mc = S.Machine(m, hooks=rng.hooks())
    set_fd(S, mc, fd)
    img = S.Image(w, h)
    mc.global_by_binding(0)[0] = img
    mc.global_by_binding(1)[0] = [[[[S.f32(float(c)) for c in v] for v in t] for t in HOST_TRIANGLE]]
    gid = mc.global_by_builtin(28)
    trace_ray = m.function("trace_ray(")
    out = {}
    for (x, y) in pixels:
        rng.begin_pixel(y * w + x)
        gid[0] = [x, y, 0]
        mc.run(m.entry)
        # the primary nearest hit: trace_ray() called directly on the Ray main() built, far bound 3000 (Tracer.comp:444)
        isect = S.Pointer([[[[0.0] * 3, [0.0] * 3, 0.0, 0.0, 0], 3000.0, [0.0] * 3, [0.0] * 3]])
        found = mc.run(trace_ray, [S.Pointer([rng.primary_ray]), isect])
        hid = 0 if (scene3 or scene4) else S.tracer_hit_id(found, isect.load())       # tracer_hit_id knows the default scene's materials only
        out[(x, y)] = (img.texels[(x, y)], list(rng.radiance_sum), hid, isect.load()[1] if found else 0.0)
    return out


def path_rows(rows):
    return path_pixels([(x, y) for y in rows for x in range(PATH_WH[0])])


def path3_rows(rows):
    return path_pixels([(x, y) for y in rows for x in range(PATH3_WH[0])], scene3=True)


def path4_rows(rows):
    return path_pixels([(x, y) for y in rows for x in range(PATH4_WH[0])], scene4=True)


def path2_rows(rows):
    return path_pixels([(x, y) for y in rows for x in range(PATH2_WH[0])], True)


FH_WH, FH_SPP, FH_FSEED = (24, 18), 48, 0.37


def floathash_rows(rows):
    """Tracer.comp.spv with its OWN rand() (the float hash, Tracer.comp:221-234), nothing substituted: main() runs once
    per pixel to set rand_salt / coords and build the primary Ray (its four radiance() calls are skipped), then
    radiance() is called FH_SPP times.  -> {(x, y): mean radiance, clamped at 0}.  Column 0 is skipped: uv.y / uv.x is
    inf or NaN there and so is every rand()."""
    import spirv_interp as S
    w, h = FH_WH
    m = S.Module(os.path.join(SPV, "Tracer.comp.spv"))
    rad = m.function("radiance(")
    mc = S.Machine(m)
    set_fd(S, mc, frame_data(w / h, FH_FSEED))
    mc.global_by_binding(0)[0] = S.Image(w, h)
    mc.global_by_binding(1)[0] = [[[[S.f32(float(c)) for c in v] for v in t] for t in HOST_TRIANGLE]]
    gid = mc.global_by_builtin(28)
    cap = {}

    def capture(machine, args, site):
        cap["ray"] = args[0].load()
        return [0.0, 0.0, 0.0]
    mc.hooks[rad] = capture
    out = {}
    for y in rows:
        for x in range(1, w):
            gid[0] = [x, y, 0]
            mc.run(m.entry)
            acc = np.zeros(3)
            for _ in range(FH_SPP):
                r = np.array(mc.run(rad, [S.Pointer([cap["ray"]])]), np.float64)
                acc += np.where(np.isnan(r), 0.0, np.maximum(r, 0.0))
            out[(x, y)] = acc / FH_SPP
    return out


def present_inputs():
    rs = np.random.RandomState(11)
    t = PRESENT_TEX
    yy, xx = np.mgrid[0:t, 0:t]
    base = np.stack([(xx * 8) % 256, (yy * 8) % 256, ((xx + yy) * 4) % 256, np.full_like(xx, 255)], -1).astype(np.uint8)
    a = base.copy()
    b = base.copy()
    b[8:20, 6:26, :3] = rs.randint(0, 256, size=(12, 20, 3))       # a region where the temporal-variance gate opens
    a[::5, ::7, :3] = rs.randint(0, 256, size=a[::5, ::7, :3].shape)
    return a, b


def present_all():
    import spirv_interp as S
    a, b = present_inputs()
    m = S.Module(os.path.join(SPV, "Fullscreen.frag.spv"))
    mc = S.Machine(m)
    mc.global_by_binding(0)[0] = S.bilinear_sampler(a.tolist())
    mc.global_by_binding(1)[0] = S.bilinear_sampler(b.tolist())
    uv_cell, out_cell = mc.global_cell("uv_coords"), mc.global_cell("frag_color")
    w, h = PRESENT_WH
    out = np.zeros((h, w, 4), np.float32)
    for y in range(h):
        for x in range(w):
            # Fullscreen.vert:8-10 interpolated at the pixel centre
            uv_cell[0] = [S._fdiv(S.f32(x + 0.5), float(w)), S._fdiv(S.f32(y + 0.5), float(h))]
            mc.run(m.entry)
            out[y, x] = out_cell[0]
    return a, b, out


def kats():
    import spirv_interp as S
    m = S.Module(os.path.join(SPV, "Tracer.comp.spv"))
    mc = S.Machine(m)
    rs = np.random.RandomState(5)
    n = 512
    o = rs.uniform(-60, 60, size=(n, 3)).astype(np.float32)
    o[:, 1] = rs.uniform(1, 120, size=n).astype(np.float32)
    mat = [[0.5, 0.5, 0.5], [0.0, 0.0, 0.0], 0.4, 0.0, 0]
    sph = np.concatenate([rs.uniform(-50, 50, size=(n, 3)), rs.uniform(2, 25, size=(n, 1))], 1).astype(np.float32)
    tri = rs.uniform(-30, 30, size=(n, 3, 3)).astype(np.float32)
    # rays aimed near their sphere (offset up to 1.3 r: hits, grazing rays, misses; some origins inside) or, for the
    # second half, at a point of their triangle's plane (barycentrics in [-0.3, 1.3]: inside and just outside)
    tgt = sph[:, :3] + rs.normal(size=(n, 3)) * (sph[:, 3:4] * rs.uniform(0, 1.3, size=(n, 1)) / np.sqrt(3.0))
    bu, bv = rs.uniform(-0.3, 1.3, size=(n, 1)), rs.uniform(-0.3, 1.3, size=(n, 1))
    ttgt = tri[:, 0] + bu * (tri[:, 1] - tri[:, 0]) + bv * (tri[:, 2] - tri[:, 0])
    tgt[n // 2:] = ttgt[n // 2:]
    o[::9] = (sph[::9, :3] + 0.3 * sph[::9, 3:4]).astype(np.float32)          # inside the sphere: near root < 0 (Tracer.comp:326)
    d = (tgt - o).astype(np.float64)
    d = (d / np.sqrt((d ** 2).sum(1, keepdims=True))).astype(np.float32)
    # Tracer.comp:348 culls back faces (det < EPSILON): turn three of four aimed-at triangles towards their ray
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    back = (nrm * d).sum(1) > 0
    back[::4] = False
    back[:n // 2] = False
    tri[back] = tri[back][:, [0, 2, 1]]
    pln = rs.normal(size=(n, 4)).astype(np.float32)
    pln[:, :3] = (pln[:, :3] / np.sqrt((pln[:, :3].astype(np.float64) ** 2).sum(1, keepdims=True))).astype(np.float32)
    pln[:, 3] = rs.uniform(-64, 128, size=n).astype(np.float32)
    pln[::16, :3] = np.array([0, 1, 0], np.float32)                 # axis-aligned planes like the room's
    d[8::16, 1] = 0.0                                               # ... with rays parallel to them (Tracer.comp:337)
    f_s, f_p, f_t = m.function("calc_sphere_intersect("), m.function("calc_plane_intersect("), m.function("calc_tri_intersect(")
    P = lambda v: S.Pointer([v])
    F = lambda a: [float(x) for x in a]
    t_s, t_p, t_t = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for i in range(n):
        ray = [F(o[i]), F(d[i])]
        t_s[i] = mc.run(f_s, [P(ray), P([mat, F(sph[i, :3]), float(sph[i, 3])])])
        t_p[i] = mc.run(f_p, [P(ray), P([mat, F(pln[i, :3]), float(pln[i, 3])])])
        t_t[i] = mc.run(f_t, [P(ray), P([F(tri[i, 0]), F(tri[i, 1]), F(tri[i, 2])])])
    return dict(kat_o=o, kat_d=d, kat_sphere=sph, kat_plane=pln, kat_tri=tri, kat_t_sphere=t_s, kat_t_plane=t_p, kat_t_tri=t_t)


def main():
    assert os.path.isdir(SPV), "the reference tree is not mounted"
    import oracle as O
    O.build()
    n = max(1, (os.cpu_count() or 2) - 0)
    with mp.Pool(n) as pool:
        w, h = WHITTED_WH
        chunks = [list(range(y, min(y + 4, h))) for y in range(0, h, 4)]
        whitted = np.zeros((h, w, 4), np.float32)
        for part in pool.imap_unordered(whitted_rows, chunks):
            for (x, y), t in part.items():
                whitted[y, x] = t
        w, h = PATH_WH
        chunks = [[y] for y in range(h)]
        path_tex, path_rad = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 3), np.float32)
        path_id, path_t = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
        for part in pool.imap_unordered(path_rows, chunks):
            for (x, y), (t, r, hid, ht) in part.items():
                path_tex[y, x], path_rad[y, x], path_id[y, x], path_t[y, x] = t, r, hid, ht
        w, h = WHITTED2_WH
        whitted2 = np.zeros((h, w, 4), np.float32)
        for part in pool.imap_unordered(whitted2_rows, [list(range(y, min(y + 4, h))) for y in range(0, h, 4)]):
            for (x, y), t in part.items():
                whitted2[y, x] = t
        w, h = PATH2_WH
        path2_tex, path2_rad = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 3), np.float32)
        path2_id, path2_t = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
        for part in pool.imap_unordered(path2_rows, [[y] for y in range(h)]):
            for (x, y), (t, r, hid, ht) in part.items():
                path2_tex[y, x], path2_rad[y, x], path2_id[y, x], path2_t[y, x] = t, r, hid, ht
        w, h = PATH3_WH
        path3_tex, path3_rad, path3_t = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.float32)
        for part in pool.imap_unordered(path3_rows, [[y] for y in range(h)]):
            for (x, y), (t, r, _, ht) in part.items():
                path3_tex[y, x], path3_rad[y, x], path3_t[y, x] = t, r, ht
        w, h = PATH4_WH
        path4_tex, path4_rad = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 3), np.float32)
        for part in pool.imap_unordered(path4_rows, [[y] for y in range(h)]):
            for (x, y), (t, r, _, _) in part.items():
                path4_tex[y, x], path4_rad[y, x] = t, r
        w, h = FH_WH
        fh = np.zeros((h, w, 3), np.float32)
        for part in pool.imap_unordered(floathash_rows, [[y] for y in range(h)]):
            for (x, y), v in part.items():
                fh[y, x] = v
    a, b, present = present_all()
    k = kats()
    np.savez_compressed(OUT, whitted_texels=whitted, path_texels=path_tex, path_radiance_sum=path_rad,
                        path_primary_id=path_id, path_primary_t=path_t,
                        view2_camera=np.stack(view2_camera()), view2_light=np.array(VIEW2["light"], np.float32),
                        whitted2_texels=whitted2, path2_texels=path2_tex, path2_radiance_sum=path2_rad,
                        path2_primary_id=path2_id, path2_primary_t=path2_t, path2_seed=np.array([PATH2_SEED]),
                        path2_frame_seed=np.array([PATH2_FSEED], np.float32), path2_aspect=np.array([PATH2_ASPECT], np.float32),
                        path2_frame_index=np.array([PATH2_FRAME]),
                        path3_texels=path3_tex, path3_radiance_sum=path3_rad, path3_primary_t=path3_t, path3_seed=np.array([PATH3_SEED]),
                        path3_frame_seed=np.array([PATH3_FSEED], np.float32), path3_aspect=np.array([PATH3_ASPECT], np.float32),
                        path3_frame_index=np.array([PATH3_FRAME]),
                        scene3_materials=np.array([m[0] + [m[2]] + m[1] + [m[3], float(m[4])] for m in SCENE3_MATS], np.float32),
                        scene3_spheres=np.array([list(p) + [r] for _, p, r in SCENE3_SPHERES], np.float32),
                        scene3_sphere_mat=np.array([i for i, _, _ in SCENE3_SPHERES], np.uint32),
                        scene3_planes=np.array([list(n) + [l] for _, n, l in SCENE3_PLANES], np.float32),
                        scene3_plane_mat=np.array([i for i, _, _ in SCENE3_PLANES], np.uint32),
                        path4_texels=path4_tex, path4_radiance_sum=path4_rad, path4_seed=np.array([PATH4_SEED]),
                        path4_frame_seed=np.array([PATH4_FSEED], np.float32),
                        scene4_materials=np.array([m[0] + [m[2]] + m[1] + [m[3], float(m[4])] for m in SCENE4_MATS], np.float32),
                        scene4_spheres=np.array([list(p) + [r] for _, p, r in SCENE4_SPHERES], np.float32),
                        scene4_sphere_mat=np.array([i for i, _, _ in SCENE4_SPHERES], np.uint32),
                        floathash_mean=fh, floathash_spp=np.array([FH_SPP]), floathash_frame_seed=np.array([FH_FSEED], np.float32),
                        path_seed=np.array([PATH_SEED]), path_frame_seed=np.array([PATH_FSEED], np.float32),
                        path_aspect=np.array([PATH_ASPECT], np.float32),
                        present_binding0=a, present_binding1=b, present_color=present, **k)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def main_full():
    """python tests/golden/make_spirv_vectors.py --full : BASELINE configs[0] at its full size (a few minutes)."""
    assert os.path.isdir(SPV), "the reference tree is not mounted"
    w, h = FULL_WH
    img = np.zeros((h, w, 3), np.float32)
    with mp.Pool(os.cpu_count() or 2) as pool:
        for part in pool.imap_unordered(whitted_full_rows, [list(range(y, min(y + 2, h))) for y in range(0, h, 2)]):
            for (x, y), t in part.items():
                img[y, x] = t
    u8 = np.floor(np.clip(np.where(np.isnan(img), 0.0, img), 0.0, 1.0).astype(np.float32) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    # the 8-bit image in full, the float colours as binary16 of (value * 255 - unorm8) would not compress: keep every 4th row / column exactly
    np.savez_compressed(OUT_FULL, whitted_full_rgb8=u8, whitted_full_float_4x4=img[::4, ::4].copy())
    print("wrote", OUT_FULL, os.path.getsize(OUT_FULL), "bytes")


if __name__ == "__main__":
    main_full() if "--full" in sys.argv else main()
