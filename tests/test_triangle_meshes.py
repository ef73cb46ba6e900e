"""Triangle meshes (ref: Tracer.comp:378-396 and the TODO at Raytracer.comp:10): per-triangle
materials, the OBJ loader, rule T in the oracle (CPU) and the triangles' own LBVH on the device (-m gpu)."""
import numpy as np
import pytest

from helpers import apply_scene, bits_equal, mismatch_report


def test_obj_loader_round_trip(vk, tmp_path):
    """vkrt_load_obj reads back exactly the triangles write_obj wrote (float32 round-trips through %.9g), fans polygons,
    resolves negative indices and ignores /vt/vn suffixes; a missing or malformed file is an error, not an empty mesh."""
    V = vk
    tris = V.scenes.torus_triangles(n_major=12, n_minor=8)
    p = str(tmp_path / "torus.obj")
    V.scenes.write_obj(p, tris)
    got = V.load_obj(p)
    assert got.shape == tris.shape and np.array_equal(got.view(np.uint32), tris.view(np.uint32))
    moved = V.load_obj(p, xform=[1, 0, 0, 5, 0, 1, 0, -3, 0, 0, 1, 2])
    assert np.allclose(moved[:, 0:3], tris[:, 0:3] + np.array([5, -3, 2], np.float32), atol=1e-5)
    q = str(tmp_path / "quad.obj")
    open(q, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1 4/1/1\nf -4 -3 -2\n")
    quad = V.load_obj(q)
    assert quad.shape == (3, 12)
    assert quad[0, [0, 1, 4, 5, 8, 9]].tolist() == [0, 0, 1, 0, 1, 1] and quad[1, [0, 1, 4, 5, 8, 9]].tolist() == [0, 0, 1, 1, 0, 1]
    assert np.array_equal(quad[2], quad[0])
    with pytest.raises(V.VkrtError):
        V.load_obj(str(tmp_path / "missing.obj"))
    open(q, "w").write("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(V.VkrtError):
        V.load_obj(q)


def test_rule_t_is_the_literal_loop_outside_its_tie_band(vk, oracle):
    """The oracle's rule T (order-independent, what a hierarchy can reproduce) against the reference's literal in-order
    triangle loop (Tracer.comp:378-396) on a closed mesh: primary hit ids and the whole image identical -- a torus has
    no two surfaces within EPSILON of each other along a ray except at silhouettes, and those pixels are counted."""
    V = vk
    tris = V.scenes.torus_triangles()
    scene = V.scenes.mesh_scene(tris, n_spheres=0)
    w, h = 96, 64
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.3)
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    lit = sc.render(fd, w, h, spp=2, max_depth=4, sphere_mode=oracle.LITERAL, seed=9)
    rt = sc.render(fd, w, h, spp=2, max_depth=4, sphere_mode=oracle.S_LINEAR, seed=9)
    assert (rt[1] >> 28 == 1).sum() > 150                          # the mesh is in view
    band = int((lit[1] != rt[1]).sum())
    assert band <= 2, band
    same = (lit[0].view(np.uint32) == rt[0].view(np.uint32)).all(axis=-1)
    assert same.mean() > 0.995


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_mesh_in_its_own_lbvh_bit_exact(vk, oracle, variant, tmp_path):
    """A 1,296-triangle mesh loaded from an OBJ file, per-triangle materials (diffuse, mirror, plastic, glass), among 300
    spheres: the device holds the triangles in their own LBVH (rule T) next to the spheres' (rule S) and the frame equals
    the oracle's linear scans bit for bit -- path integrator (both kernel variants) and whitted."""
    V = vk
    p = str(tmp_path / "torus.obj")
    V.scenes.write_obj(p, V.scenes.torus_triangles())
    scene = V.scenes.mesh_scene(V.load_obj(p))
    w, h = 160, 100
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.3)
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    r = V.Renderer(w, h, spp=4, max_depth=6, variant=variant, flags=V.FLAG_HIT_IDS)
    r.set_scene(scene)
    info = r.build_bvh()
    assert info.n_triangles == 1296 and info.n_tri_nodes == 1295 and 10 <= info.tri_depth <= 64
    r.set_seed(11)
    r.draw(fd)
    acc, ids, rgba, c = r.read_accum(), r.read_hit_ids(), r.read_rgba8(), r.counters()
    r.close()
    oacc, oids, orgba, oc = sc.render(fd, w, h, spp=4, max_depth=6, sphere_mode=oracle.S_BVH, seed=11)
    assert (oids >> 28 == 1).sum() > 400
    assert np.array_equal(ids, oids)
    assert bits_equal(acc, oacc), mismatch_report(acc, oacc)
    assert np.array_equal(rgba, orgba)
    assert (c.closest_rays, c.shadow_rays, c.paths) == (oc.closest_rays, oc.shadow_rays, oc.paths)
    if variant == 0:
        r = V.Renderer(w, h, spp=1, max_depth=2, integrator=V.INTEGRATOR_WHITTED, flags=V.FLAG_HIT_IDS)
        r.set_scene(scene); r.build_bvh()
        r.draw(fd)
        acc, ids = r.read_accum(), r.read_hit_ids()
        r.close()
        oacc, oids, _, _ = sc.render(fd, w, h, spp=1, max_depth=2, integrator=oracle.WHITTED, sphere_mode=oracle.S_BVH)
        assert np.array_equal(ids, oids)
        assert bits_equal(acc, oacc), mismatch_report(acc, oacc)


@pytest.mark.gpu
def test_triangle_material_argument_checks(vk):
    V = vk
    r = V.Renderer(32, 32)
    scene = V.scenes.mesh_scene(V.scenes.torus_triangles(n_major=6, n_minor=4), n_spheres=4)
    r.set_scene(scene)
    with pytest.raises(V.VkrtError):
        r.set_triangle_materials(np.zeros(5, np.uint32))                 # not one id per triangle
    r.set_triangle_materials(np.full(scene.triangles.shape[0], 999, np.uint32))
    with pytest.raises(V.VkrtError):
        r.draw(V.default_frame_data())                                   # id out of range
    r.set_triangle_materials(None)                                       # back to the shared material
    r.draw(V.default_frame_data())
    r.close()
