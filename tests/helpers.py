"""Shared helpers of the parity tests."""
import numpy as np


def bits_equal(a, b):
    """Bit-exact float comparison (NaN payloads included)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def mismatch_report(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    bad = (a.view(np.uint32) != b.view(np.uint32)).any(axis=-1)
    ys, xs = np.nonzero(bad)
    head = [(int(y), int(x), a[y, x].tolist(), b[y, x].tolist()) for y, x in list(zip(ys, xs))[:5]]
    return "%d of %d pixels differ; first: %s" % (bad.sum(), bad.size, head)


def radiance_parity(acc_gpu, acc_ref):
    """BASELINE.json's radiance criterion: fraction of pixels within 1e-3 relative, and image RMSE."""
    g = acc_gpu[..., :3] / np.maximum(acc_gpu[..., 3:4], 1.0)
    r = acc_ref[..., :3] / np.maximum(acc_ref[..., 3:4], 1.0)
    both_nan = np.isnan(g) & np.isnan(r)
    g = np.where(both_nan, 0.0, g)
    r = np.where(both_nan, 0.0, r)
    rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-3)
    ok = (rel <= 1e-3).all(axis=-1)
    rmse = float(np.sqrt(np.mean((g - r) ** 2)))
    return float(ok.mean()), rmse


def apply_scene(oracle_mod, scene, fast=False):
    """Builds an oracle scene from a vk_renderer_b200.scenes.Scene (same byte layouts)."""
    sc = oracle_mod.Scene(fast=fast)
    sc.set_materials(scene.materials)
    sc.set_spheres(scene.spheres, scene.sphere_mat)
    sc.set_planes(scene.planes, scene.plane_mat)
    sc.set_triangles(scene.triangles, scene.tri_mat)
    if getattr(scene, "tri_mats", None) is not None:
        sc.set_triangle_materials(scene.tri_mats)
    return sc
