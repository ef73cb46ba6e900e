"""The frame the target is quoted on, at its full size, against the full CPU oracle render (runs last: it is the one
GPU test whose CPU side takes more than a few seconds -- about 10 s on the GPU box's host cores, 25 s on 8 cores)."""
import numpy as np
import pytest

from helpers import apply_scene, mismatch_report, radiance_parity

pytestmark = pytest.mark.gpu


def test_target_frame_config4_full_size_against_the_oracle(vk, oracle):
    """The frame the north_star's target is quoted on -- 100k spheres, 1920x1080, 16 spp, depth 8 -- rendered once by the
    wavefront kernels and once, in full, by the CPU oracle (its LBVH, rule S): 33 M paths, 208 M rays, every pixel's
    accumulated radiance bit for bit (a NaN, should one occur, only has to be a NaN on both sides), ray counts equal.
    BASELINE.json's own bar (1e-3 relative for >= 99.9 % of the pixels, RMSE <= 1e-4) is asserted as well."""
    V = vk
    w, h = 1920, 1080
    scene = V.scenes.grid_spheres()
    fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
    r = V.Renderer(w, h, spp=16, max_depth=8, variant=V.VARIANT_WAVEFRONT, flags=V.FLAG_NO_RESOLVE)
    r.set_scene(scene); r.build_bvh(); r.set_seed(2026)
    r.draw(fd)
    acc, cnt = r.read_accum(), r.counters()
    r.close()
    sc = apply_scene(oracle, scene, fast=True).build_bvh()
    oacc, _, _, ocnt = sc.render(fd, w, h, spp=16, max_depth=8, integrator=oracle.PATH, sphere_mode=oracle.S_BVH, seed=2026,
                                 want_ids=False, want_rgba=False)
    a, b = acc.view(np.uint32), np.ascontiguousarray(oacc, dtype=np.float32).view(np.uint32)
    same = (a == b) | (np.isnan(acc) & np.isnan(oacc))
    assert same.all(), mismatch_report(acc, oacc)
    assert (cnt.closest_rays, cnt.shadow_rays, cnt.paths) == (ocnt.closest_rays, ocnt.shadow_rays, ocnt.paths)
    frac, rmse = radiance_parity(acc, oacc)
    assert frac >= 0.999 and rmse <= 1e-4
