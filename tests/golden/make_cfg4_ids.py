#!/usr/bin/env python
"""Primary nearest-hit primitive ids of BASELINE.json configs[3] at its FULL size (100k-sphere scene, 1920x1080),
computed by the CPU oracle under rule S through its own LBVH (DESIGN.md "Rule S"), and the size of the band of
grazing ties in which the reference's literal in-order loop (Tracer.comp:398-412) names another sphere.

    python tests/golden/make_cfg4_ids.py        # ~2 min on 8 cores; writes tests/golden/cfg4_primary_ids.npz

The ids are RNG-independent (every sample of a pixel shares one primary ray, Tracer.comp:574-581).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "cfg4_primary_ids.npz")


def main():
    import oracle as O
    import vk_renderer_b200.scenes as scenes
    from vk_renderer_b200.device import default_frame_data
    from helpers import apply_scene
    w, h = 1920, 1080
    scene = scenes.grid_spheres()
    sc = apply_scene(O, scene, fast=True)
    sc.build_bvh()
    fd = default_frame_data(aspect_ratio=w / h, seed=0.5)
    _, ids, _, c = sc.render(fd, w, h, spp=1, max_depth=1, sphere_mode=O.S_BVH, seed=2026)
    np.savez_compressed(OUT, ids=ids, tie_band=np.array([c.literal_vs_s_mismatch]), scene_sha=np.array([scene.digest()]))
    print("wrote", OUT, os.path.getsize(OUT), "bytes; tie band", c.literal_vs_s_mismatch, "of", ids.size)


if __name__ == "__main__":
    main()
