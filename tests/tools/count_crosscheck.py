#!/usr/bin/env python
"""Cross-check of a committed GPU bench line against the CPU oracle WITHOUT a GPU: bench.py counts, on the device, the
nearest-hit and shadow rays of its K timed full-size frames (frame indices warmup .. warmup+K-1, seeds as in
bench.frame_data_for); this script renders the same frames with the oracle and compares the totals.  Ray counts are a
sensitive fingerprint of the whole computation -- one different Russian-roulette decision or hit / miss changes them.

    python tests/tools/count_crosscheck.py [profiles/r01_bench_cfg4.json]      # ~25 s per frame on 8 cores
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import oracle as O
    import vk_renderer_b200.scenes as scenes
    from vk_renderer_b200.device import default_frame_data
    from helpers import apply_scene
    import bench
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_bench_cfg4.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    cfg = d["config"]
    assert d["n_gpus"] == 1, "written for single-GPU lines"
    w, h, spp, depth, k, w0 = cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"], d["steps"], d["warmup"]
    if cfg["workload"].startswith("synthetic 100k"):          # cfg4: LBVH, rule S
        scene, mode = scenes.grid_spheres(), O.S_BVH
    elif cfg["workload"].startswith("reference default scene"):   # cfg2: the literal primitive loop
        scene, mode = scenes.tracer_default(), O.LITERAL
    else:
        raise SystemExit("no cross-check for this workload (cfg3 would take the oracle ~10 minutes per frame)")
    assert scene.digest() == cfg["scene_sha"], "the scene generator changed since the bench line was taken"
    sc = apply_scene(O, scene, fast=True)
    if mode == O.S_BVH:
        sc.build_bvh()
    closest = shadow = 0
    for i in range(w0, w0 + k):
        fd = default_frame_data(aspect_ratio=float(w) / float(h), seed=float((i * 0.61803398875) % 1.0))
        _, _, _, c = sc.render(fd, w, h, spp=spp, max_depth=depth, integrator=O.PATH, sphere_mode=mode, seed=bench.SEED,
                               frame_index=i, want_ids=False, want_rgba=False)
        closest += c.closest_rays
        shadow += c.shadow_rays
        print("frame %d: %d nearest-hit + %d shadow rays" % (i, c.closest_rays, c.shadow_rays), flush=True)
    gpu_c, gpu_s = round(d["closest_rays"] * k), round(d["shadow_rays"] * k)
    print("oracle total: %d nearest-hit, %d shadow" % (closest, shadow))
    print("GPU    total: %d nearest-hit, %d shadow   (%s)" % (gpu_c, gpu_s, os.path.relpath(path, ROOT)))
    ok = (closest, shadow) == (gpu_c, gpu_s)
    print("IDENTICAL" if ok else "DIFFERENT")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
