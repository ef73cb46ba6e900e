#!/bin/bash
# usage (GPU box): tools/sanitize.sh <tag>  -- compute-sanitizer memcheck / racecheck / synccheck over small frames of every kernel family
T=${1:-san}; O=gpurun_out; mkdir -p $O
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import vk_renderer_b200 as V
fd = V.default_frame_data(aspect_ratio=1.5, seed=0.3)
def run(scene, w, h, spp, depth, variant, bvh, integrator=V.INTEGRATOR_PATH, frames=2, **kw):
    r = V.Renderer(w, h, spp=spp, max_depth=depth, variant=variant, integrator=integrator, flags=V.FLAG_HIT_IDS, **kw)
    r.set_scene(scene)
    if bvh: r.build_bvh()
    for i in range(frames):
        r.set_frame_index(i); r.draw(fd)
    a = r.read_accum(); r.read_rgba8(); r.close()
    return a
mesh = V.scenes.mesh_scene(V.scenes.torus_triangles(n_major=10, n_minor=6), n_spheres=60)
for variant in (0, 1):
    run(V.scenes.tracer_default(), 48, 32, 4, 4, variant, False)
    run(V.scenes.random_spheres(300), 60, 40, 3, 6, variant, True)
    run(V.scenes.random_spheres(300), 60, 40, 20, 4, variant, True, tile_shard=(1, 3))      # two waves, a tile shard
    run(mesh, 48, 32, 2, 4, variant, True)
two = V.scenes.random_spheres(60); two.materials[8, 4:7] = 40.0
run(two, 40, 30, 2, 3, 1, True)                                                            # four-kernel pipeline
run(V.scenes.raytracer_default(), 48, 32, 1, 2, 0, False, integrator=V.INTEGRATOR_WHITTED)
# frame exchange on one device: rank 1 attaches to rank 0
g = V.Renderer(64, 64, spp=2, max_depth=3, variant=1, tile_shard=(0, 2)); p = V.Renderer(64, 64, spp=2, max_depth=3, variant=1, tile_shard=(1, 2))
g.exchange_create(); p.exchange_attach(g)
for x in (g, p):
    x.set_scene(V.scenes.random_spheres(100)); x.build_bvh()
for i in range(3):
    for x in (p, g):
        x.set_frame_index(i); x.draw(fd)
g.read_accum(); p.close(); g.close()
print("sanitizer cases done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > $O/${T}_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $O/${T}_$tool.log) $(grep 'ERROR SUMMARY' $O/${T}_$tool.log | tail -1)"
done
