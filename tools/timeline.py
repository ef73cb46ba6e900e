"""Renders a few frames of (a tile shard of) the cfg4 workload and prints the per-launch timeline of the last one."""
import sys
sys.path.insert(0, '/root/repo')
import vk_renderer_b200 as V
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = V.scenes.grid_spheres()
w, h = 1920, 1080
r = V.Renderer(w, h, spp=16, max_depth=8, variant=1, tile_shard=(0, shard), flags=V.FLAG_LAUNCH_TIMING)
r.set_scene(scene); r.build_bvh(); r.set_seed(1)
fd = V.default_frame_data(aspect_ratio=w / h)
for i in range(6):
    r.draw(fd)
r.wait_idle()
r.dump_timeline('gpurun_out/timeline_%d.txt' % shard)
print(open('gpurun_out/timeline_%d.txt' % shard).read())
