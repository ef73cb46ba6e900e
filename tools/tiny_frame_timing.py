import sys, time
sys.path.insert(0, '/root/repo')
import vk_renderer_b200 as V
scene = V.scenes.grid_spheres()
for (w, h) in ((64, 64), (256, 256), (680, 384)):
    for variant in (1, 0):
        r = V.Renderer(w, h, spp=16, max_depth=8, variant=variant, flags=V.FLAG_LAUNCH_TIMING)
        r.set_scene(scene); r.build_bvh(); r.set_seed(1)
        fd = V.default_frame_data(aspect_ratio=w / h)
        for i in range(3): r.draw(fd)
        r.wait_idle()
        ts = []
        for i in range(5):
            r.draw(fd); r.wait_idle()
            ts.append((r.last_frame_timing(), r.last_frame_traversal_timing()))
        print(w, h, 'variant', variant, 'frame ms', ['%.3f' % t[0][1] for t in ts], 'traversal ms/launches', ts[-1][1])
        r.close()
