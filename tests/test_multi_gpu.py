"""N > 1 on real GPUs: these tests need at least two CUDA devices and skip otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import bits_equal, mismatch_report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_ranks_over_nccl_and_peer_stores_match_one_gpu():
    """One process per GPU (torchrun, NCCL): tile shards exchanged by peer stores over NVLink (CUDA IPC) and by an NCCL
    gather are bit-identical to each other and to the single-GPU frame, for both kernel variants, with frames in flight,
    with progressive accumulation; tile x sample shards agree between the two exchanges."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tests", "tools", "mgpu_worker.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    assert "rank0_ok" in r.stdout and "rank1_ok" in r.stdout


@pytest.mark.gpu
def test_one_context_on_several_devices_matches_one_gpu(vk):
    """In-library multi-GPU (vkrt_create_info.device_ids / n_devices): ONE context, one vkrt_draw per frame like
    GraphicsDevice::Draw, the devices' tile shards land in device_ids[0]'s accumulator through peer memory."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    V = vk
    w, h = 500, 300
    scene = V.scenes.random_spheres(3000)
    out = {}
    for key, kw in (("one", dict(device_id=0)), ("two", dict(device_ids=list(range(min(n_gpus(), 4)))))):
        for variant in (V.VARIANT_WAVEFRONT, V.VARIANT_MEGAKERNEL):
            r = V.Renderer(w, h, spp=8, max_depth=6, variant=variant, flags=V.FLAG_HIT_IDS, **kw)
            r.set_scene(scene); r.build_bvh(); r.set_seed(5)
            frames = []
            for i in range(4):
                r.set_frame_index(i)
                r.draw(V.default_frame_data(aspect_ratio=w / h, seed=0.1 * (i + 1)))
            frames.append((r.read_accum(), r.read_rgba8()))
            c = r.counters()
            out[(key, variant)] = (frames, c.closest_rays, c.shadow_rays, c.paths)
            r.close()
    for variant in (V.VARIANT_WAVEFRONT, V.VARIANT_MEGAKERNEL):
        a, b = out[("one", variant)], out[("two", variant)]
        assert bits_equal(a[0][0][0], b[0][0][0]), mismatch_report(a[0][0][0], b[0][0][0])
        assert np.array_equal(a[0][0][1], b[0][0][1])
        assert a[1:] == b[1:]          # every ray of the frame was traced exactly once, somewhere


@pytest.mark.gpu
def test_exchange_argument_checks(vk):
    V = vk
    r = V.Renderer(64, 64, spp=1, max_depth=2, tile_shard=(1, 2))
    with pytest.raises(V.VkrtError):
        r.exchange_create()                       # only the gathering rank (0, 0) owns the block
    with pytest.raises(V.VkrtError):
        r.exchange_open(b"\0" * 96)               # not a handle
    g = V.Renderer(64, 64, spp=1, max_depth=2, tile_shard=(0, 2))
    other = V.Renderer(32, 64, spp=1, max_depth=2, tile_shard=(1, 2))
    g.exchange_create()
    with pytest.raises(V.VkrtError):
        g.exchange_create()                       # already attached
    with pytest.raises(V.VkrtError):
        other.exchange_attach(g)                  # another frame size
    # same device, same process: rank 1 attaches and both draw -- the exchange works on one GPU too
    r.exchange_attach(g)
    for x in (g, r):
        x.use_default_scene(V.SCENE_TRACER); x.set_seed(3)
    full = V.Renderer(64, 64, spp=1, max_depth=2)
    full.use_default_scene(V.SCENE_TRACER); full.set_seed(3)
    fd = V.default_frame_data(aspect_ratio=1.0, seed=0.25)
    for i in range(3):
        for x in (r, g, full):                    # the gathering rank's draw waits (on the device) for rank 1's
            x.set_frame_index(i); x.draw(fd)
    assert bits_equal(g.read_accum(), full.read_accum())
    assert np.array_equal(g.read_rgba8(), full.read_rgba8())
    for x in (r, g, other, full):
        x.close()


@pytest.mark.gpu
def test_cpp_graphics_device_on_two_gpus_draws_the_same_image(tmp_path):
    """The C++ GraphicsDevice drop-in (csrc/host): `--devices 0,1` changes nothing but vkrt_create_info.device_ids -- one
    Draw per frame, the image is byte-identical to the single-GPU run (same camera script, same srand seed)."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_vkrt_build", os.path.join(ROOT, "vk-renderer_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    exe = b.build_host()
    imgs = []
    for extra in ([], ["--devices", "0,1"]):
        for wf in ([], ["--wavefront"]):
            out = tmp_path / ("frame%d.ppm" % len(imgs))
            r = subprocess.run([exe, "--frames", "6", "--res", "320", "--spp", "4", "--depth", "4", "--seed", "7", "--out", str(out)] + extra + wf,
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            imgs.append(open(out, "rb").read())
    assert imgs[0] == imgs[1] == imgs[2] == imgs[3]
