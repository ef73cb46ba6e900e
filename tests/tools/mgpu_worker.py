"""Worker of tests/test_multi_gpu.py (one process per GPU, torchrun): renders shards of the same frames on every rank,
exchanges them to rank 0 over (a) peer stores / CUDA IPC and (b) an NCCL gather, and compares rank 0's result bit for
bit with a single-GPU render of the whole frame made by rank 0 itself."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vk_renderer_b200 as V  # noqa: E402
from helpers import bits_equal, mismatch_report  # noqa: E402
from vk_renderer_b200.sharding import FrameGather, PeerExchange, shard_layout  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
stream = torch.cuda.Stream(device=device)

w, h, spp, depth = 500, 300, 8, 6          # not a multiple of the 32-px tile
scene = V.scenes.random_spheres(3000)
FRAMES = 5


def fd(i):
    return V.default_frame_data(aspect_ratio=w / h, seed=(0.37 * (i + 1)) % 1.0)


def make(tile=(0, 1), samp=(0, 1), flags=0, variant=V.VARIANT_WAVEFRONT):
    r = V.Renderer(w, h, spp=spp, max_depth=depth, variant=variant, flags=flags, device_id=local, tile_shard=tile,
                   sample_shard=samp, stream=stream.cuda_stream)
    r.set_scene(scene); r.build_bvh(); r.set_seed(77)
    return r


def single(progressive=False):
    """Rank 0's own single-GPU render of the same frames: accumulators and images of every frame."""
    r = make(flags=V.FLAG_PROGRESSIVE if progressive else 0)
    out = []
    for i in range(FRAMES):
        r.set_frame_index(i); r.draw(fd(i))
        out.append((r.read_accum(), r.read_rgba8()))
    r.close()
    return out


ref = single() if rank == 0 else None
ref_prog = single(progressive=True) if rank == 0 else None

for variant in (V.VARIANT_WAVEFRONT, V.VARIANT_MEGAKERNEL):
    for sample_shards in ((1, 2) if world % 2 == 0 else (1,)):
        tile, samp = shard_layout(rank, world, sample_shards)
        results = {}
        for mode in ("peer", "nccl", "peer-progressive"):
            peer = mode.startswith("peer")
            prog = mode.endswith("progressive")
            flags = (V.FLAG_NO_RESOLVE if (rank != 0 or not peer) else 0) | (V.FLAG_PROGRESSIVE if (prog and rank == 0) else 0)
            r = make(tile, samp, flags, variant)
            xch = PeerExchange(r, rank, world) if peer else None
            gather = None if peer else FrameGather(r, rank, world, sample_shards, stream, device)
            got = []
            for i in range(FRAMES):                       # frames are enqueued back to back: two in flight, no host sync
                r.set_frame_index(i); r.draw(fd(i))
                if gather is not None:
                    gather.gather()
                    if rank == 0:
                        r.resolve()
                if rank == 0 and i in (0, FRAMES - 1):
                    got.append((r.read_accum(), r.read_rgba8()))
            r.wait_idle()
            if xch is not None:
                xch.close()
            r.close()
            results[mode] = got
            dist.barrier()
        if rank == 0:
            for k, i in enumerate((0, FRAMES - 1)):
                a_peer, img_peer = results["peer"][k]
                a_nccl, img_nccl = results["nccl"][k]
                assert bits_equal(a_peer, a_nccl), "peer vs NCCL gather, frame %d: %s" % (i, mismatch_report(a_peer, a_nccl))
                assert np.array_equal(img_peer, img_nccl)
                if sample_shards == 1:       # tile shards are disjoint pixels: bit-identical to one GPU
                    assert bits_equal(a_peer, ref[i][0]), "tile shards vs 1 GPU, frame %d: %s" % (i, mismatch_report(a_peer, ref[i][0]))
                    assert np.array_equal(img_peer, ref[i][1])
                    a_prog, img_prog = results["peer-progressive"][k]
                    assert bits_equal(a_prog, ref_prog[i][0]), "progressive, frame %d: %s" % (i, mismatch_report(a_prog, ref_prog[i][0]))
                    assert np.array_equal(img_prog, ref_prog[i][1])
                else:                        # sample groups are added in rank order: a different (fixed) summation order
                    g = a_peer[..., :3] / a_peer[..., 3:4]
                    o = ref[i][0][..., :3] / ref[i][0][..., 3:4]
                    assert np.allclose(g, o, rtol=1e-5, atol=1e-6) and np.array_equal(a_peer[..., 3], ref[i][0][..., 3])
            print("variant %d sample_shards %d ok" % (variant, sample_shards)); sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
print("rank%d_ok" % rank); sys.stdout.flush()
