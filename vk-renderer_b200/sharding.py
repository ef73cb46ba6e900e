"""Multi-GPU plumbing: one process (and one vkrt context) per GPU, torch.distributed for the one
exchange step of the path -- the gather of every rank's owned accumulator tiles to rank 0
(SURVEY.md 8e).  Tile shards are disjoint pixels, so the recombined image is bit-identical to a
single-GPU render; sample shards are summed on rank 0 in rank order (deterministic).

Rank layout for world = T * S: tile_rank = rank % T, sample_rank = rank // T.

Two exchange modes: `PeerExchange` (default in bench.py) -- every rank's last kernel stores its pixels straight into
rank 0's memory over NVLink (csrc/vkrt_exchange.cu); torch.distributed only carries the 96-byte CUDA-IPC handle once.
`FrameGather` -- pack -> NCCL gather -> unpack, the plain-collective form of the same step.
"""
import numpy as np
import torch
import torch.distributed as dist

TILE = 32
TILE_PX = TILE * TILE


def shard_layout(rank, world, sample_shards=1):
    assert world % sample_shards == 0
    tiles = world // sample_shards
    return (rank % tiles, tiles), (rank // tiles, sample_shards)


def owned_pixels(width, height, tile_rank, tile_count):
    """Pixel index (y * width + x, or -1 outside the image) of every packed slot of a tile shard: the
    host-side statement of the kernels' slot_to_pixel (csrc/vkrt_render.cu).  Slots go tile by tile
    (tiles tile_rank, tile_rank + tile_count, ...), 32 consecutive slots are an 8x4 pixel block."""
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    n_tiles = tiles_x * tiles_y
    owned = (n_tiles - tile_rank + tile_count - 1) // tile_count if tile_rank < n_tiles else 0
    w = np.arange(owned * TILE_PX, dtype=np.int64)
    lt, inn = w >> 10, w & 1023
    gt = tile_rank + lt * tile_count
    tx, ty = gt % tiles_x, gt // tiles_x
    blk, l = inn >> 5, inn & 31
    px = tx * TILE + (blk & 3) * 8 + (l & 7)
    py = ty * TILE + (blk >> 2) * 4 + (l >> 3)
    pix = py * width + px
    pix[(px >= width) | (py >= height)] = -1
    return pix


def gather_packed(send, rank, world, group=None):
    """dist.gather of equally sized packed shard buffers to rank 0; returns the list there, else None."""
    recv = [torch.zeros_like(send) for _ in range(world)] if rank == 0 else None
    if world > 1:
        dist.gather(send, recv, dst=0, group=group)
    elif rank == 0:
        recv[0].copy_(send)
    return recv


class FrameGather:
    """Owns the send/receive tensors; gather() enqueues pack -> dist.gather -> unpack on `stream`."""

    def __init__(self, renderer, rank, world, sample_shards, stream, device, group=None):
        self.r, self.rank, self.world, self.group = renderer, rank, world, group
        self.stream, self.device = stream, device
        (self.tile_rank, self.tiles), (self.sample_rank, self.samples) = shard_layout(rank, world, sample_shards)
        self.n_floats = renderer.shard_floats(0)          # the largest shard; every rank sends this size
        self.send = torch.zeros(self.n_floats, dtype=torch.float32, device=device)
        self.recv = [torch.zeros(self.n_floats, dtype=torch.float32, device=device) for _ in range(world)] if rank == 0 else None

    def gather(self):
        """After renderer.draw(): collects every shard into rank 0's full accumulator."""
        with torch.cuda.stream(self.stream):
            self.r.pack_shard_into(self.send.data_ptr(), self.n_floats)
            if self.world > 1:
                dist.gather(self.send, self.recv, dst=0, group=self.group)
            if self.rank == 0:
                bufs = self.recv if self.world > 1 else [self.send]
                for src, buf in enumerate(bufs):
                    (tile_rank, tiles), (sample_rank, _) = shard_layout(src, self.world, self.samples)
                    self.r.unpack_shard(buf.data_ptr(), tile_rank, tiles, add=(sample_rank > 0))


class PeerExchange:
    """Frame exchange over peer memory: rank 0 creates the exchange block, the other ranks map it (CUDA IPC).  After
    attach(), renderer.draw() on every rank IS the exchange: no further call is needed, rank 0's accumulator and
    rgba8 image hold the whole frame."""

    def __init__(self, renderer, rank, world, group=None):
        self.r, self.rank, self.world = renderer, rank, world
        box = [renderer.exchange_create() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                renderer.exchange_open(box[0])
            dist.barrier(group=group)

    def close(self):
        self.r.wait_idle()
        if self.world > 1:
            dist.barrier()
        self.r.exchange_close()
