"""Multi-GPU plumbing: one process (and one vkrt context) per GPU, torch.distributed for the one
exchange step of the path -- the gather of every rank's owned accumulator tiles to rank 0
(SURVEY.md 8e).  Tile shards are disjoint pixels, so the recombined image is bit-identical to a
single-GPU render; sample shards are summed on rank 0 in rank order (deterministic).

Rank layout for world = T * S: tile_rank = rank % T, sample_rank = rank // T.
"""
import torch
import torch.distributed as dist


def shard_layout(rank, world, sample_shards=1):
    assert world % sample_shards == 0
    tiles = world // sample_shards
    return (rank % tiles, tiles), (rank // tiles, sample_shards)


class FrameGather:
    """Owns the send/receive tensors; gather() enqueues pack -> dist.gather -> unpack on `stream`."""

    def __init__(self, renderer, rank, world, sample_shards, stream, device, group=None):
        self.r, self.rank, self.world, self.group = renderer, rank, world, group
        self.stream, self.device = stream, device
        (self.tile_rank, self.tiles), (self.sample_rank, self.samples) = shard_layout(rank, world, sample_shards)
        self.n_floats = renderer.shard_floats(0)          # the largest shard; every rank sends this size
        dev = device if device.type == "cuda" else "cpu"
        self.send = torch.zeros(self.n_floats, dtype=torch.float32, device=dev)
        self.recv = [torch.zeros(self.n_floats, dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None

    def gather(self):
        """After renderer.draw(): collects every shard into rank 0's full accumulator."""
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            self.r.pack_shard_into(self.send.data_ptr(), self.n_floats)
            if self.world > 1:
                dist.gather(self.send, self.recv, dst=0, group=self.group)
            if self.rank == 0:
                bufs = self.recv if self.world > 1 else [self.send]
                for src, buf in enumerate(bufs):
                    (tile_rank, tiles), (sample_rank, _) = shard_layout(src, self.world, self.samples)
                    self.r.unpack_shard(buf.data_ptr(), tile_rank, tiles, add=(sample_rank > 0))


class _Null:
    def __enter__(self): return self
    def __exit__(self, *a): return False
