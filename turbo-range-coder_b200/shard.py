"""Multi-GPU sharding of the batch path: independent blocks per rank, one gather of the compressed chunks.

The path has no data dependency between chunks, so ranks never exchange anything while coding.  The only
exchange is the collection of the packed streams on one rank: an all-gather of the
per-rank compressed lengths (8 bytes each) fixes every rank's offset, then each rank sends its payload to
the destination rank, which receives it at that offset (NCCL has no gatherv; grouped send/recv is its
idiom).  Works with any torch.distributed backend: NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def partition(n_blocks: int, world: int):
    """Contiguous block ranges per rank: block b belongs to rank b * world // n_blocks."""
    bounds = [-(-r * n_blocks // world) for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_bytes(total_len: int, block_len: int, world: int):
    """Byte ranges [start, end) of every rank's slice when the buffer is cut into block_len blocks."""
    n_blocks = -(-total_len // block_len)
    return [(min(a * block_len, total_len), min(b * block_len, total_len)) for a, b in partition(n_blocks, world)]


def gather_compressed(payload: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor = None):
    """Collect every rank's packed stream (1-D uint8 tensor, any length) on rank `dst`.

    Returns (buffer, offsets) on `dst` -- offsets is an int64 tensor of world+1 prefix sums, rank r's
    stream is buffer[offsets[r]:offsets[r+1]] -- and (None, offsets) elsewhere.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
    lens = torch.empty(world, dtype=torch.int64, device=payload.device)
    dist.all_gather_into_tensor(lens, mine, group=group)
    lens_h = lens.cpu()
    offs = torch.zeros(world + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens_h, 0)
    if world == 1:
        return payload, offs
    if rank == dst:
        total = int(offs[-1])
        buf = out if out is not None and out.numel() >= total else torch.empty(total, dtype=torch.uint8, device=payload.device)
        buf[int(offs[dst]):int(offs[dst + 1])] = payload
        ops = [dist.P2POp(dist.irecv, buf[int(offs[r]):int(offs[r + 1])], r, group) for r in range(world)
               if r != dst and lens_h[r] > 0]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return buf[:total], offs
    if payload.numel() > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, dst, group)]):
            w.wait()
    return None, offs
