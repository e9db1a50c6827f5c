"""Multi-GPU sharding of independent streams: no collective on the hot path.

Streams never interact (every LyraEncoder/LyraDecoder object of the reference owns all of its state,
lyra/lyra_encoder.h:112-120, lyra/lyra_decoder.h:130-160), so global stream g lives on rank
``g // streams_per_rank`` as local stream ``g % streams_per_rank`` and each rank runs its own Context.
The only exchange is optional and at the edge: gathering the packets (8-23 B per stream-frame) on one rank.
"""
import numpy as np


def shard_range(num_streams, world_size, rank):
    """Contiguous block partition: -> (first_global_stream, count) owned by `rank`."""
    per = -(-num_streams // world_size)
    first = min(rank * per, num_streams)
    return first, max(0, min(per, num_streams - first))


def owner(global_stream, num_streams, world_size):
    per = -(-num_streams // world_size)
    return global_stream // per, global_stream % per


def gather_packets(local_packets, num_streams, group=None, dst=0):
    """Edge gather of per-rank packet arrays [count, P] to rank `dst` with torch.distributed (NCCL on GPUs,
    gloo in the CPU tests).  Returns the [num_streams, P] array on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-num_streams // world)
    P = local_packets.shape[1]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.zeros((per, P), dtype=torch.uint8, device=dev)
    buf[:local_packets.shape[0]] = torch.as_tensor(np.ascontiguousarray(local_packets), device=dev)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat(out)[:num_streams].cpu().numpy()
