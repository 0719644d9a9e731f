"""Multi-GPU host logic: a batch of independent images is split contiguously across ranks; every rank runs
the whole pipeline on its shard (no collective on the data path); labels are gathered at the end.

The reference has no distributed code at all; this is the only place a collective
appears, and it is optional (`gather_labels`)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous split of n_items over `world` ranks: the first (n_items % world) ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_labels(local_labels, n_total, group=None):
    """All-gather of ragged per-rank label shards [n_local, H, W] (int16) into [n_total, H, W] on every rank.

    Shards are padded to the largest shard so a single all_gather (NCCL on GPUs, gloo on CPU) moves them;
    u16/int16 stays on the wire."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(e - s for s, e in sizes)
    H, W = local_labels.shape[1:]
    padded = torch.zeros((n_max, H, W), dtype=local_labels.dtype, device=local_labels.device)
    padded[: local_labels.shape[0]] = local_labels
    wire = padded.view(torch.uint8)  # raw bytes: gloo has no int16 all_gather, and NCCL does not care
    out_b = [torch.empty_like(wire) for _ in range(world)]
    dist.all_gather(out_b, wire, group=group)
    out = [o.view(local_labels.dtype) for o in out_b]
    assert sizes[rank][1] - sizes[rank][0] == local_labels.shape[0]
    return torch.cat([o[: e - s] for o, (s, e) in zip(out, sizes)], dim=0)
