"""Multi-GPU sharding of the hot path: scan pairs are independent units (GroupNorm statistics are per pair,
geotransformer/modules/kpconv/modules.py:47), so the path shards exactly like the reference's own testers do - one
process per GPU, pairs dealt round-robin (DistributedSampler without shuffling, geotransformer/utils/torch.py:58-60;
experiments/test_batchoffline.py:255-266) - and needs NO data-path collective. The only communication is the final
reduction of per-rank counters / timings (the analogue of the reference's all_reduce of logged scalars,
geotransformer/utils/torch.py:16-21), done here in ONE collective instead of one per key."""
import torch
import torch.distributed as dist


def shard_pair_ids(num_pairs, rank, world_size, pad=False):
    """Pair ids of `rank`: rank, rank + world, ... With pad=True every rank gets ceil(num_pairs / world) ids by wrapping
    around (what torch's DistributedSampler(shuffle=False, drop_last=False) hands out), otherwise no pair is repeated."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    ids = list(range(rank, num_pairs, world_size))
    if pad and num_pairs > 0:
        per = -(-num_pairs // world_size)
        padded = (list(range(num_pairs)) * (1 + (per * world_size) // num_pairs))[:per * world_size]
        ids = padded[rank::world_size]
    return ids


def reduce_job_stats(pairs_done, elapsed_s, extra=None, group=None, device=None):
    """Whole-job statistics from per-rank counters in one all_gather: total pairs, max elapsed time over ranks (the
    job is as slow as its slowest rank) and pairs/s = total / max. `extra`: dict of per-rank floats, returned summed.
    Works without an initialised process group (single process)."""
    keys = sorted(extra) if extra else []
    mine = torch.tensor([float(pairs_done), float(elapsed_s)] + [float(extra[k]) for k in keys], dtype=torch.float64,
                        device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        rows = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(rows, mine, group=group)
        allv = torch.stack(rows).cpu()
    else:
        allv = mine[None].cpu()
    total, tmax = float(allv[:, 0].sum()), float(allv[:, 1].max())
    out = {"pairs": total, "elapsed_s": tmax, "pairs_per_s": total / tmax if tmax > 0 else 0.0,
           "per_rank_pairs": [float(x) for x in allv[:, 0]], "per_rank_elapsed_s": [float(x) for x in allv[:, 1]]}
    for i, k in enumerate(keys):
        out[k] = float(allv[:, 2 + i].sum())
    return out


def run_sharded(num_pairs, rank, world_size, register_fn, load_fn):
    """Processes this rank's share: results[pair_id] = register_fn(*load_fn(pair_id)). No collective inside."""
    return {pid: register_fn(*load_fn(pid)) for pid in shard_pair_ids(num_pairs, rank, world_size)}
