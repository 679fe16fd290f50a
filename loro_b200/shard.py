"""Multi-GPU plumbing of the path: documents are independent units, so a batch shards across ranks with no
data-path collective; the only exchange is one all-gather of the per-shard summary counters
(SURVEY.md 8e).  torch.distributed is plumbing only (NCCL on the GPU box, gloo in CPU tests)."""
import torch

COUNTER_KEYS = ("docs", "docs_ok", "blob_bytes", "blocks", "changes", "op_rows", "atom_ops", "pending_changes",
                "json_bytes", "state_hash")


def shard_range(total_docs, rank, world):
    """Contiguous balanced split: ranks [0, total % world) own one extra document."""
    base, extra = divmod(total_docs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_counters(counters, device=None):
    """All-gather the fixed-size counter struct of every shard; returns a list of dicts (one per rank)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    vals = [int(counters[k]) & 0x7FFFFFFFFFFFFFFF for k in COUNTER_KEYS]
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    if world == 1:
        return [dict(zip(COUNTER_KEYS, vals))]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [dict(zip(COUNTER_KEYS, o.tolist())) for o in out]
