"""Row sharding of packed quantized layers (tensor parallel over OUTPUT rows).

Every packed layout of the reference (TCQ trellis, VQ/SQ tensor-core qweight, SIMT qweight) is strip-major: a block of
32*k consecutive output rows is ONE contiguous byte range of the buffer.  A rank's shard of a (possibly merged) projection
is therefore a zero-copy slice per merged member; the partial outputs are concatenated by an all-gather
(`gather_rows`).  Nothing here needs a GPU: the same functions run under gloo in the CPU tests.
"""
import torch


def shard_plan(group_sizes, rank, world):
    """[(first_row, n_rows)] of `rank` for a projection made of merged members with `group_sizes` rows each (e.g.
    [4096, 1024, 1024] for merge_qkv): the rank keeps rows [rank, rank+1) * size/world of EVERY member, so q/k/v heads
    (or up/gate channels) stay aligned on the same rank."""
    plan, base = [], 0
    for s in group_sizes:
        if s % (32 * world) != 0:
            raise ValueError(f"{s} rows cannot be split into {world} shards of whole 32-row strips")
        per = s // world
        plan.append((base + rank * per, per))
        base += s
    return plan


def shard_rows(packed, out_features, plan):
    """rows listed in `plan` of a packed buffer (any dtype/shape, `out_features` rows in total) as one flat tensor"""
    flat = packed.reshape(-1)
    if len(plan) == 1 and plan[0] == (0, out_features):
        return flat
    per_row = flat.numel() // out_features
    assert per_row * out_features == flat.numel()
    return torch.cat([flat[r0 * per_row:(r0 + n) * per_row] for r0, n in plan]).contiguous()


def gather_rows(local, world, group=None):
    """all-gather of per-rank row shards (1-D, equal sizes) into the full vector, rank-major"""
    import torch.distributed as dist
    out = torch.empty(local.numel() * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def unshard_index(group_sizes, world):
    """permutation that maps the rank-major concatenation of all ranks' shard outputs back to the original row order:
    full[idx[i]] = gathered[i]"""
    idx = []
    for rank in range(world):
        for r0, n in shard_plan(group_sizes, rank, world):
            idx.extend(range(r0, r0 + n))
    return torch.tensor(idx, dtype=torch.long)
