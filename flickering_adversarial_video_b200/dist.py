"""Multi-GPU plumbing for the sharded (universal / class-generalisation) attacks: one process per
GPU, clips sharded on dim 0, delta replicated, ONE sum-all-reduce of the packed [T*3 | scalars] buffer
per step.  The reference never had a working collective on this path (the TF
MirroredStrategy is disabled at i3d_adversarial_main_universal.py:309-312, torch uses single-process
nn.DataParallel, model.py:576-578); the semantics preserved are "sum of per-sample gradients" for the
margin loss and "mean over the global batch" for the CE loss."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_range(global_batch, rank, world):
    """Contiguous equal shards; the reference drops remainders (`drop_remainder=True`,
    i3d_adversarial_main_universal.py:241), so the global batch must divide evenly."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def allreduce_sum_(t, group=None):
    """In-place sum over ranks of the packed gradient/scalar buffer (no-op for one rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def ce_grad_divisor(local_batch, world):
    """The CE loss is a mean over the GLOBAL batch (utils/kinetics_i3d_utils.py:305): every rank
    divides its per-sample terms by local_batch*world so that the sum over ranks is the global mean."""
    return local_batch * world


def sum_counts(values, device="cpu", group=None):
    """Sum a few host counters (validation miss / total counts of the fooling ratio) over ranks; returns floats.
    One rank: the values unchanged.  Every rank must call it the same number of times."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()


def agree_min(value, device="cpu", group=None):
    """MIN over ranks of one host integer (e.g. the number of batches every rank can supply this epoch)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def lockstep(batches, world=1, device="cpu", group=None, num_batches=None):
    """Iterate a per-rank batch source so that EVERY rank runs the same number of steps.  Sharded attacks exchange
    their gradient with one collective per step: a rank whose shard holds fewer batches (record files of unequal
    length — the conversion script skips videos shorter than 90 frames) would leave the loop early and issue a
    different collective (the validation count sum) while the others still wait in the step's all-reduce.
    With `num_batches` (the local batch count, when the source knows it) the ranks agree once on the minimum and every
    rank stops there; otherwise one MIN all-reduce of a has-next flag precedes every batch.  One rank: plain iteration."""
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        yield from batches
        return
    it = iter(batches)
    if num_batches is not None:
        n = agree_min(num_batches, device, group)
        for _ in range(n):
            yield next(it)
        return
    while True:
        item = next(it, None)
        if not agree_min(0 if item is None else 1, device, group):
            return
        yield item


def shard_indices(indices, rank, world):
    """Equal-count shard of an index list for one rank: every world-th entry starting at `rank`, cut to the common
    length len(indices) // world (the remainder is dropped, like the reference's drop_remainder batching)."""
    per = len(indices) // world
    return list(indices[rank::world])[:per]


def broadcast_ints(values, src=0, device="cpu", group=None):
    """Rank `src`'s list of integers on every rank (e.g. the random train / test split permutation)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return [int(v) for v in values]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.broadcast(t, src=src, group=group)
    return t.tolist()


def replicas_equal(t, group=None):
    """True when `t` (the replicated perturbation / Adam state) is bit-identical on every rank: one MAX all-reduce of
    [t, -t] gives the element-wise max and min over ranks, which coincide only if all ranks agree.  The sharded attacks
    rely on this (every rank applies the same update to its own copy after the gradient all-reduce); drivers call it
    periodically.  One rank: trivially True."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return True
    both = torch.cat([t.reshape(-1), -t.reshape(-1)]).contiguous()
    dist.all_reduce(both, op=dist.ReduceOp.MAX, group=group)
    n = t.numel()
    return bool(torch.equal(both[:n], -both[n:]))


def is_writer(world=None):
    """True on the rank that writes result files / summaries of a sharded run (rank 0); always True for one rank or for
    per-rank replicas (`world` = the attack object's world size: 1 when it does not take part in collectives)."""
    if world is not None and world <= 1:
        return True
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
