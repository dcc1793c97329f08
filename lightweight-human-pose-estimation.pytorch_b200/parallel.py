"""Data-parallel sharding of frames over the GPUs of one box: one process per GPU, no collective on the
compute path (frames are independent; the reference has no multi-GPU inference at all).  The only
exchange is an optional gather of the fixed-capacity pose tables over torch.distributed (NCCL on
NVLink/NVSwitch on the GPU box, gloo in the CPU tests)."""
import torch


def shard_range(n_items, rank, world_size):
    """Contiguous split [lo, hi) of n_items for `rank`; the first n_items % world_size ranks get one extra."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_pose_tables(n_poses, pose_entries, group=None, total=None):
    """all_gather of the per-rank results: n_poses int32 [b], pose_entries float64 [b, cap, 20] ->
    ([B], [B, cap, 20]) in frame order, on every rank.  With equal shards (total=None) B = world * b; with
    total = the global frame count, the ranks hold the (possibly unequal) shards of shard_range(total, rank, world):
    every rank pads its tables to the largest shard for the collective and the padding is dropped afterwards."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return n_poses, pose_entries
    if total is not None:
        rows = -(-total // world)
        if n_poses.shape[0] < rows:
            pad = rows - n_poses.shape[0]
            n_poses = torch.cat([n_poses, n_poses.new_zeros((pad,))], 0)
            pose_entries = torch.cat([pose_entries, pose_entries.new_zeros((pad,) + tuple(pose_entries.shape[1:]))], 0)
    ns = [torch.empty_like(n_poses) for _ in range(world)]
    ps = [torch.empty_like(pose_entries) for _ in range(world)]
    dist.all_gather(ns, n_poses.contiguous(), group=group)
    dist.all_gather(ps, pose_entries.contiguous(), group=group)
    if total is not None:
        sizes = [shard_range(total, r, world) for r in range(world)]
        ns = [t[:hi - lo] for t, (lo, hi) in zip(ns, sizes)]
        ps = [t[:hi - lo] for t, (lo, hi) in zip(ps, sizes)]
    return torch.cat(ns, 0), torch.cat(ps, 0)


def max_over_ranks(value, device=None, group=None):
    """max of a Python float over all ranks (timing: the slowest rank defines the step time)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
