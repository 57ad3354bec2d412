"""Data-parallel plumbing: trajectories are sharded across ranks, the complex is replicated, and the
only exchange per optimizer step is one all-reduce of the flat [grads | nll_sum | count] buffer."""
import os


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n trajectories for `rank` (the first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(buf):
    """In-place sum over ranks of a torch tensor (CUDA -> NCCL, CPU -> gloo); no-op when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    return buf
