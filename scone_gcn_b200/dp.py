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


def is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(device_index=None):
    """Joins the process group `torchrun` set up (RANK / WORLD_SIZE / MASTER_* in the environment): NCCL with one GPU per
    rank when CUDA is there, gloo otherwise.  Returns (rank, world); (0, 1) and no group when WORLD_SIZE <= 1."""
    rank, world = env_rank_world()
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if torch.cuda.is_available():
            local = int(os.environ.get('LOCAL_RANK', rank)) if device_index is None else device_index
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group('gloo')
    return dist.get_rank(), dist.get_world_size()


def shard_rows(rows):
    """This rank's contiguous share of the row list every rank computed identically (same RNG stream on every rank)."""
    import torch.distributed as dist
    if not is_distributed():
        return rows
    lo, hi = shard_range(len(rows), dist.get_rank(), dist.get_world_size())
    return rows[lo:hi]
