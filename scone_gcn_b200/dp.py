"""Data-parallel plumbing: trajectories are sharded across ranks, the complex is replicated, and the
only exchange per optimizer step is one sum over ranks of the flat [grads | nll_sum | count] buffer — an NCCL all-reduce, or
(PeerExchange, the default on one node) the library's own kernel that pushes the buffer through NVLink peer memory and applies
the Adam update in the same launch."""
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


class PeerExchange:
    """Gradient exchange fused with the Adam step over NVLink peer memory (csrc/scone_dp.cu, include/scone_b200.h): every rank
    allocates an exchange buffer, the CUDA IPC handles are all-gathered once through the process group, and every optimizer step
    is ONE kernel per rank (push to all peers, wait, sum in rank order, Adam).  One process per GPU of one node, at most 8 ranks."""

    def __init__(self, n_floats, device):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.handle = None
        nbytes = int(L.scone_dp_handle_bytes())
        mine = (ctypes.c_ubyte * nbytes)()
        err = None
        # every rank runs the same collectives whatever fails locally; the outcome is agreed with a MIN all-reduce
        try:
            h = ctypes.c_void_p()
            _lib.check(L.scone_dp_create(self.rank, self.world, int(n_floats), ctypes.byref(h)), 'scone_dp_create')
            self.handle = h
            _lib.check(L.scone_dp_get_handle(self.handle, ctypes.cast(mine, ctypes.c_void_p)), 'scone_dp_get_handle')
        except Exception as e:                             # noqa: BLE001
            err = e
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=device)
        allh = torch.empty(self.world * nbytes, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, t)
        flag = torch.tensor([0 if err else 1], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)        # every rank created its buffer
        if int(flag.item()) and err is None:
            try:
                buf = (ctypes.c_ubyte * (self.world * nbytes)).from_buffer_copy(allh.cpu().numpy().tobytes())
                _lib.check(L.scone_dp_open(self.handle, ctypes.cast(buf, ctypes.c_void_p)), 'scone_dp_open')
            except Exception as e:                         # noqa: BLE001
                err = e
        flag = torch.tensor([0 if err else 1], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)        # every rank has mapped every buffer (also the barrier before the first push)
        if not int(flag.item()):
            self.close()
            raise RuntimeError('peer exchange set-up failed on %s rank: %s' % ('this' if err else 'another', err))

    def adam_step(self, net, step, lr, weight_decay, stream=None):
        """Sum of the ranks' gradient buffers (left in net's buffer, as after an all-reduce) + Adam update, one launch."""
        from . import _lib
        _lib.check(_lib.lib().scone_model_dp_adam_step(net.handle, self.handle, int(step), float(lr), float(weight_decay), stream),
                   'scone_model_dp_adam_step')

    def status(self, stream=None):
        from . import _lib
        _lib.check(_lib.lib().scone_dp_status(self.handle, stream), 'scone_dp_status')

    def close(self):
        if getattr(self, 'handle', None):
            from . import _lib
            _lib.lib().scone_dp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_exchange(n_floats, device):
    """PeerExchange when the job is one NCCL process group of <= 8 CUDA ranks and SCONE_DP_EXCHANGE != 'nccl'; None otherwise
    (the caller all-reduces with the process group and calls the plain Adam step).  A failure to map the peers' buffers (e.g. GPUs
    hidden from each other by CUDA_VISIBLE_DEVICES) is reported once and also returns None — on EVERY rank: the decision is agreed."""
    import torch
    import torch.distributed as dist
    if not is_distributed() or os.environ.get('SCONE_DP_EXCHANGE', 'peer') == 'nccl':
        return None
    if dist.get_backend() != 'nccl' or dist.get_world_size() > 8 or not torch.cuda.is_available():
        return None
    try:
        return PeerExchange(n_floats, device)              # (raises on every rank or on none)
    except Exception as e:                                 # noqa: BLE001 - any failure means "use NCCL"
        if dist.get_rank() == 0:
            print('scone_gcn_b200.dp: peer exchange unavailable (%s); using NCCL all-reduce' % e, file=__import__('sys').stderr)
        return None
