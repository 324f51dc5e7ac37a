"""Host-side helpers of the sharded 1-point RANSAC hypothesis sweep.

The sharded sweep itself lives in the library (include/rslam.h: rslam_comm_init / rslam_comm_init_rank / rslam_support_sweep_multi --
NCCL on the handles' own streams).  This module only (a) states the sharding and key conventions in Python for the CPU tests
(world-size-2 gloo) and (b) carries the 128-byte NCCL id from rank 0 to the other ranks of a torchrun job.

Hypotheses [0, H) are split in contiguous ranges over the ranks (or, with deduplication, the match indices are); every rank scores its
shard on its own GPU (state, matches and P replicated) and produces ONE packed 64-bit key; a MAX all-reduce of that key gives every
rank the same winner: highest support, ties -> lowest hypothesis id (the reference keeps the FIRST best hypothesis: strict '>' at
src/Tracking.cpp:507).  The winner's inlier mask follows from the rank that scored it.
"""
import numpy as np


def shard_range(n_hyp, world, rank):
    """contiguous hypothesis range of `rank`"""
    return rank * n_hyp // world, (rank + 1) * n_hyp // world


def pack_key(support, hyp_id):
    return (int(support) << 32) | (0xFFFFFFFF - int(hyp_id))


def decode_key(key):
    key = int(key)
    return key >> 32, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def local_key(supports, begin):
    """key of a local support array covering hypotheses [begin, begin + len): first maximum wins; 0 for an empty/zero range"""
    supports = np.asarray(supports)
    if supports.size == 0:
        return 0
    i = int(np.argmax(supports))  # first maximum
    return pack_key(int(supports[i]), begin + i)


def allreduce_key(key_tensor):
    """in-place MAX all-reduce of an int64 tensor holding the packed key (support < 2^31 keeps it non-negative)"""
    import torch.distributed as dist

    dist.all_reduce(key_tensor, op=dist.ReduceOp.MAX)
    return key_tensor


def shard_filters(n_filters, world, rank):
    """batched independent filters: contiguous block of filters per rank, no collective on the data path"""
    return rank * n_filters // world, (rank + 1) * n_filters // world


def comm_from_torch_distributed(device):
    """one process per GPU under torchrun: rank 0's NCCL id reaches the other ranks through torch.distributed (plumbing only);
    the communicator itself is the library's (rslam_comm_init_rank)."""
    import torch
    import torch.distributed as dist

    from ransac_slam_b200 import capi

    world, rank = dist.get_world_size(), dist.get_rank()
    uid = capi.Comm.unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(uid.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=0)
    return capi.Comm.from_rank(world, rank, t.cpu().numpy(), device)
