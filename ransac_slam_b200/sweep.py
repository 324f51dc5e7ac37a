"""Host-side logic of the sharded 1-point RANSAC hypothesis sweep.

Hypotheses [0, H) are split in contiguous ranges over the ranks; every rank scores its range on its own GPU (state, matches and
P replicated) and produces ONE packed 64-bit key; a single MAX all-reduce of that key (NCCL on the GPUs, gloo in the CPU tests)
gives every rank the same winner: highest support, ties -> lowest hypothesis id (the reference keeps the FIRST best hypothesis:
strict '>' at src/Tracking.cpp:507).  Nothing else crosses the links.
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
