"""Worker of tests/test_gpu_multi.py (one process per GPU under torchrun, or a single process driving all GPUs with --single):
the sweep sharded over the ranks by the library's own NCCL path must give, on every rank, the key AND the inlier mask of the
un-sharded sweep, on a map where the reference's (quirk Q1) support scoring finds real inliers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ransac_slam_b200 import capi, sweep  # noqa: E402
from tests import helpers as H  # noqa: E402


def build_filter(device, dedupe, N, seed):
    cam, x, P, z, ic = H.q1_consistent_state(N, seed=seed, n_loose=0)
    f = capi.Filter(cam.as9(), N, device=device, dedupe=dedupe)
    f.upload_state(x, P, prior=True)
    f.search_ic_matches()
    f.set_matches(z, ic.astype(np.uint8))
    return f, int(ic.sum())


def check(comm, filters, hyp, N, tag):
    ref_key, ref_mask, _ = filters[0].support_sweep(hyp)  # un-sharded, on this rank's replica
    s, hid = capi.decode_key(ref_key)
    assert s > N // 4, f"{tag}: the map must give a real winner, got support {s}"
    for shard in (capi.SHARD_BY_MATCH, capi.SHARD_BY_HYPOTHESIS):
        key, mask, pairs = comm.support_sweep(filters, hyp, shard=shard, want_pairs=True)
        assert key == ref_key, (tag, shard, capi.decode_key(key), (s, hid))
        assert (mask == ref_mask).all(), (tag, shard, int((mask != ref_mask).sum()))
        assert pairs > 0
    return s, hid


def main():
    N, NH = 1000, 20000
    single = "--single" in sys.argv
    hyp = np.random.Generator(np.random.MT19937(99)).integers(0, N, NH).astype(np.int32)
    if single:
        import torch

        G = torch.cuda.device_count()
        comm = capi.Comm.single_process(list(range(G)))
        out = []
        for dedupe in (True, False):
            fs = [build_filter(d, dedupe, N, seed=7)[0] for d in range(G)]
            hyp_c = np.minimum(hyp, build_filter(0, dedupe, N, seed=7)[1] - 1).astype(np.int32)
            out.append(check(comm, fs, hyp_c, N, f"single-process x{G} dedupe={dedupe}"))
            for f in fs:
                f.close()
        comm.close()
        print(f"ok single-process {G} GPUs winners {out}", flush=True)
        return
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    comm = sweep.comm_from_torch_distributed(local)
    assert comm.size == world and comm.local_size == 1
    out = []
    for dedupe in (True, False):
        f, nic = build_filter(local, dedupe, N, seed=7)
        hyp_c = np.minimum(hyp, nic - 1).astype(np.int32)
        out.append(check(comm, [f], hyp_c, N, f"rank {rank}/{world} dedupe={dedupe}"))
        # device-side key (enqueue only), as the bench uses it
        dk = torch.zeros(1, dtype=torch.int64, device=f"cuda:{local}")
        dh = torch.from_numpy(hyp_c).cuda(local)
        comm.support_sweep([f], dh.data_ptr(), shard=capi.SHARD_BY_MATCH, key_device_ptr=dk.data_ptr(), n_hyp=NH)
        f.sync()
        assert capi.decode_key(int(dk.item()) & 0xFFFFFFFFFFFFFFFF) == out[-1]
        f.close()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"ok rank {rank}/{world} winners {out}", flush=True)


if __name__ == "__main__":
    main()
