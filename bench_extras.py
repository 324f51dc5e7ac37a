"""Extra workloads reported next to the main bench line (BASELINE.json configs[1], [3], [4]):
  C2  ONE 100-feature filter (the latency-bound configuration): frames/s, end to end, kernel shares
  C4  1-point RANSAC support sweep, 1e5 hypotheses x 5000 matches, sharded over the ranks inside the library (NCCL MAX all-reduce of
      the packed (support, hypothesis id) key + the winner's mask): hypothesis-matches/s + HBM roofline of the support kernel
  C5  batch of independent 100-feature filters (the main line when --gpus N > 1)
Each returns a dict; failures are reported, never raised (the main line must survive)."""
import os
import time

import numpy as np

from ransac_slam_b200 import synth
import bench as B


def _events(stream):
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def consistent_state(scene, device, t0=3, rho_err=0.02, seed=0, lowrank=8):
    """truth-consistent (x, P) for a large scene, P assembled on the GPU"""
    rng = np.random.default_rng(seed + 1)
    N = scene.N
    x = scene.x0.copy()
    r, q = synth.truth_pose(t0)
    x[0:3] = r + rng.normal(0, 0.001, 3)
    qq = q + np.array([0.0, *rng.normal(0, 0.0004, 3)])
    x[3:7] = qq / np.linalg.norm(qq)
    x[7:10] = rng.normal(0, 0.005, 3)
    x[10:13] = rng.normal(0, 0.001, 3)
    d = np.linalg.norm(scene.landmarks - scene.x0[:3], axis=1)
    idx = 13 + 6 * np.arange(N) + 5
    x[idx] = (1.0 / d) * (1.0 + rho_err * rng.standard_normal(N))
    import torch

    P = synth.assemble_P_torch(scene, device, rho_std_rel=2 * rho_err, x=x, lowrank=lowrank, seed=seed)
    P[0:3, 0:3] += 0.002**2 * torch.eye(3, dtype=P.dtype, device=P.device)
    P[3:7, 3:7] += 0.0008**2 * torch.eye(4, dtype=P.dtype, device=P.device)
    return x, P


def bench_c2(args, world, rank, local):
    """C2: ONE 100-feature filter (BASELINE.json configs[1]) -- the latency-bound configuration (a frame touches ~3 MB and ~1e8 flop)."""
    import torch

    if rank != 0:
        return None
    K, W = 200, 10
    T = W + K
    scene, seq = B.make_c2(1234, T)
    N, n = scene.N, scene.x0.size
    rows, cols = seq.images.shape[1:]
    g = B.new_gpu_filter(scene, device=local)
    stream = torch.cuda.ExternalStream(g.stream)
    d_images = torch.from_numpy(seq.images).cuda()
    d_u01 = torch.from_numpy(seq.u01).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B.run_frames_resident(g, d_images, d_u01, range(W), False, flush, stream)
    l0 = g.launches
    ms = B.run_frames_resident(g, d_images, d_u01, range(W, T), True, flush, stream)
    launches = g.launches - l0
    pose_resident = g.download_pose()
    g.close()
    # per-kernel breakdown (instrumented pass)
    gp = B.new_gpu_filter(scene, device=local)
    sp = torch.cuda.ExternalStream(gp.stream)
    B.run_frames_resident(gp, d_images, d_u01, range(W), False, flush, sp)
    gp.profile(True)
    stats = []
    PF = 50
    for k in range(W, W + PF):
        with torch.cuda.stream(sp):
            flush.zero_()
        B.run_frames_resident(gp, d_images, d_u01, [k], False, flush, sp)
        stats.append(B.frame_stats(gp))
    prof = gp.profile_read()
    gp.profile(False)
    gp.close()
    st = {k: float(np.mean([s_[k] for s_ in stats])) for k in stats[0]}
    tot = sum(v[1] for v in prof.values())
    breakdown = {k: dict(launches_per_frame=v[0] / PF, us_per_frame=1e3 * v[1] / PF, share=v[1] / tot) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    # end to end
    ge = B.new_gpu_filter(scene, device=local)
    h_images = torch.from_numpy(seq.images).pin_memory()
    h_u01 = torch.from_numpy(seq.u01).pin_memory()
    img_b, u_b = rows * cols, seq.u01.shape[1] * 8

    T_e = h_images.shape[0]

    def e2e_in(k):
        k %= T_e
        return (h_images.data_ptr() + k * img_b, rows, cols, cols), (h_u01.data_ptr() + k * u_b, seq.u01.shape[1])

    def e2e_step(k):
        ge.frame(*e2e_in(k), predict=True)
        ge.prefetch(*e2e_in(k + 1))  # next frame's host->device copy beside this frame's kernels (one copy per frame)
        return ge.download_pose()

    for k in range(W):
        e2e_step(k)
    t0 = time.perf_counter()
    for k in range(W, T):
        pose = e2e_step(k)
    ge.sync()
    t_e2e = time.perf_counter() - t0
    assert np.allclose(pose, pose_resident, rtol=1e-9, atol=1e-12)
    ge.close()
    # the reference's own sources on the first frames of the same trajectory (whole frames: they take about a second each)
    cpu = None
    if not args.no_cpu:
        try:
            rc = B.RefC2(scene, seq)
            dts = [rc.step() for _ in range(5)][1:]
            cpu = dict(value=len(dts) / float(np.sum(dts)), unit="frames/s", cores=1, kind="reference", sample="frames 1-4 of the same trajectory, whole frames")
        except Exception as e:
            cpu = dict(error=repr(e))
    return dict(workload="C2: ONE synthetic 100-feature filter, 320x240, bounded trajectory (latency-bound by construction); L2 flushed between frames",
                value=K / (ms * 1e-3), unit="frames/s", ms_per_frame=ms / K, steps=K, warmup=W, gpu_launches_per_frame=launches / K,
                e2e=dict(value=K / t_e2e, unit="frames/s", h2d_bytes_per_step=img_b + u_b, d2h_bytes_per_step=104), frame_stats=st, kernels=breakdown,
                cpu_baseline=cpu)


def make_c4(device, N=5000, seed=77):
    cam = synth.scaled_camera(6)
    scene = synth.make_scene(N=N, seed=seed, cam=cam, margin=30, min_sep=18, assemble_P=False)
    x, P = consistent_state(scene, device, seed=seed)
    rng = np.random.default_rng(seed + 9)
    r, q = synth.truth_pose(3)
    uv, _ = synth.project(cam, r, q, scene.landmarks)
    uv = uv + 0.5 * rng.standard_normal(uv.shape)
    gross = rng.random(N) < 0.05
    uv[gross] += rng.uniform(20, 40, (int(gross.sum()), 2)) * rng.choice([-1, 1], (int(gross.sum()), 2))
    z = np.rint(uv)
    return scene, x, P, z


def bench_c4(args, world, rank, local):
    """C4: the support sweep sharded over the ranks INSIDE the library (rslam_support_sweep_multi: NCCL MAX all-reduce of the packed key on
    the handle's own stream, the winner's mask handed to every rank)."""
    import torch

    from ransac_slam_b200 import capi, sweep

    N, H = 5000, 100000
    dev = torch.device("cuda", local)
    scene, x, P, z = make_c4(dev, N)
    n = x.size
    hyp = np.random.Generator(np.random.MT19937(99)).integers(0, N, H).astype(np.int32)
    comm = sweep.comm_from_torch_distributed(local) if world > 1 else capi.Comm.single_process([local])
    res = {}
    for dedupe in (True, False):
        g = capi.Filter(scene.cam.as9(), N, batch=1, device=local, dedupe=dedupe, quirks=0x6)
        xd = torch.from_numpy(x).to(dev)
        g.upload_state_device(xd.data_ptr(), P.data_ptr(), n, n, N, prior=True)
        g.search_ic_matches()  # h, H, S at x_k_km1 (no image: matches are injected)
        g.set_matches(z, np.ones(N, dtype=np.uint8))
        d_hyp = torch.from_numpy(hyp).to(dev)
        d_key = torch.zeros(1, dtype=torch.int64, device=dev)
        # sharding axis: dedupe -> by match index (each distinct hypothesis scored on exactly one GPU); brute force -> by hypothesis id
        shard = capi.SHARD_BY_MATCH if dedupe else capi.SHARD_BY_HYPOTHESIS
        stream = torch.cuda.ExternalStream(g.stream)
        reps = 10 if dedupe else 2

        def sweep_once():
            comm.support_sweep([g], d_hyp.data_ptr(), shard=shard, key_device_ptr=d_key.data_ptr(), n_hyp=H)

        for _ in range(3):
            sweep_once()
        g.sync()
        B.barrier(world)
        e0, e1 = _events(stream)
        e0.record(stream)
        for _ in range(reps):
            sweep_once()
        e1.record(stream)
        g.sync()
        ms = B.max_over_ranks(e0.elapsed_time(e1) / reps, world)
        # the same sweep returning key AND the winner's mask to the host of every rank (wall clock around the blocking call)
        B.barrier(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            comm.support_sweep([g], d_hyp.data_ptr(), shard=shard, want_mask=True, n_hyp=H)
        ms_mask = B.max_over_ranks((time.perf_counter() - t0) / reps * 1e3, world)
        key, mask, pairs = comm.support_sweep([g], hyp, shard=shard, want_mask=True, want_pairs=True)
        assert key == (int(d_key.item()) & 0xFFFFFFFFFFFFFFFF)
        support, hid = capi.decode_key(key)
        assert int(mask.sum()) == support, (int(mask.sum()), support)
        # kernel-only time of the support kernel (instrumented pass)
        g.profile(True)
        comm.support_sweep([g], hyp, shard=shard, want_mask=False)
        prof = g.profile_read()
        g.profile(False)
        k_ms = B.max_over_ranks(prof.get("k_ransac_support", (1, 0.0))[1], world)
        setup_ms = B.max_over_ranks(sum(v[1] for kname, v in prof.items() if kname != "k_ransac_support"), world)
        pairs_all = B.sum_over_ranks(float(pairs), world)
        pk = B.peaks()
        name = "dedupe" if dedupe else "brute_force"
        res[name] = dict(ms_per_sweep=ms, value=H * float(N) / (ms * 1e-3), unit="hypothesis-matches/s", winner=dict(support=support, hypothesis=hid),
                         mask_bits=int(mask.sum()), pairs_scored_on_device=pairs_all, ms_per_sweep_key_and_mask_on_host=ms_mask,
                         roofline=dict(kernel="k_ransac_support", bound="hbm", unit="GB/s", achieved=288.0 * pairs_all / world / (ms * 1e-3) / 1e9,
                                       peak=pk["hbm_gbs"], frac=288.0 * pairs_all / world / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                       kernel_only=dict(ms=k_ms, achieved=288.0 * pairs_all / world / (k_ms * 1e-3) / 1e9 if k_ms else None,
                                                        frac=288.0 * pairs_all / world / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"] if k_ms else None,
                                                        other_kernels_ms=setup_ms,
                                                        note="per-launch CUDA events of an instrumented sweep (not the timed region)"),
                                       note="288 B per scored (hypothesis, match) pair (SURVEY 8d), per GPU; whole sweep timed (compact + hypotheses + marks + "
                                            "support + reduce + the NCCL all-reduce of the key); peak " + pk["source"]))
        g.close()
    comm.close()
    return dict(workload="C4: support sweep 1e5 hypotheses x 5000 matches (n=30013, P 7.2 GB replicated per GPU) on a consistent map (quirk Q1 off: real "
                         "inlier sets), sharded over the ranks inside the library (rslam_support_sweep_multi, NCCL on the handle's stream)",
                n_gpus=world, scaling="strong", **res)


def bench_c5(args, world, rank, local):
    """C5 as an extra of the --gpus 1 line (at --gpus N > 1 it IS the main line: bench.c5_line)"""
    if world > 1:
        return dict(skipped="main line of this run")
    r = B.bench_c5(args, world, rank, local, steps=6, warmup=3, with_e2e=True)
    K = r["steps"]
    return dict(workload=f"C5: {B.C5_FILTERS} independent 100-feature filters on one GPU", value=B.C5_FILTERS * K / (r["ms"] * 1e-3), unit="frames/s",
                ms_per_batch_frame=r["ms"] / K, steps=K, warmup=r["warmup"], e2e=dict(value=B.C5_FILTERS * r["e2e_steps"] / r["e2e_seconds"], unit="frames/s",
                                                                                   h2d_bytes_per_step=r["e2e_bytes"][0], d2h_bytes_per_step=r["e2e_bytes"][1]),
                frame_stats=r["frame_stats"], kernels=r["kernels"], roofline=r["roofline"], whole_frame_hbm=r["whole_frame_hbm"])


def run(args, world, rank, local):
    out = {}
    which = [w.strip() for w in args.extras.split(",") if w.strip()]
    for name, fn in (("c2", bench_c2), ("c4", bench_c4), ("c5", bench_c5)):
        if name not in which:
            continue
        t0 = time.time()
        try:
            out[name] = fn(args, world, rank, local) or {}
        except Exception as e:
            import traceback

            out[name] = dict(error=repr(e), trace=traceback.format_exc()[-800:])
        out[name]["wall_s"] = time.time() - t0
        import torch

        torch.cuda.empty_cache()
    return out
