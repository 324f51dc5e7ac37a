"""Extra workloads reported next to the main bench line (BASELINE.json configs[2..4]):
  C3  N = 2000 features (state dim 12013, P = 1.15 GB): frames/s + the fp64 DMMA covariance downdate against the fp64 peak
  C4  1-point RANSAC support sweep, 1e5 hypotheses x 5000 matches, sharded over the ranks with one MAX all-reduce of the
      packed (support, hypothesis id) key: hypothesis-matches/s + HBM roofline of the support kernel
  C5  batch of independent 100-feature filters split over the ranks (no inter-GPU traffic): frames/s
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


def bench_c3(args, world, rank, local):
    import torch

    from ransac_slam_b200 import capi

    if world > 1:
        return dict(skipped="a single filter's update stays on one GPU (replicas only); measured at N=1")
    N = 2000
    cam = synth.scaled_camera(4)
    scene = synth.make_scene(N=N, seed=1234, cam=cam, margin=30, min_sep=18, assemble_P=False, motion_scale=0.25)
    W, K = 2, 4
    n_u01 = 8192
    seq = synth.make_sequence(scene, T=W + K, seed=1235, n_u01=n_u01)
    n = scene.x0.size
    dev = torch.device("cuda", local)
    P0 = synth.assemble_P_torch(scene, dev)
    g = capi.Filter(cam.as9(), N, batch=1, device=local, std_a=0.007 * 0.25, std_alpha=0.007 * 0.25)
    x0 = torch.from_numpy(scene.x0).to(dev)
    g.upload_state_device(x0.data_ptr(), P0.data_ptr(), n, n, N)
    g.upload_patches(scene.templates.astype(np.float64))
    del P0
    stream = torch.cuda.ExternalStream(g.stream)
    d_images = torch.from_numpy(seq.images).to(dev)
    d_u01 = torch.from_numpy(seq.u01).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B.run_frames_resident(g, d_images, d_u01, range(W), False, flush, stream)
    ms = B.run_frames_resident(g, d_images, d_u01, range(W, W + K), True, flush, stream)
    # instrumented pass for the kernel breakdown (continues the same trajectory: frames W+K.. are not available, so re-run the last)
    g.profile(True)
    B.run_frames_resident(g, d_images, d_u01, [W + K - 1], False, flush, stream)
    st = B.frame_stats(g)
    prof = g.profile_read()
    g.profile(False)
    tot = sum(v[1] for v in prof.values())
    breakdown = {k: dict(launches=v[0], ms=v[1], share=v[1] / tot) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    fp64_peak = B.fp64_gemm_peak()
    kk = 2.0 * (st["m_li"] + st["m_hi"])
    gemm_ms = sum(v[1] for kname, v in prof.items() if kname.startswith("k_gemm_dmma"))
    syrk_ms = prof.get("k_gemm_dmma/syrk_P", (0, 0.0))[1]
    # fp64 flops issued on DMMA tiles in that frame: SYRK n^2 k (lower triangle) + TRSM trailing n k^2 + Cholesky trailing k^3/3
    k_li, k_hi = 2.0 * st["m_li"], 2.0 * st["m_hi"]
    flops = sum(float(n) * n * k + float(n) * k * k + k**3 / 3.0 for k in (k_li, k_hi))
    out = dict(workload="C3: synthetic 2000-feature map, 1280x960 camera (4x pixel density, motion and noise scaled 1/4), state dim 12013, P 1.15 GB fp64",
               value=K / (ms * 1e-3), unit="frames/s", ms_per_frame=ms / K, steps=K, warmup=W, frame_stats=st, kernels=breakdown,
               roofline=dict(kernel="k_gemm_dmma", bound="tensor", unit="TFLOP/s", achieved=flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                             peak=fp64_peak, frac=(flops / (gemm_ms * 1e-3) / 1e12 / fp64_peak) if gemm_ms else None,
                             flops_per_frame=flops, gemm_ms_per_frame=gemm_ms, k_li=k_li, k_hi=k_hi,
                             syrk=dict(ms=syrk_ms, tflops=(sum(float(n) * n * k for k in (k_li, k_hi)) / (syrk_ms * 1e-3) / 1e12) if syrk_ms else None),
                             note="peak = cuBLAS fp64 GEMM (torch.matmul 6144^3) measured in the same process; flops = n^2 k + n k^2 + k^3/3 per update"),
               cpu_baseline=dict(value=None, note="dense reference path at N=2000 is ~7e12 flop/frame plus up to ~9000 dense RANSAC hypotheses (SURVEY Appendix B): hours per frame on one core, not run"))
    g.close()
    return out


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
    import torch

    from ransac_slam_b200 import capi

    N, H = 5000, 100000
    dev = torch.device("cuda", local)
    scene, x, P, z = make_c4(dev, N)
    n = x.size
    hyp = np.random.Generator(np.random.MT19937(99)).integers(0, N, H).astype(np.int32)
    res = {}
    for dedupe in (True, False):
        g = capi.Filter(scene.cam.as9(), N, batch=1, device=local, dedupe=dedupe)
        xd = torch.from_numpy(x).to(dev)
        g.upload_state_device(xd.data_ptr(), P.data_ptr(), n, n, N, prior=True)
        g.set_matches(z, np.ones(N, dtype=np.uint8))
        g.search_ic_matches()  # h, H, S at x_k_km1 (no image: matches were injected)
        g.set_matches(z, np.ones(N, dtype=np.uint8))
        d_hyp = torch.from_numpy(hyp).to(dev)
        d_key = torch.zeros(1, dtype=torch.int64, device=dev)
        # sharding axis: dedupe -> by match index (each distinct hypothesis scored on exactly one GPU); brute force -> by hypothesis id
        if dedupe:
            h0, h1, t0, t1 = 0, H, rank * N // world, (rank + 1) * N // world
        else:
            h0, h1, t0, t1 = rank * H // world, (rank + 1) * H // world, 0, N
        stream = torch.cuda.ExternalStream(g.stream)
        reps = 5 if dedupe else 2

        def sweep():
            g.support_sweep(d_hyp.data_ptr(), h0, h1, want_mask=False, key_device_ptr=d_key.data_ptr(), n_hyp=H, match_begin=t0, match_end=t1)
            if world > 1:
                import torch.distributed as dist

                torch.cuda.current_stream().wait_stream(stream)
                dist.all_reduce(d_key, op=dist.ReduceOp.MAX)

        sweep()
        B.barrier(world)
        e0, e1 = _events(stream)
        e0.record(stream)
        for _ in range(reps):
            sweep()
        if world > 1:
            stream.wait_stream(torch.cuda.current_stream())
        e1.record(stream)
        g.sync()
        torch.cuda.synchronize()
        ms = B.max_over_ranks(e0.elapsed_time(e1) / reps, world)
        key = int(d_key.item())
        support, hid = capi.decode_key(key)
        # distinct pairs actually scored (host read, outside the timed region) + kernel-only time of the support kernel (instrumented pass)
        g.profile(True)
        _, _, pairs = g.support_sweep(hyp, h0, h1, want_mask=False, match_begin=t0, match_end=t1)
        prof = g.profile_read()
        g.profile(False)
        k_ms = B.max_over_ranks(prof.get("k_ransac_support", (1, 0.0))[1], world)
        pairs_all = B.sum_over_ranks(float(pairs), world)
        pk = B.peaks()
        name = "dedupe" if dedupe else "brute_force"
        res[name] = dict(ms_per_sweep=ms, value=H * float(N) / (ms * 1e-3), unit="hypothesis-matches/s", winner=dict(support=support, hypothesis=hid),
                         pairs_scored_on_device=pairs_all,
                         roofline=dict(kernel="k_ransac_support", bound="hbm", unit="GB/s", achieved=288.0 * pairs_all / world / (ms * 1e-3) / 1e9,
                                       peak=pk["hbm_gbs"], frac=288.0 * pairs_all / world / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                       kernel_only=dict(ms=k_ms, achieved=288.0 * pairs_all / world / (k_ms * 1e-3) / 1e9 if k_ms else None,
                                                        frac=288.0 * pairs_all / world / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"] if k_ms else None),
                                       note="288 B per scored (hypothesis, match) pair (SURVEY 8d), per GPU; whole sweep timed (compact + hyp + mark + support + reduce); peak " + pk["source"]))
        g.close()
    return dict(workload="C4: support sweep 1e5 hypotheses x 5000 matches (n=30013, P 7.2 GB replicated per GPU), hypotheses sharded over ranks, MAX all-reduce of the packed key",
                n_gpus=world, scaling="strong", **res)


def bench_c5(args, world, rank, local):
    import torch

    from ransac_slam_b200 import capi

    Btot = int(os.environ.get("RSLAM_C5_FILTERS", "4096"))
    Bl = Btot // world
    NS = 8  # distinct scenes cycled over the batch
    W, K = 2, 4
    scenes, seqs = [], []
    for s in range(NS):
        sc, sq = B.make_c2(1234 + s + NS * rank, W + K)
        scenes.append(sc)
        seqs.append(sq)
    cam = scenes[0].cam
    dev = torch.device("cuda", local)
    g = capi.Filter(cam.as9(), 100, batch=Bl, device=local)
    n = scenes[0].x0.size
    for s in range(NS):
        xd = torch.from_numpy(scenes[s].x0).to(dev)
        Pd = torch.from_numpy(np.ascontiguousarray(scenes[s].P0)).to(dev)
        for b in range(s, Bl, NS):
            g.upload_state_device(xd.data_ptr(), Pd.data_ptr(), n, n, 100, b=b)
            g.upload_patches(scenes[s].templates.astype(np.float64), b=b)
    idx = torch.arange(Bl, device=dev) % NS
    imgs = torch.stack([torch.from_numpy(sq.images) for sq in seqs]).to(dev)  # NS x T x r x c
    u01s = torch.stack([torch.from_numpy(sq.u01) for sq in seqs]).to(dev)
    d_images = imgs[idx].transpose(0, 1).contiguous()  # T x Bl x r x c
    d_u01 = u01s[idx].transpose(0, 1).contiguous()  # T x Bl x n_u01
    stream = torch.cuda.ExternalStream(g.stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B.run_frames_resident(g, d_images, d_u01, range(W), False, flush, stream)
    B.barrier(world)
    ms = B.run_frames_resident(g, d_images, d_u01, range(W, W + K), False, flush, stream)
    B.barrier(world)
    ms = B.max_over_ranks(ms, world)
    # instrumented pass (events around every launch, no graph): where the batch frame goes
    g.profile(True)
    B.run_frames_resident(g, d_images, d_u01, [W + K - 1], False, flush, stream)
    prof = g.profile_read()
    g.profile(False)
    ft = [g.features(b=b) for b in range(0, Bl, max(1, Bl // 16))]
    m_hi = float(np.mean([f["hi"].sum() for f in ft]))
    m_li = float(np.mean([f["li"].sum() for f in ft]))
    tot = sum(v[1] for v in prof.values())
    breakdown = {k: dict(launches=v[0], ms=v[1], share=v[1] / tot) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    g.close()
    # HBM roofline of the batch frame: every filter's P (n x ldp fp64) is read and written once per non-empty update (the SYRK
    # downdate), and the 7 + 6 m columns of P that W = P H^T gathers are read once more; everything else is < 5 % of that
    ldp = (n + 15) // 16 * 16
    upd = (1 if m_li > 0 else 0) + (1 if m_hi > 0 else 0)
    bytes_per_filter = upd * 2.0 * n * ldp * 8 + 8.0 * n * (7 * upd + 6 * (m_li + m_hi))
    pk = B.peaks()
    ach = Bl * bytes_per_filter / (ms / K * 1e-3) / 1e9
    # the dominant kernel against ITS roofline: the covariance downdate P -= V V^T on the fp64 tensor pipe, (n + 1) n k flop per filter
    # (lower triangle x 2 flop), against the cuBLAS fp64 GEMM rate measured live
    top = next(iter(breakdown))
    kroof = None
    if top in ("k_syrk_rows", "k_gemm_dmma/syrk_P"):
        flops = Bl * float(n + 1) * n * 2.0 * (m_li + m_hi)
        tf = flops / (breakdown[top]["ms"] * 1e-3) / 1e12
        fp64_peak = B.fp64_gemm_peak()
        kroof = dict(kernel=top, bound="tensor", unit="TFLOP/s", achieved=tf, peak=fp64_peak, frac=tf / fp64_peak, share_of_step=breakdown[top]["share"],
                     note="(n + 1) n k flop per filter / kernel time; peak = cuBLAS fp64 GEMM measured live")
    return dict(workload=f"C5: {Btot} independent 100-feature filters, batch split over ranks, no inter-GPU traffic", n_gpus=world, scaling="strong",
                value=Btot * K / (ms * 1e-3), unit="filter-frames/s", ms_per_batch_frame=ms / K, filters_per_gpu=Bl,
                frame_stats=dict(m_li=m_li, m_hi=m_hi), kernels=breakdown,
                roofline=dict(bound="hbm", unit="GB/s", achieved=ach, peak=pk["hbm_gbs"], frac=ach / pk["hbm_gbs"], bytes_per_filter_frame=bytes_per_filter,
                              note="whole batch frame against HBM: P read+written once per non-empty update + the P columns gathered by W = P H^T; peak " + pk["source"]),
                roofline_dominant_kernel=kroof,
                note="working set (P 12.5 GB per 4096 filters) exceeds L2; no flush needed")


def run(args, world, rank, local):
    out = {}
    which = [w.strip() for w in args.extras.split(",") if w.strip()]
    for name, fn in (("c3", bench_c3), ("c4", bench_c4), ("c5", bench_c5)):
        if name not in which:
            continue
        t0 = time.time()
        try:
            out[name] = fn(args, world, rank, local)
        except Exception as e:
            import traceback

            out[name] = dict(error=repr(e), trace=traceback.format_exc()[-800:])
        out[name]["wall_s"] = time.time() - t0
        import torch

        torch.cuda.empty_cache()
    return out
