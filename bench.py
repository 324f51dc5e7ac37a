#!/usr/bin/env python
"""bench.py -- frames/s of the 1-point-RANSAC EKF measurement-update path (BASELINE.json metric) on B200.

Main line (contract): workload C2 = synthetic 100-feature inverse-depth map, 320x240 camera, bounded 1000-frame trajectory;
one step = one frame through the whole hot path (begin_frame + ekf_prediction + search_IC_matches + ransac_hypotheses +
ekf_update_li_inliers + rescue_hi_inliers + ekf_update_hi_inliers) via the C ABI (include/rslam.h).
  value : frames/s with all inputs (images, uniforms) already resident in HBM, CUDA events on the library's stream
  e2e   : frames/s through the same C-ABI call with HOST (pinned) images/uniforms copied in and the pose copied out every frame
  roofline, cpu_baseline : see DESIGN.md section "Measurement"
Extra workloads (same JSON line, key "workloads"): C3 N=2000 frames/s, C4 support-sweep hypothesis-matches/s, C5 batched filters.
`--impl reference` times the CPU oracle (the reference cannot be built here) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ransac_slam_b200 import synth  # noqa: E402

METRIC = "frames/s (EKF+1-pt RANSAC) at N=100 features"
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def dist_setup(n_gpus):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier(world):
    import torch

    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic work per launch (DESIGN.md "Measurement")
# ---------------------------------------------------------------------------------------------------------------------
def algorithmic_work(kernel, N, n, st):
    """returns (bound, amount, unit_label): bytes for HBM-bound kernels, flops for the fp64 tensor GEMM."""
    nic, mid, mli, mhi = st["nic"], st["mid"], st["m_li"], st["m_hi"]
    if kernel == "k_predict":
        return "hbm", 1310.0 * N + 1460.0
    if kernel == "k_search":
        return "hbm", N * ((2 * 3 + 13) ** 2 + 169 * 4 + 16.0)
    if kernel == "k_ransac_support":
        return "hbm", 288.0 * nic * mid
    if kernel == "k_upd_W":
        m = 0.5 * (mli + mhi)
        return "hbm", 8.0 * n * (7 + 6 * m) + 8.0 * n * 2 * m
    if kernel.startswith("k_gemm_dmma"):
        # fp64 flops issued on DMMA tiles per frame, by use: SYRK n^2 k (lower triangle), TRSM trailing ~ n k^2, Cholesky ~ k^3/3
        ks = [2.0 * mli, 2.0 * mhi]
        if kernel.endswith("syrk_P"):
            return "tensor", sum(float(n) * n * k for k in ks)
        if kernel.endswith("trsm_trail"):
            return "tensor", sum(float(n) * k * k for k in ks)
        if kernel.endswith("chol_outer") or kernel.endswith("chol_inner"):
            return "tensor", sum(k**3 / 3.0 for k in ks)
        return "tensor", sum(float(n) * n * k + float(n) * k * k + k**3 / 3.0 for k in ks)
    if kernel in ("k_chol_small", "k_chol_panel"):
        # Cholesky k^3/3 + the explicit inverses of the 64 x 64 diagonal blocks (64^3/3 each): fp64, one CTA, a serial column chain
        ks = [2.0 * mli, 2.0 * mhi]
        return "tensor", sum(k**3 / 3.0 + np.ceil(k / 64.0) * 64.0**3 / 3.0 for k in ks if k > 0)
    if kernel in ("k_trsm_small", "k_trsm_ll"):
        return "tensor", sum(float(n + 1) * k * k for k in (2.0 * mli, 2.0 * mhi))
    if kernel == "k_syrk_rows":
        return "tensor", sum(float(n + 1) * n * k for k in (2.0 * mli, 2.0 * mhi))
    if kernel == "k_ekf_prediction":
        return "hbm", 2 * 13 * n * 8.0 * 2
    if kernel == "k_upd_jnorm":
        return "hbm", 2 * 4 * n * 8.0 * 2
    return "hbm", None


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs.  "reference": the reference's OWN sources (oracle/_ref/libref.so = /root/reference/src/*.cpp compiled unmodified over the
# stand-in Eigen / OpenCV headers of oracle/ref_shim/), single threaded like the reference.  "port": the oracle restatement in dense
# (reference-faithful) mode.  Both run whole frames of the SAME C2 trajectory.  The synthetic images carry white-noise templates that
# the reference's bilinear patch warp decorrelates, so the reference's own matching finds fewer matches than the workload defines;
# to keep the work of the later stages identical to the GPU arm's, the match list of each frame is overwritten (untimed) with the
# workload's matches after the reference has done ALL of its own search work (prediction, Jacobians, S_i, patch warp, ZNCC).
# ---------------------------------------------------------------------------------------------------------------------
def cpu_frames(scene, seq, first, count, threads, prefer_reference=True):
    """time `count` frames starting at `first`; returns (frames/s, seconds, kind, cores, description)"""
    from oracle import oracle_py as O

    N = scene.N
    O.set_threads(threads)
    o = O.OracleFilter(scene.cam.as9(), std_z=scene.std_z, quirks=O.Q_ALL, sparse=False, fast_corr=False, warp_patches=False)
    for i in range(N):
        o.add_feature(0, None, scene.templates[i].astype(np.float64), scene.x0[:3], np.eye(3), scene.uv0[i])
    o.set_state(scene.x0, scene.P0)
    r = None
    if prefer_reference:
        try:
            from oracle import ref_py as R

            if R.available():
                r = R.ReferenceFilter()
                if not np.array_equal(r.camera9(), scene.cam.as9()):
                    r = None
        except Exception:
            r = None
    if r is None:
        for k in range(first):
            o.frame(seq.images[k], seq.u01[k])
        t0 = time.perf_counter()
        for k in range(first, first + count):
            o.frame(seq.images[k], seq.u01[k])
        dt = time.perf_counter() - t0
        return count / dt, dt, "port", threads, f"oracle restatement, dense mode (g++ -O3 -march=x86-64-v3), {threads} thread(s)"
    from oracle import ref_py as R

    o.set_options(O.Q_ALL, sparse=True, fast_corr=True, warp_patches=False)  # fast mode: only supplies each frame's match list
    for i in range(N):
        init = np.zeros((41, 41), np.uint8)
        init[14:27, 14:27] = scene.templates[i]
        r.add_feature(0, init, None, scene.x0[:3], np.eye(3), scene.uv0[i])
    r.set_state(scene.x0, scene.P0)
    dt = 0.0
    for k in range(first + count):
        o.frame(seq.images[k], seq.u01[k])
        fo = o.features()
        draws = np.minimum((seq.u01[k] * R.RAND_MAX).astype(np.int64), R.RAND_MAX - 1).astype(np.int32)
        r.set_draws(draws)
        t0 = time.perf_counter()
        r.reset_flags()  # Map::map_management's per-frame reset (src/Map.cpp:34-55); the synthetic map is fixed
        r.ekf_prediction()
        r.search_ic_matches(seq.images[k])
        t1 = time.perf_counter()
        r.set_matches(fo["z"], fo["ic"])  # untimed: the workload's matches
        t2 = time.perf_counter()
        r.ransac_hypotheses()
        r.update_li()
        r.rescue_hi()
        r.update_hi()
        t3 = time.perf_counter()
        if k >= first:
            dt += (t1 - t0) + (t3 - t2)
    xr, _ = r.get_state()
    xo, _ = o.get_state()
    assert np.allclose(xr[:13], xo[:13], rtol=1e-6, atol=1e-9), "reference and oracle trajectories diverged"
    return count / dt, dt, "reference", 1, ("the reference's own src/{ExtendKF,Tracking,Converter,Map}.cpp (compiled unmodified, -O3, over this repo's "
                                            "stand-in Eigen/OpenCV headers; single threaded like the reference)")


def make_c2(seed, frames, n_u01=1000):
    scene = synth.make_scene(N=100, seed=seed)
    seq = synth.make_sequence(scene, T=frames, seed=seed + 1, n_u01=n_u01, u01_seed=42 + seed)
    return scene, seq


def new_gpu_filter(scene, batch=1, device=0, quirks=0x7, x=None, P=None, max_features=None):
    from ransac_slam_b200 import capi

    g = capi.Filter(scene.cam.as9(), max_features or scene.N, batch=batch, device=device, quirks=quirks, std_z=scene.std_z)
    for b in range(batch):
        g.upload_state(scene.x0 if x is None else x, scene.P0 if P is None else P, b=b)
        g.upload_patches(scene.templates.astype(np.float64), b=b)
    return g


def run_frames_resident(g, d_images, d_u01, frames, l2_flush, flush_buf, stream):
    """times `frames` (list of frame indices) with CUDA events on the library's stream; returns total ms"""
    import torch

    T, rows, cols = d_images.shape[0], d_images.shape[-2], d_images.shape[-1]
    n_u01 = d_u01.shape[-1]
    img_bytes = rows * cols * (d_images.shape[1] if d_images.dim() == 4 else 1)
    u_bytes = n_u01 * 8 * (d_u01.shape[1] if d_u01.dim() == 3 else 1)
    if l2_flush:
        evs = []
        for k in frames:
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.frame((d_images.data_ptr() + k * img_bytes, rows, cols, cols), (d_u01.data_ptr() + k * u_bytes, n_u01), predict=True)
            e1.record(stream)
            evs.append((e0, e1))
        g.sync()
        return sum(a.elapsed_time(b) for a, b in evs)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in frames:
        g.frame((d_images.data_ptr() + k * img_bytes, rows, cols, cols), (d_u01.data_ptr() + k * u_bytes, n_u01), predict=True)
    e1.record(stream)
    g.sync()
    return e0.elapsed_time(e1)


def fp64_gemm_peak():
    """cuBLAS fp64 GEMM throughput measured live (MEASURED_PEAKS.json has no fp64 figure); measurement probe only."""
    import torch

    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def frame_stats(g):
    ft = g.features()
    r = g.ransac_result()
    return dict(nic=int(ft["ic"].sum()), mid=int(ft["ic"].sum()), m_li=int(ft["li"].sum()), m_hi=int(ft["hi"].sum()), hyp_run=r["hyp_run"])


def bench_c2(args, world, rank, local):
    import torch

    from oracle import oracle_py as O  # only for the cpu_baseline leg (rank 0, N == 1)

    K, W = args.steps, args.warmup
    T = W + K
    scene, seq = make_c2(1234 + rank, T)
    N, n = scene.N, scene.x0.size
    rows, cols = seq.images.shape[1:]
    # ---- device-resident throughput ------------------------------------------------------------------------------
    g = new_gpu_filter(scene, device=local)
    stream = torch.cuda.ExternalStream(g.stream)
    d_images = torch.from_numpy(seq.images).cuda()
    d_u01 = torch.from_numpy(seq.u01).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    run_frames_resident(g, d_images, d_u01, range(W), False, flush, stream)
    barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = g.launches
    ms = run_frames_resident(g, d_images, d_u01, range(W, T), True, flush, stream)
    launches = g.launches - l0
    barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    ms_max = max_over_ranks(ms, world)
    value = world * K / (ms_max * 1e-3)
    pose_resident = g.download_pose()
    # warm-L2 back-to-back figure (informational)
    g2 = new_gpu_filter(scene, device=local)
    run_frames_resident(g2, d_images, d_u01, range(W), False, flush, stream=torch.cuda.ExternalStream(g2.stream))
    ms_warm = run_frames_resident(g2, d_images, d_u01, range(W, T), False, flush, torch.cuda.ExternalStream(g2.stream))
    g2.close()
    # ---- per-kernel breakdown (separate, instrumented pass: events around every launch, no graph) --------------------
    gp = new_gpu_filter(scene, device=local)
    sp = torch.cuda.ExternalStream(gp.stream)
    run_frames_resident(gp, d_images, d_u01, range(W), False, flush, sp)
    gp.profile(True)
    stats = []
    PF = min(K, 50)
    for k in range(W, W + PF):
        with torch.cuda.stream(sp):
            flush.zero_()
        run_frames_resident(gp, d_images, d_u01, [k], False, flush, sp)
        stats.append(frame_stats(gp))
    prof = gp.profile_read()
    gp.profile(False)
    gp.close()
    st = {k: float(np.mean([s[k] for s in stats])) for k in stats[0]}
    tot = sum(v[1] for v in prof.values())
    breakdown = {k: dict(launches_per_frame=v[0] / PF, us_per_frame=1e3 * v[1] / PF, share=v[1] / tot) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    top = next(iter(breakdown))
    pk = peaks()
    roof = dict(kernel=top, share_of_step=breakdown[top]["share"])
    bound, amount = algorithmic_work(top, N, n, st)
    per_launch_us = 1e3 * prof[top][1] / prof[top][0]
    if bound == "tensor":
        fp64_peak = fp64_gemm_peak()
        per_frame_us = 1e3 * prof[top][1] / PF
        roof.update(bound="tensor", achieved=amount / (per_frame_us * 1e-6) / 1e12, peak=fp64_peak, unit="TFLOP/s",
                    note="fp64 (DMMA) work of this kernel per frame / its time per frame; peak = cuBLAS fp64 GEMM measured live (no fp64 figure in "
                         "MEASURED_PEAKS.json).  C2 is ONE 613-state filter: every kernel of the step is launch/latency bound by construction "
                         "(a frame touches ~3 MB and ~1e8 flop), so this fraction is small; the roofline fractions that mean something are "
                         "under workloads.C3 (DMMA), workloads.C4 (support sweep, HBM) and workloads.C5 (batched filters)")
    else:
        roof.update(bound="hbm", achieved=(amount or 0.0) / (per_launch_us * 1e-6) / 1e9, peak=pk["hbm_gbs"], unit="GB/s", note="peak: " + pk["source"])
    roof["frac"] = roof["achieved"] / roof["peak"] if roof.get("peak") else None
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/r01_s5_chol_small_c2.md)
    roof["traffic"] = {"k_chol_small": 158208}.get(top)
    roof["launch_us"] = per_launch_us
    # ---- end to end: host (pinned) inputs copied in, pose copied out, every frame -------------------------------------
    ge = new_gpu_filter(scene, device=local)
    h_images = torch.from_numpy(seq.images).pin_memory()
    h_u01 = torch.from_numpy(seq.u01).pin_memory()
    img_b, u_b = rows * cols, seq.u01.shape[1] * 8

    def e2e_step(k):
        ge.frame((h_images.data_ptr() + k * img_b, rows, cols, cols), (h_u01.data_ptr() + k * u_b, seq.u01.shape[1]), predict=True)
        return ge.download_pose()

    for k in range(W):
        e2e_step(k)
    barrier(world)
    t0 = time.perf_counter()
    for k in range(W, T):
        pose = e2e_step(k)
    ge.sync()
    t_e2e = time.perf_counter() - t0
    barrier(world)
    t_e2e = max_over_ranks(t_e2e, world)
    e2e = dict(value=world * K / t_e2e, unit=UNIT, h2d_bytes_per_step=img_b + u_b, d2h_bytes_per_step=13 * 8)
    assert np.allclose(pose, pose_resident, rtol=1e-9, atol=1e-12), "resident and end-to-end runs must produce the same trajectory"
    ge.close()
    g.close()
    # ---- CPU baseline: the oracle (dense, reference-faithful), 1 thread, bounded sample -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nf = min(args.cpu_frames, T)
        v, dt, kind, cores, what = cpu_frames(scene, seq, 0, nf, 1)
        cpu = dict(value=v, unit=UNIT, cores=cores, kind=kind, sample=f"first {nf} frames of the same C2 trajectory, {what}, {dt:.1f} s")
        if kind == "reference":  # the restatement beside it, same frames
            vp, dtp, _, _, whatp = cpu_frames(scene, seq, 0, min(nf, 4), 1, prefer_reference=False)
            cpu["port"] = dict(value=vp, unit=UNIT, cores=1, sample=f"first {min(nf, 4)} frames, {whatp}, {dtp:.1f} s")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_max / K, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="C2: synthetic 100-feature inverse-depth map, 320x240, bounded 1000-frame trajectory, 1-pt RANSAC + li/hi EKF update",
                            features=N, state_dim=n, quirks="reference (Q1,Q4,Q6 on)", l2="flushed between steps (256 MiB memset, untimed)",
                            per_gpu="one independent filter per GPU (replicas)" if world > 1 else "single filter", n_u01=int(seq.u01.shape[1])),
                e2e=e2e, gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu,
                kernels=breakdown, frame_stats=st, value_warm_l2=world * K / (max_over_ranks(ms_warm, world) * 1e-3))
    return line


def bench_reference(args, world, rank):
    """CPU arm: the reference's own sources (oracle/_ref) when that library exists, else the oracle port; rank 0 only."""
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    K = min(args.steps, 20)
    W = min(args.warmup, 1)
    scene, seq = make_c2(1234, W + K)
    v, dt, kind, used, what = cpu_frames(scene, seq, W, K, cores)
    sample = f"{K} frames (of the requested {args.steps}) of the C2 trajectory after {W} warm-up, {what}; host has {cores} cores"
    return dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=1e3 * dt / K, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="C2: synthetic 100-feature inverse-depth map, 320x240, bounded 1000-frame trajectory, 1-pt RANSAC + li/hi EKF update",
                            features=scene.N, state_dim=int(scene.x0.size), quirks="reference (Q1,Q4,Q6 on)"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=used, kind=kind, sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="the reference is single threaded (no OpenMP in its CMakeLists.txt); ROS, the Map feature management and visualisation are not in "
                     "the timed path; third-party arithmetic (Eigen, OpenCV) is this repository's stand-in, not the real libraries")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="c3,c4,c5")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    # stdout carries exactly ONE line, the JSON: everything libraries print while we run (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        line = bench_reference(args, world, rank)
        if line is not None:
            emit(line)
        return
    world, rank, local = dist_setup(args.gpus)
    line = bench_c2(args, world, rank, local)
    if not args.no_extras:
        try:
            import bench_extras

            line["workloads"] = bench_extras.run(args, world, rank, local)
        except Exception as e:  # extras must never cost the main line
            line["workloads"] = dict(error=repr(e))
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
