#!/usr/bin/env python
"""bench.py -- local bundle-adjustment throughput on KITTI-00-shaped windows (BASELINE.json metric).

A "step" is one full ORB-SLAM2-style local BA (robust 5 LM iterations -> chi2/depth outlier exclusion ->
10 LM iterations) over one batch of independent synthetic stereo windows of the C0 shape (20 free + 10 fixed
keyframes, ~6k points, ~70k observations each; BASELINE.json configs[0]/[4]).  Windows are the unit that shards:
every rank owns `--windows-per-gpu` windows, no data-path collective (weak scaling).

  value   observations x LM trials per second with the batch already resident in HBM (solve only)
  e2e     the same metric through the C ABI with HOST buffers: sqrtba_set_problem_batch (H2D) + solve + read-back (D2H)
  roofline  PCG matvec kernel: SURVEY 8(d)'s algorithmic bytes (216 B per free-pose stereo observation) / CUDA-event time
            per launch, beside the bytes the kernel's own layout moves (104 B: the 3x6 pose block is rebuilt from four numbers)
  cpu_baseline  the oracle (g2o Schur-LM restatement, oracle/refba.cpp) on a bounded sample, 1 thread

`--impl reference` times that CPU implementation with all host threads (one window per thread).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_MATVEC = 216.0     # canonical bytes per free-pose stereo observation of the matvec / back-substitution
LAYOUT_MATVEC = 104.0  # bytes per free-pose observation the kernels actually stream (4 geometry rows + 9 rows of Q1)
METRIC = "local-BA observations/s (KITTI-00-shaped stereo windows, observations x LM trials per second)"
UNIT = "obs/s"


def load_pkg():
    name = "sqrtlm_slam_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "sqrtlm-slam_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_batch(pkg, n_windows: int, seed0: int):
    wins = [pkg.synth.config_c0(seed0 + i) for i in range(n_windows)]
    prob, pp, tp, op = pkg.synth.concat_windows(wins)
    return wins, prob, pp, tp, op


def pinned_copy(arr):
    """Host staging buffers in pinned memory (the e2e leg copies from these)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    try:
        t = t.pin_memory()
    except Exception:
        pass
    return t


def trials_times_obs(ba, wins):
    tot = 0
    ntr = 0
    for i, w in enumerate(wins):
        n = len(ba.trace(i))
        tot += n * w.n_obs
        ntr += n
    return tot, ntr


def cpu_solve_windows(wins, threads_total: int, keep: bool = False):
    """Solve windows with the oracle, one window per worker thread (ctypes releases the GIL). Returns
    (seconds, sum(n_obs * trials), sum(trials)[, per-window (trace, poses, outliers)])."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import refba
    refba.lib()

    def one(w):
        r = refba.RefBA(w, threads=1)
        r.solve_local(0)
        tr = r.trace()
        return len(tr) * w.n_obs, len(tr), ((tr, r.poses(), r.outliers()) if keep else None)

    t0 = time.perf_counter()
    if threads_total <= 1:
        res = [one(w) for w in wins]
    else:
        with ThreadPoolExecutor(max_workers=threads_total) as ex:
            res = list(ex.map(one, wins))
    dt = time.perf_counter() - t0
    out = (dt, sum(r[0] for r in res), sum(r[1] for r in res))
    return out + ([r[2] for r in res],) if keep else out


def quat_angle(qa, qb):
    d = np.abs(np.sum(qa * qb, axis=-1)).clip(0, 1)
    rel_v = np.linalg.norm(qa[..., :3] * qb[..., 3:4] - qb[..., :3] * qa[..., 3:4] - np.cross(qa[..., :3], qb[..., :3]), axis=-1)
    return 2 * np.arctan2(rel_v, d)


def parity_report(tg, pg, fg, tr, pr, fr, free):
    """north_star's tolerances on one window: identical trial sequence, per-trial cost <= 1e-6 relative, pose RMS
    <= 1e-5 m / 1e-6 rad, identical outlier flags.  Returns (ok, worst cost error, pose RMS m, pose RMS rad)."""
    same = len(tg) == len(tr) and np.array_equal(tg[:, [0, 1, 2, 7]], tr[:, [0, 1, 2, 7]])
    cost = float(np.max(np.abs(tg[:, 5] - tr[:, 5]) / np.abs(tr[:, 5]))) if same else float("inf")
    t_rms = float(np.sqrt(np.mean(np.sum((pg[free, :3] - pr[free, :3]) ** 2, axis=1))))
    r_rms = float(np.sqrt(np.mean(quat_angle(pg[free, 3:], pr[free, 3:]) ** 2)))
    flags = fg is None or bool(np.array_equal(fg, fr))
    ok = bool(same and cost <= 1e-6 and t_rms <= 1e-5 and r_rms <= 1e-6 and flags)
    return ok, cost, t_rms, r_rms


WORKLOAD = ("C4: batch of independent KITTI-00-shaped stereo local-BA windows (C0 shape: 20 free + 10 fixed keyframes, "
            "~6k points, ~70k observations), two-pass 5+10 LM with chi2 outlier exclusion")


def bench_config(args, n_obs_rank=None, nobs_all=None):
    cfg = {"workload": WORKLOAD, "windows_per_gpu": args.windows_per_gpu, "l2_policy": "inputs_larger_than_l2",
           "pcg_rtol": "auto (1e-7 for two-pass local BA on windows of up to 128 free keyframes, 1e-8 for global BA; include/sqrtba.h)",
           "pcg_mode": args.pcg_mode}
    if n_obs_rank is not None:
        cfg["observations_per_gpu"] = n_obs_rank
        cfg["observations_total"] = int(nobs_all)
    return cfg


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; g2o itself cannot be built in this image)
    on the same workload shape, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = load_pkg()
    cores = os.cpu_count() or 1
    # the GPU arm's step is args.windows_per_gpu windows (seeds 0..W-1 on rank 0); a CPU step is the first n_sample of
    # them, sized from a timed warm-up so that `--steps` steps end within a few minutes on this box's cores
    first = [pkg.synth.config_c0(i) for i in range(min(cores, args.windows_per_gpu))]
    t_w = cpu_solve_windows(first, cores)[0]           # ~ one window per core: seconds per window and thread
    if args.cpu_windows > 0:
        n_sample = args.cpu_windows
    else:
        budget_s = 150.0
        n_sample = int(budget_s * cores / (max(args.steps, 1) * max(t_w, 1e-3)))
        n_sample = max(min(n_sample, args.windows_per_gpu), min(cores, args.windows_per_gpu))
    wins = first + [pkg.synth.config_c0(i) for i in range(len(first), n_sample)]
    wins = wins[:n_sample]
    tot_t = tot_work = tot_tr = 0
    for _ in range(args.steps):
        dt, work, ntr = cpu_solve_windows(wins, cores)
        tot_t += dt
        tot_work += work
        tot_tr += ntr
    value = tot_work / tot_t
    sample = (f"the first {n_sample} of the batch's {args.windows_per_gpu} windows per step "
              f"({sum(w.n_obs for w in wins)} observations), full two-pass local BA each, one window per thread")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": bench_config(args),
        "sample_windows_per_step": n_sample,
        "lm_iters_per_s": tot_tr / tot_t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r02_matvec_ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def single_window_leg(pkg, prob, local, label, reps=5):
    """Latency of ONE window -- the call shape of Optimizer::LocalBundleAdjustment: device time of the two-pass solve
    (best of `reps` after two warm-up solves) and wall time of the whole C-ABI call sequence from host buffers."""
    ba = pkg.SqrtBA(device=local)
    ba.set_problem(prob)
    ms = []
    st = None
    for _ in range(reps):
        ba.reset_state()
        st = ba.solve_local()
        ms.append(st["ms_total"])
    tr = ba.trace()
    best = min(ms[2:])
    out = {"workload": label, "n_obs": prob.n_obs, "n_free_poses": prob.n_free, "ms_per_local_ba": best,
           "lm_trials": len(tr), "lm_iters_per_s": len(tr) / (best * 1e-3), "obs_per_s": len(tr) * prob.n_obs / (best * 1e-3),
           "cg_iters": st["cg_iters_total"], "persistent_pcg": st["persistent_pcg"], "kernel_launches": st["kernel_launches"]}
    try:  # the whole C-ABI call sequence of one LocalBundleAdjustment from host buffers (wall clock, best of reps)
        wall = []
        for _ in range(reps):
            t0 = time.perf_counter()
            ba.set_problem(prob)
            ba.solve_local()
            ba.poses(), ba.points(), ba.outliers()
            wall.append(time.perf_counter() - t0)
        out["ms_per_call_set_solve_get"] = 1e3 * min(wall[1:])
    except Exception as exc:  # an extra, never fatal for the bench line
        out["ms_per_call_set_solve_get_error"] = repr(exc)
    res = (ba.trace(), ba.poses(), ba.outliers())
    ba.close()
    return out, res


def seeds_median_leg(pkg, local, make, seeds=(0, 1, 2, 3, 4)):
    """median over seeds 0-4 (device time of the two-pass solve, best of 3 after a warm-up each)."""
    ba = pkg.SqrtBA(device=local)
    ms, iters = [], []
    for seed in seeds:
        prob = make(seed)
        ba.set_problem(prob)
        t = []
        for _ in range(4):
            ba.reset_state()
            st = ba.solve_local()
            t.append(st["ms_total"])
        ms.append(min(t[1:]))
        iters.append(len(ba.trace()) / (ms[-1] * 1e-3))
    ba.close()
    return {"seeds": list(seeds), "ms_per_local_ba": [float(x) for x in ms], "ms_per_local_ba_median": float(np.median(ms)),
            "lm_iters_per_s_median": float(np.median(iters))}


def oracle_local(prob, threads):
    from oracle import refba
    r = refba.RefBA(prob, threads=threads)
    t0 = time.perf_counter()
    r.solve_local(0)
    return time.perf_counter() - t0, (r.trace(), r.poses(), r.outliers())


def pose_only_leg(pkg, local, cpu_ok):
    """Optimizer::PoseOptimization (tracking thread, every frame): one 1500-point frame and a batch of 64 relocalisation
    candidates through sqrtba_pose_opt (host buffers in, pose + flags out: the call IS end to end)."""
    import numpy as np
    ba = pkg.SqrtBA(device=local)
    p0, cam, xyz, meas, truth = pkg.synth.frame_problem(seed=5, n_points=1500)
    dev, wall = [], []
    for _rep in range(8):
        t0 = time.perf_counter()
        gp, gf, gi, st = ba.pose_opt([0, len(xyz)], p0, cam, xyz, meas)
        wall.append(time.perf_counter() - t0)
        dev.append(st["ms_total"])
    out = {"workload": "one frame, 1500 matched map points, 4 x optimize(10) with chi2 re-classification (g2oOptimizer.cc:385-559)",
           "ms_per_frame_device": min(dev[2:]), "ms_per_frame_call": 1e3 * min(wall[2:]), "inliers": int(gi[0]),
           "lm_trials": len(ba.pose_opt_trace(0))}
    # the same frame with this fork's lidar block (g2oOptimizer.cc:560-640): 800 flat + 200 sharp points against a
    # 20 000-point local lidar map, exact nearest neighbour on the device, fifth optimize(10) over visual + lidar edges
    ld = pkg.synth.frame_lidar(truth, seed=5)
    dev = []
    for _k in range(6):
        lp, lf, li, lnm, st = ba.pose_opt_lidar(p0, cam, xyz, meas, ld)
        dev.append(st["ms_total"])
    out["with_lidar_block"] = {"ms_per_frame_device": min(dev[2:]), "flat_matches": lnm[0], "corner_matches": lnm[1],
                               "map_points": int(len(ld.map_xyz)), "kernel_launches": st["kernel_launches"], "inliers": li}
    frames = [pkg.synth.frame_problem(seed=100 + k, n_points=1500) for k in range(64)]
    ptr = np.concatenate([[0], np.cumsum([len(f[2]) for f in frames])])
    P0 = np.stack([f[0] for f in frames]); CAM = np.stack([f[1] for f in frames])
    XYZ = np.concatenate([f[2] for f in frames]); MEAS = np.concatenate([f[3] for f in frames])
    wall = []
    for _ in range(5):
        t0 = time.perf_counter()
        ba.pose_opt(ptr, P0, CAM, XYZ, MEAS)
        wall.append(time.perf_counter() - t0)
    out["batch_64_frames_ms_per_call"] = 1e3 * min(wall[1:])
    out["frames_per_s_batched"] = 64 / min(wall[1:])
    if cpu_ok:
        from oracle import refba
        t0 = time.perf_counter()
        rp, rf, ri, _ = refba.pose_opt(p0, cam, xyz, meas)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1e3 * dt, "unit": "ms per frame (lower is better)", "cores": 1, "kind": "port",
                               "sample": "the same frame"}
        out["parity_vs_oracle"] = {"ok": bool(ri == int(gi[0]) and np.array_equal(gf, rf) and np.abs(gp[0] - rp).max() <= 1e-7)}
        t0 = time.perf_counter()
        rp, rf, ri, _t, rnm = refba.pose_opt_lidar(p0, cam, xyz, meas, ld)
        out["with_lidar_block"]["cpu_baseline_ms"] = 1e3 * (time.perf_counter() - t0)
        out["with_lidar_block"]["parity_vs_oracle"] = {"ok": bool(ri == li and np.array_equal(lf, rf) and rnm == lnm and np.abs(lp - rp).max() <= 1e-6)}
    ba.close()
    return out


def sim3_leg(pkg, local, cpu_ok):
    """Optimizer::OptimizeSim3 (LoopClosing::ComputeSim3, per loop candidate): one candidate pair with 150 matches and a
    batch of 16 candidates through sqrtba_optimize_sim3 (host buffers in, S12 + flags out: the call IS end to end)."""
    import numpy as np
    ba = pkg.SqrtBA(device=local)
    c = pkg.synth.sim3_pair(seed=5, n_matches=150)
    dev, wall = [], []
    for _ in range(8):
        t0 = time.perf_counter()
        S, keep, n_in, st = ba.optimize_sim3([0, len(c[2])], c[0], c[1], c[2], c[3], c[4], 10.0, False)
        wall.append(time.perf_counter() - t0)
        dev.append(st["ms_total"])
    out = {"workload": "one keyframe pair, 150 matches (300 edges), optimize(5) + chi2 test + optimize(10), numeric Jacobians (g2oOptimizer.cc:1560-1796)",
           "ms_per_pair_device": min(dev[2:]), "ms_per_pair_call": 1e3 * min(wall[2:]), "inliers": int(n_in[0]),
           "lm_trials": len(ba.optimize_sim3_trace(0))}
    cases = [pkg.synth.sim3_pair(seed=200 + k, n_matches=150) for k in range(16)]
    ptr = np.concatenate([[0], np.cumsum([len(k[2]) for k in cases])])
    arrs = [np.stack([k[0] for k in cases]), np.stack([k[1] for k in cases])] + [np.concatenate([k[i] for k in cases]) for i in (2, 3, 4)]
    wall = []
    for _ in range(5):
        t0 = time.perf_counter()
        ba.optimize_sim3(ptr, *arrs, 10.0, False)
        wall.append(time.perf_counter() - t0)
    out["batch_16_pairs_ms_per_call"] = 1e3 * min(wall[1:])
    if cpu_ok:
        from oracle import refba
        t0 = time.perf_counter()
        So, keep_o, nin_o, _ = refba.optimize_sim3(c[0], c[1], c[2], c[3], c[4], 10.0, False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1e3 * dt, "unit": "ms per pair (lower is better)", "cores": 1, "kind": "port",
                               "sample": "the same pair"}
        out["parity_vs_oracle"] = {"ok": bool(nin_o == int(n_in[0]) and np.array_equal(keep, keep_o) and np.abs(S[0] - So).max() <= 1e-6)}
    ba.close()
    return out


def essential_graph_leg(pkg, local, cpu_ok):
    """Optimizer::OptimizeEssentialGraph at KITTI-00 length: 1500 keyframes on a drifting loop, spanning tree +
    covisibility + loop edges, Levenberg with lambda_0 = 1e-16, 20 iterations (g2oOptimizer.cc:1212-1460)."""
    ba = pkg.SqrtBA(device=local)
    v0, fixed, edges, meas = pkg.synth.pose_graph(5, n_kf=1500, fix_scale=True, n_loop=10)
    ms = []
    for _ in range(3):
        V, tr, st = ba.pose_graph(v0, fixed, True, edges, meas, iters=20)
        ms.append(st["ms_total"])
    acc = tr[tr[:, 7] == 1]
    out = {"workload": f"{len(v0)} keyframes, {len(edges)} Sim3 edges, fixed scale", "ms_per_call": min(ms[1:]), "lm_trials": len(tr),
           "chi2_initial": float(tr[0, 4]), "chi2_final": float(acc[-1, 5]) if len(acc) else None,
           "cpu_baseline": None,
           "note": "no CPU figure at this size: the oracle's dense LDL^T of the 10 493-unknown system would take hours "
                   "(the reference uses a sparse Cholesky); parity is tested on 24-160 keyframe graphs"}
    ba.close()
    return out


def run_extras(pkg, torch, dist, world, rank, local, barrier, args):
    """The other shapes of BASELINE.json's metric, beside the headline batch:
    (1) C0 / C1 single-window latency and LM iterations/s on rank 0;
    (2) C2 (100 keyframes, ~600k observations, one GPU): latency, per-kernel GB/s against the HBM peak, the rate of the
        persistent PCG kernel's matvec phase, oracle beside it;
    (3) C3 global BA, landmarks sharded over all `world` ranks (strong scaling): time-to-converge of
        solve_global(10 iterations, non-robust), the reference's loop-closing call (LoopClosing.cc:987-991); at N=1 the
        oracle's time-to-converge on the host cores (1 thread and all cores) and a parity check against it."""
    out = {}
    cpu_ok = rank == 0 and world == 1 and not args.skip_cpu_baseline
    cores = os.cpu_count() or 1
    peaks, _ = measured_peaks()
    if rank == 0:
        for key, prob, label in (("single_window", pkg.synth.config_c0(0), "C0: one KITTI-00-shaped stereo window, two-pass 5+10 local BA"),
                                 ("mono_window", pkg.synth.config_c1(0), "C1: the same window in monocular mode (EdgeSE3ProjectXYZ), gauge fixed by the first window keyframe")):
            leg, res = single_window_leg(pkg, prob, local, label)
            if cpu_ok:
                dt, ref = oracle_local(prob, 1)
                ok, cost, t_rms, r_rms = parity_report(*res, *ref, prob.pose_fixed == 0)
                dta, _ = oracle_local(prob, cores)
                leg["cpu_baseline"] = {"value": len(ref[0]) * prob.n_obs / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                       "ms_per_local_ba": 1e3 * dt, "lm_iters_per_s": len(ref[0]) / dt,
                                       "ms_per_local_ba_all_cores": 1e3 * dta, "all_cores": cores,
                                       "sample": "the same window, full two-pass local BA"}
                leg["parity_vs_oracle"] = {"ok": ok, "cost_rel_err_max": cost, "pose_t_rms_m": t_rms, "pose_r_rms_rad": r_rms}
            leg["seeds_0_4"] = seeds_median_leg(pkg, local, pkg.synth.config_c0 if key == "single_window" else pkg.synth.config_c1)
            out[key] = leg
        # ---- C2
        prob = pkg.synth.config_c2(0)
        leg, res = single_window_leg(pkg, prob, local, "C2: large stereo window, 100 keyframes (99 free), ~50k points, ~600k observations, two-pass 5+10 local BA", reps=4)
        free_obs = int((prob.pose_fixed[prob.obs_pose] == 0).sum())
        ba = pkg.SqrtBA(device=local, stage_timing=True)
        ba.set_problem(prob)
        ba.solve_local()
        ba.reset_state()
        st = ba.solve_local()
        us_cg = 1e3 * st["ms_pcg"] / max(st["cg_iters_total"], 1)
        stage_ms = {"k_matvec_pipe": ba.time_stage(0, 3, 20), "k_linearize_pipe": ba.time_stage(1, 2, 10),
                    "k_qr_pipe2": ba.time_stage(2, 2, 10), "k_backsub": ba.time_stage(4, 2, 10)}
        ba.close()
        n_lm = prob.n_point
        # SURVEY 8(d) / DESIGN.md section 4 figures (canonical layout), as in round 1
        alg = {"k_matvec_pipe": free_obs * ALG_MATVEC, "k_linearize_pipe": prob.n_obs * 288.0,
               "k_qr_pipe2": prob.n_obs * 312.0 + n_lm * 72.0, "k_backsub": free_obs * ALG_MATVEC + n_lm * 168.0}
        leg["roofline"] = {"bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "kernels": {k: {"ms": v, "algorithmic_bytes": alg[k], "achieved": alg[k] / (v * 1e-3) / 1e9,
                                           "frac": alg[k] / (v * 1e-3) / 1e9 / peaks["hbm_gbs"]} for k, v in stage_ms.items()},
                           "persistent_pcg": {"us_per_cg_iteration": us_cg, "achieved": free_obs * ALG_MATVEC / (us_cg * 1e-6) / 1e9,
                                              "frac": free_obs * ALG_MATVEC / (us_cg * 1e-6) / 1e9 / peaks["hbm_gbs"],
                                              "achieved_moved": free_obs * LAYOUT_MATVEC / (us_cg * 1e-6) / 1e9,
                                              "chunk_precond": st["chunk_precond"], "coarse_level": st["coarse_level"],
                                              "note": "PCG stage time / CG iterations: matvec over all tiles + grid barriers + vector update; a "
                                                      "window of 64 or more free keyframes runs the big-window kernel with the two-level "
                                                      "preconditioner (half the iterations), whose build per LM iteration is inside this figure"},
                           "note": "SURVEY 8(d) bytes; the layout moves 104 B per free observation, so one CG iteration touches "
                                   "~61 MB: L2-resident on this GPU (126 MB) -- these are L2 rates, not HBM rates"}
        if cpu_ok:
            dt, ref = oracle_local(prob, 1)
            dta, _ = oracle_local(prob, cores)
            ok, cost, t_rms, r_rms = parity_report(*res, *ref, prob.pose_fixed == 0)
            leg["cpu_baseline"] = {"value": len(ref[0]) * prob.n_obs / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                   "ms_per_local_ba": 1e3 * dt, "ms_per_local_ba_all_cores": 1e3 * dta, "all_cores": cores,
                                   "sample": "the same window, full two-pass local BA"}
            leg["parity_vs_oracle"] = {"ok": ok, "cost_rel_err_max": cost, "pose_t_rms_m": t_rms, "pose_r_rms_rad": r_rms}
        out["large_window"] = leg
    if rank == 0:
        out["pose_only"] = pose_only_leg(pkg, local, cpu_ok)
        out["essential_graph"] = essential_graph_leg(pkg, local, cpu_ok)
        out["sim3_candidates"] = sim3_leg(pkg, local, cpu_ok)
    prob = pkg.synth.config_c3(0, scale=args.gba_scale, n_kf=max(int(1500 * args.gba_scale), 160))
    shard, _, _ = pkg.multi.shard_by_landmark(prob, rank, world)

    def gba_solve(pcg_mode):
        h = pkg.SqrtBA(device=local, pcg_mode=pcg_mode)
        if world > 1:
            pkg.multi.init_comm(h, rank, world)
        h.set_problem(shard)
        ts, s_last = [], None
        for _ in range(3):
            h.reset_state()
            barrier()
            t0 = time.perf_counter()
            s_last = h.solve_global(10, False)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        tt = torch.tensor([min(ts[1:])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return h, s_last, tt

    # A/B first: the 6x6 block-Jacobi preconditioner of the local-BA path on the same map (pcg_mode 5)
    ba6, st6, tsec6 = gba_solve(5)
    tr6 = ba6.trace()
    ab = {"time_to_converge_s": float(tsec6.item()), "cg_iters_total": int(tr6[:, 8].sum()), "final_chi2": float(tr6[-1, 5]),
          "note": "same solve with pcg_mode = 5: 6x6 block-Jacobi blocks instead of the 20-keyframe chunk blocks + coarse level"}
    ba6.close()
    ba, st, tsec = gba_solve(0)
    tr = ba.trace()
    free_obs = int((prob.pose_fixed[prob.obs_pose] == 0).sum())
    cg_total = max(int(tr[:, 8].sum()), 1)
    us_iter = 1e6 * float(tsec.item()) / cg_total
    out["global_ba"] = {"workload": f"C3: {prob.n_pose} keyframes on a loop, {prob.n_point} points, {prob.n_obs} observations, "
                                    "10 LM iterations, non-robust; landmarks sharded over the ranks",
                        "n_gpus": world, "scaling": "strong", "time_to_converge_s": float(tsec.item()),
                        "lm_trials": len(tr), "cg_iters_total": int(tr[:, 8].sum()), "final_chi2": float(tr[-1, 5]),
                        "persistent_pcg": st["persistent_pcg"], "peer_exchange": st["peer_exchange"],
                        "preconditioner": {"chunk_blocks_20_keyframes": st["chunk_precond"], "coarse_level": st["coarse_level"],
                                           "block_jacobi_6x6_ab": ab},
                        "us_per_cg_iteration_incl_everything": us_iter,
                        "matvec_algorithmic_bytes_all_ranks": free_obs * ALG_MATVEC,
                        "matvec_layout_bytes_all_ranks": free_obs * LAYOUT_MATVEC,
                        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"] * world,
                                     "achieved": free_obs * ALG_MATVEC / (us_iter * 1e-6) / 1e9,
                                     "frac": free_obs * ALG_MATVEC / (us_iter * 1e-6) / 1e9 / (peaks["hbm_gbs"] * world),
                                     "note": "whole time-to-converge / CG iterations, i.e. linearise, QR, the preconditioner "
                                             "build, barriers and the NVLink exchange all charged to the matvec's algorithmic "
                                             "bytes; peak = N x one GPU.  The two-level preconditioner cuts the iterations ~6x, so "
                                             "the fixed per-trial work weighs more in this figure than with the 6x6 blocks"}}
    if cpu_ok:  # the reference's CPU algorithm on the same map: time-to-converge, 1 thread (faithful) and all cores
        from oracle import refba
        pg = ba.poses()
        res = {}
        for thr in ([1, cores] if cores > 1 else [1]):
            r = refba.RefBA(prob, threads=thr)
            t0 = time.perf_counter()
            r.solve_global(10, False)
            res[thr] = time.perf_counter() - t0
        ok, cost, t_rms, r_rms = parity_report(tr, pg, None, r.trace(), r.poses(), None, prob.pose_fixed == 0)
        out["global_ba"]["cpu_baseline"] = {"value": res[1], "unit": "s (time-to-converge, lower is better)", "cores": 1,
                                            "kind": "port", "all_cores_value": res.get(cores), "all_cores": cores,
                                            "sample": "the same C3 map, solve_global(10 iterations, non-robust), whole solve"}
        out["global_ba"]["parity_vs_oracle"] = {"ok": ok, "cost_rel_err_max": cost, "pose_t_rms_m": t_rms, "pose_r_rms_rad": r_rms}
    ba.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sqrtba", choices=["sqrtba", "reference"])
    ap.add_argument("--windows-per-gpu", type=int, default=256)
    ap.add_argument("--cpu-windows", type=int, default=0, help="windows in the CPU sample (0 = auto)")
    ap.add_argument("--pcg-mode", type=int, default=0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="skip the single-window latency and global-BA legs")
    ap.add_argument("--gba-scale", type=float, default=1.0, help="scale of the C3 global-BA leg (1.0 = 1500 KFs, 3M obs)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pkg = load_pkg()
    pkg.capi.lib()  # fails loudly if the CUDA library is missing
    W = args.windows_per_gpu
    wins, prob, pp, tp, op = make_batch(pkg, W, seed0=rank * 100000)
    n_obs_rank = prob.n_obs
    free_obs = int((prob.pose_fixed[prob.obs_pose] == 0).sum())

    ba = pkg.SqrtBA(device=local, pcg_mode=args.pcg_mode)
    ba.set_problem_batch(prob, pp, tp, op)

    # ---- resident-input leg: value -------------------------------------------------------------
    launches = 0
    for _ in range(args.warmup):
        ba.reset_state()
        ba.solve_local()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    stats = None
    for _ in range(args.steps):
        ba.reset_state()
        stats = ba.solve_local()
        dev_ms += stats["ms_total"]
        launches += stats["kernel_launches"]
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    work, ntr = trials_times_obs(ba, wins)       # per step (identical every step: same inputs)
    sec = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(work), float(ntr), float(n_obs_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    sec_max = float(sec.item())
    work_all, ntr_all, nobs_all = (float(x) for x in tot.tolist())
    value = work_all * args.steps / sec_max

    # ---- roofline of the dominant kernel (PCG matvec), full batch active, inputs >> L2 -----------
    ms_matvec = ba.time_stage(0, warmup=3, reps=20)
    ms_lin = ba.time_stage(1, warmup=2, reps=5)
    ms_qr = ba.time_stage(2, warmup=2, reps=5)
    peaks, peak_kind = measured_peaks()
    alg_bytes = free_obs * ALG_MATVEC          # the figure roofline.achieved is defined on
    moved_bytes = free_obs * LAYOUT_MATVEC     # what this kernel's operand layout streams per launch
    achieved = alg_bytes / (ms_matvec * 1e-3) / 1e9
    tr_info = load_traffic()
    traffic = None
    if tr_info and tr_info.get("dram_bytes_per_free_obs"):
        # per launch like `achieved`: the capture's DRAM bytes per free-pose observation x this launch's observations
        traffic = tr_info["dram_bytes_per_free_obs"] * free_obs
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": (tr_info or {}).get("source"),
                "peak_kind": peak_kind, "kernel": "k_matvec_pipe<2,false>",
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_matvec,
                "moved": {"layout_bytes_per_launch": moved_bytes, "achieved": moved_bytes / (ms_matvec * 1e-3) / 1e9,
                          "frac": moved_bytes / (ms_matvec * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "note": "the operand stores {x/z, y/z, 1/z, w} + Q1 (104 B per free-pose observation) and rebuilds "
                                  "the 3x6 pose block in registers: DRAM traffic is about half of the algorithmic figure, and "
                                  "the kernel is bound by the per-tile dependency chain (issue slots 58 % busy at 20 warps/SM, "
                                  "profiles/r02_matvec_final_ncu_full.txt), no longer by HBM"},
                "note": "achieved = SURVEY 8(d)'s 216 B per free-pose stereo observation (Jp 3x6 + Q1 3x3, FP64) x the "
                        "launch's observations / time, as in round 1; observations of fixed keyframes have no pose columns",
                "other_kernels_ms": {"k_linearize": ms_lin, "k_qr": ms_qr},
                "pcg_share_of_step": None}

    # share of the timed step spent in the dominant kernel, from event pairs around every launch INSIDE the timed solves
    # (launches late in a PCG solve only touch the windows that have not converged, so their average is shorter than
    # the full-batch launch the roofline is quoted on)
    roofline["pcg_share_of_step"] = stats["ms_matvec"] / max(stats["ms_total"], 1e-9)
    roofline["ms_per_launch_inside_timed_steps"] = stats["ms_matvec"] / max(stats["cg_iters_total"], 1)

    # ---- end-to-end leg through the C ABI with host buffers -------------------------------------
    host = [pinned_copy(a) for a in (prob.pose_qt, prob.pose_fixed, prob.cam, prob.point_xyz, prob.obs_pose,
                                     prob.obs_point, prob.obs_meas)]
    hprob = pkg.synth.Problem(*[t.numpy() for t in host])
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = prob.n_pose * 7 * 8 + prob.n_point * 3 * 8 + prob.n_obs
    # host-side preprocessing threads of set_problem: the ranks of one box share its cores
    host_threads = max(1, (os.cpu_count() or 1) // max(int(os.environ.get("LOCAL_WORLD_SIZE", world)), 1))
    ba2 = pkg.SqrtBA(device=local, pcg_mode=args.pcg_mode, host_threads=host_threads)
    # results land in pinned host buffers too (the caller of the C ABI owns them; pinned = plain DMA)
    o_pose = pinned_copy(np.empty((prob.n_pose, 7))).numpy()
    o_point = pinned_copy(np.empty((prob.n_point, 3))).numpy()
    o_flag = pinned_copy(np.empty(prob.n_obs, np.uint8)).numpy()
    for _ in range(min(args.warmup, 1)):
        ba2.set_problem_batch(hprob, pp, tp, op)
        ba2.solve_local()
        ba2.poses(o_pose); ba2.points(o_point); ba2.outliers(o_flag)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ba2.set_problem_batch(hprob, pp, tp, op)
        st2 = ba2.solve_local()
        launches += st2["kernel_launches"]
        ba2.poses(o_pose); ba2.points(o_point); ba2.outliers(o_flag)
    barrier()
    t1 = time.perf_counter()
    sec2 = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sec2, op=dist.ReduceOp.MAX)
    e2e_value = work_all * args.steps / float(sec2.item())
    ba2.close()

    # ---- the other two shapes of BASELINE.json's metric (reported beside the headline, not part of `value`) ----
    extras = {}
    if not args.skip_extras:
        extras = run_extras(pkg, torch, dist, world, rank, local, barrier, args)

    # ---- CPU baseline (oracle port, 1 thread, bounded sample) on rank 0 at N=1 -------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        n_cpu = args.cpu_windows if args.cpu_windows > 0 else 32
        dt, wk, _, refs = cpu_solve_windows(wins[:n_cpu], 1, keep=True)
        cpu = {"value": wk / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {n_cpu} windows of the batch, full two-pass local BA each, {dt:.1f} s"}
        # the oracle solved these windows anyway: compare them with what the GPU produced for the same windows in the
        # last timed step (trial sequence, per-trial cost, pose RMS, outlier flags -- north_star's tolerances)
        Pg, Fg = ba.poses(), ba.outliers()
        bad, worst = [], [0.0, 0.0, 0.0]
        for i, (w, (rt, rp, rf)) in enumerate(zip(wins[:n_cpu], refs)):
            ok, cost, t_rms, r_rms = parity_report(ba.trace(i), Pg[pp[i]:pp[i + 1]], Fg[op[i]:op[i + 1]], rt, rp, rf, w.pose_fixed == 0)
            worst = [max(worst[0], cost), max(worst[1], t_rms), max(worst[2], r_rms)]
            if not ok:
                bad.append(i)
        parity = {"parity_checked_windows": n_cpu, "windows_failing": bad, "cost_rel_err_max": worst[0],
                  "pose_t_rms_m_max": worst[1], "pose_r_rms_rad_max": worst[2],
                  "tolerances": "identical trial sequence and outlier flags, cost 1e-6, pose RMS 1e-5 m / 1e-6 rad"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args),
            "observations_per_gpu": n_obs_rank, "observations_total": int(nobs_all),
            "lm_iters_per_s": ntr_all * args.steps / sec_max,
            "windows_per_s": W * world * args.steps / sec_max,
            "ms_per_step_device_events": dev_ms / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * float(sec2.item()) / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_checked_windows": (parity or {}).get("parity_checked_windows", 0),
            "parity": parity,
            "solve_stats_last_step": stats,
            "single_window": extras.get("single_window"),
            "mono_window": extras.get("mono_window"),
            "large_window": extras.get("large_window"),
            "pose_only": extras.get("pose_only"),
            "essential_graph": extras.get("essential_graph"),
            "sim3_candidates": extras.get("sim3_candidates"),
            "global_ba": extras.get("global_ba"),
        }
        print(json.dumps(line), flush=True)
    ba.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
