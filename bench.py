#!/usr/bin/env python
"""bench.py — LM iterations/sec of the calibration solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2] [--impl reference]

A "step" is one trust-region (LM) iteration over the whole synthetic problem: solve the damped
arrow system, update the state, evaluate residuals + Jacobians at the trial point, rebuild the
block normal equations, accept/reject.  One JSON line is printed by rank 0.

* value      — K iterations / device time (CUDA events on the library's launch stream, max over
               ranks); inputs resident in HBM; L2 flushed before every iteration (config 2's inputs
               are 12 MB, far below the 126 MB L2) unless --no-flush.  The single-/multi-GPU vision
               solve runs in the persistent kernel: one launch per iteration in the flushed region,
               one per solve otherwise (config.value_no_flush).
* e2e        — the same metric through the C-ABI with HOST buffers: upload (set_* calls), K
               iterations, state read-back, all inside the timed region (wall clock).
* roofline   — dominant kernel: useful FP64 flops per launch / its duration against the FP64
               throughput measured live (the path is FP64-pipe bound); HBM view beside it.
* cpu_baseline — the CPU oracle (port of the reference's Ceres path) on a bounded sample.
* --impl reference — the oracle port on all host threads (Ceres cannot be built here).
"""
from __future__ import annotations

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

from vicalib_b200 import synth  # noqa: E402

METRIC = "lm_iterations_per_sec"
UNIT = "iterations/s"


_RESULT_FD = None


def _emit(out):
    line = (json.dumps(out) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes_per_obs(K):
    """the fused variant no longer materialises J, so an iteration's algorithmic
    traffic is the observation read (44 B/corner); block outputs are <1 %."""
    return 44


def stage_bytes(stage, n_obs, K, n_groups=0, fused=True):
    """Algorithmic HBM bytes of one launch of a stage."""
    if stage == "eval_reproj":  # two-pass variant only: obs read + residual/Jacobian written
        return n_obs * (44 + 16 + 16 * (12 + K))
    if stage == "build_frames":
        if fused:  # obs read once; per (frame, camera): B 36 + E 6*(6+K) + g 6 + C (6+K)(7+K)/2 + gc (6+K) doubles
            NG = 6 + K
            return n_obs * 44 + n_groups * 8 * (36 + 6 * NG + 6 + NG * (NG + 1) // 2 + NG)
        return n_obs * (16 + 16 * (12 + K))
    return None


def fused_flops(n_obs, K):
    """Useful FP64 flops of the evaluate + build work per iteration: lower triangle of the Gram matrix of the
    reduced Jacobian [pose 6 | intrinsics K | residual] (2 rows per corner, 2 flop per entry) + ~300 flops of pose
    chain / projection / Jacobian per corner (213 FP64 instructions per corner in the ncu count, DFMA = 2).  The
    extrinsic blocks come from the constant map J_ck = J_pose A applied per (frame, camera), not per corner."""
    W = 7 + K
    return n_obs * (2 * (W * (W + 1) // 2) * 2 + 300)


def roofline(g, top, stages, top_bytes, achieved, peaks, peak_src, p, K0, fused, device, workload, persistent=False,
             step_s=None):
    """Roofline of the dominant kernel.  The evaluate + J^T J work is bound by the FP64 pipe (DMMA m8n8k4 + the
    projection chain), not by HBM (44 B per corner): the line is in TFLOP/s against the FP64 throughput measured
    live on this device (vcgpu_fp64_peak); the HBM view of the same launch rides along.

    Persistent engine: the dominant kernel is the whole iteration (lm_mega_kernel, one launch per iteration in the
    flushed timed region), so `achieved` = useful FP64 flops of an iteration / the launch's duration; the build
    phase alone (device clocks) is reported beside it."""
    hbm = {"achieved_gbs": achieved, "peak_gbs": peaks["hbm_gbs"], "frac": achieved / peaks["hbm_gbs"],
           "bytes_per_launch": top_bytes, "peak_source": peak_src}
    traffic = None
    kname = "lm_mega_kernel" if persistent else ("fused_build_kernel" if fused else top)
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "ncu_top_kernel.json")) as f:
            cap = json.load(f)
        if cap.get("workload") == workload and cap.get("kernel") == kname:
            traffic = cap.get("dram_bytes_per_launch")
    except Exception:
        pass
    if fused and top == "build_frames":
        dfma, dmma = g.fp64_peak(device)
        flops = fused_flops(p.n_obs, K0)
        peak = max(dfma, dmma)
        src = "FP64 DMMA/DFMA throughput measured live by vcgpu_fp64_peak (dfma %.1f, dmma %.1f TFLOP/s)" % (dfma, dmma)
        note = ("useful FP64 flops only (lower triangle of the 2-row outer products + projection chain); the DMMA "
                "tiles also compute the padded/upper parts")
        phase_tf = flops / (stages[top]["ms_per_iter"] * 1e-3) / 1e12
        if persistent:
            tf = flops / step_s / 1e12
            hbm["achieved_gbs"] = top_bytes / step_s / 1e9
            hbm["frac"] = hbm["achieved_gbs"] / peaks["hbm_gbs"]
            return {"bound": "tensor", "kernel": "lm_mega_kernel (whole LM iteration: solves, update, build, decision)",
                    "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": traffic,
                    "peak_source": src, "flops_per_launch": flops, "note": note,
                    "build_phase": {"achieved": phase_tf, "frac": phase_tf / peak, "ms": stages[top]["ms_per_iter"]},
                    "hbm": hbm}
        return {"bound": "tensor", "kernel": "fused_build_kernel (stage build_frames)", "achieved": phase_tf, "peak": peak,
                "unit": "TFLOP/s", "frac": phase_tf / peak, "traffic": traffic, "peak_source": src,
                "flops_per_launch": flops, "note": note, "hbm": hbm}
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": top_bytes}


def run_ours(args):
    from vicalib_b200.capi import Calibrator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")  # rendezvous only; the data-path collective is NCCL inside libvcgpu
    # weak scaling: every rank owns one BASELINE-config block of frames of a world x larger joint problem
    p = synth.make_config(args.workload, seed=20260924 + 100 * rank) if world > 1 else synth.make_config(args.workload)
    if world > 1 and rank > 0:  # globals (cameras) are shared: every shard uses rank 0's camera truth / guess
        p0 = synth.make_config(args.workload)
        p.models, p.intr, p.q_ck, p.p_ck = p0.models, p0.intr, p0.q_ck, p0.p_ck
    flags = dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1) if p.inertial else {}
    K0 = synth.NUM_INTR[int(p.models[0])]

    def new_cal():
        c = Calibrator(device=local)
        if world > 1:  # one NCCL unique id per communicator
            box = [Calibrator.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            c.comm_init(box[0], rank, world)
        return c

    g = new_cal()
    g.load(p)
    g.set_flags(**flags)
    g.set_options(max_iters=args.steps)
    # ---- warm-up (W untimed iterations; also builds all device buffers)
    g.set_profiling(False, not args.no_flush)
    g.iterate(max(args.warmup, 3))
    # ---- timed: exactly K iterations from the same initial guess
    g.load(p)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    s = g.iterate(args.steps)
    dev_s = s["device_seconds"]
    # extra timed repeats keep the clock sampler busy long enough to see the loaded clocks
    reps = [dev_s]
    t_end = time.time() + 1.0
    while time.time() < t_end:
        g.load(p)
        reps.append(g.iterate(args.steps)["device_seconds"])
    clocks = sampler.stop()
    dev_s = float(np.median(reps))
    launches = s["kernel_launches"]
    # ---- per-stage device time (separate profiled pass, same workload, L2 flushed the same way)
    # the single-GPU vision solve runs in the persistent kernel: its phases are clocked on the device
    # (%globaltimer, CTA 0); every other path is a sequence of launches bracketed by CUDA events
    persistent = not p.inertial and launches < 3 * args.steps
    g.load(p)
    g.set_profiling(8 if persistent else 1, not args.no_flush)
    g.iterate(args.steps)
    st = g.stage_times()
    g.set_profiling(False, False)
    stages = {k: {"ms_per_iter": v[0] / args.steps, "launches_per_iter": v[1] / args.steps} for k, v in st.items() if v[1]}
    # ---- no-flush number for information
    g.load(p)
    nf_s = g.iterate(args.steps)["device_seconds"]
    # ---- end to end through the C-ABI with host buffers
    e2e_t = []
    g2 = new_cal()  # device context / NCCL communicator creation is one-time setup, not part of a solve
    g2.load(p)
    g2.set_flags(**flags)
    g2.set_options(max_iters=args.steps)
    g2.iterate(2)
    for _ in range(5):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        g2.load(p)                 # host buffers -> sort by (camera, frame) -> H2D
        g2.iterate(args.steps)     # K iterations on the device
        st2 = g2.state()           # D2H of the solved parameters
        e2e_t.append(time.perf_counter() - t0)
    g2.close()
    e2e_s = float(np.median(e2e_t))
    h2d = (p.n_obs * (4 + 4 + 24 + 16) + p.n_frames * 88 + p.n_cams * (4 + 136) + len(p.imu_t) * 56 + 120)
    d2h = p.n_frames * 80 + p.n_cams * 136 + 120 + args.steps * 128
    # ---- max over ranks
    if world > 1:
        import torch

        t = torch.tensor([dev_s, e2e_s, nf_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, nf_s = t.tolist()
        dist.barrier()
    if rank != 0:
        return
    peaks, peak_src = _peaks()
    n_obs_total = p.n_obs * world
    n_groups = p.n_frames * p.n_cams
    fused = "eval_reproj" not in stages
    if persistent:
        stages = {("build_phase" if k == "build_frames" else k): v for k, v in stages.items()}
        stages["build_frames"] = stages["build_phase"]
    top = max((k for k in stages if stage_bytes(k, p.n_obs, K0, n_groups, fused)), key=lambda k: stages[k]["ms_per_iter"])
    top_bytes = stage_bytes(top, p.n_obs, K0, n_groups, fused)
    achieved = top_bytes / (stages[top]["ms_per_iter"] * 1e-3) / 1e9
    out = {
        # whole-job aggregate: each rank advances one BASELINE-config block per iteration (weak scaling)
        "metric": METRIC, "value": args.steps * world / dev_s, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: " + _describe(p), "n_obs": n_obs_total, "n_frames": p.n_frames,
                   "cameras": [int(m) for m in p.models], "l2": "flushed before every iteration" if not args.no_flush
                   else "not flushed", "value_no_flush": args.steps * world / nf_s,
                   "multi_gpu": None if world == 1 else f"{world} frame shards of {p.n_frames} frames each solved jointly; "
                   "2 reductions / iteration (reduced Schur system, global blocks + scalars), inside the persistent kernel through "
                   "NVLink peer stores (vision) or as NCCL all-reduces (inertial); value counts "
                   "block-iterations (N blocks per joint iteration)",
                   "algorithmic_bytes_per_obs_iter": algorithmic_bytes_per_obs(K0),
                   "iteration_hbm_frac": p.n_obs * algorithmic_bytes_per_obs(K0) / (dev_s / args.steps) / 1e9 / peaks["hbm_gbs"],
                   "engine": "persistent cooperative kernel (one launch per iteration in the flushed timed region, one per "
                   "solve otherwise)" if persistent else "multi-launch",
                   "stages_ms_per_iter": {k: round(v["ms_per_iter"], 5) for k, v in stages.items() if k != "build_phase"},
                   "accepted_steps": s["successful_steps"], "final_cost": s["final_cost"]},
        "clocks": clocks,
        "e2e": {"value": args.steps * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "note": "upload + K iterations + state read-back, wall clock"},
        "gpu_launches": launches,
        "roofline": roofline(g, top, stages, top_bytes, achieved, peaks, peak_src, p, K0, fused, local, args.workload,
                             persistent, dev_s / args.steps),
        "cpu_baseline": cpu_baseline(p, args) if world == 1 else None,  # timed on rank 0 at N=1 only
    }
    _emit(out)


def _describe(p):
    names = {v: k for k, v in synth.MODEL_IDS.items()}
    return (f"{p.n_cams} cam ({','.join(names[int(m)] for m in p.models)}), {p.n_frames} frames, "
            f"{p.n_obs // (p.n_frames * p.n_cams)} corners" + (" + IMU" if p.inertial else ", no IMU"))


def cpu_baseline(p, args, threads=None, iters=None):
    """The oracle (CPU port of the reference's Ceres path: one dual-number evaluation per residual
    block, block-sparse normal equations, block Cholesky, same trust-region loop) on the host."""
    from oracle.binding import Oracle

    cores = threads or min(os.cpu_count() or 1, 64)
    o = Oracle(p, inertial=int(p.inertial))
    iters = iters or 4
    o.set_options(max_iters=iters, function_tol=0.0, gradient_tol=0.0, param_tol=0.0, num_threads=cores)
    t0 = time.perf_counter()
    s = o.solve()
    dt = time.perf_counter() - t0
    n = max(int(s["iterations"]), 1)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} LM iterations of the full workload ({p.n_obs} observations), {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = synth.make_config(args.workload)
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    # warm-up + K timed steps, each a full LM iteration of the workload
    cpu_baseline(p, args, threads=threads, iters=max(1, min(args.warmup, 2)))
    cb = cpu_baseline(p, args, threads=threads, iters=args.steps)
    v = cb["value"]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.workload}: " + _describe(p),
                      "note": "Ceres/Calibu/Sophus/Eigen are not in the image: the reference arm is the CPU port "
                              "(oracle/) of the reference's Ceres path on the host cores",
                      "multi_gpu": None if int(os.environ.get("WORLD_SIZE", "1")) == 1 else
                      "the GPU arm counts block-iterations of an N-block joint problem; the CPU arm times one block: on the "
                      "same cores an N-block problem takes N times longer per iteration, i.e. the same block-iterations/s"},
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="config2", choices=list(synth.CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner) goes to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
