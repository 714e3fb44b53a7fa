"""End-to-end timing of one solve through the C-ABI from host buffers, split into the set_* calls, the K
iterations — through vcgpu_iterate (no callback) and through vcgpu_solve with a per-iteration callback (the
drop-in path of host/vicalibrator.h) — and the state read-back.   python tools/e2e_probe.py [workload] [K]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("E2E_TORCH"):  # bench.py runs with torch imported (torch.distributed plumbing): same timings?
    import torch  # noqa: F401
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator

wl = sys.argv[1] if len(sys.argv) > 1 else "target"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = synth.make_config(wl)
flags = dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1) if p.inertial else {}
g = Calibrator()
g.load(p)
g.set_flags(**flags)
g.set_options(max_iters=K, function_tol=0.0, gradient_tol=0.0, param_tol=0.0)
g.iterate(3)
if os.environ.get("E2E_FRESH"):  # bench.py times a second handle while the first is still alive
    g0 = g
    g = Calibrator()
    g.load(p)
    g.set_flags(**flags)
    g.set_options(max_iters=K, function_tol=0.0, gradient_tol=0.0, param_tol=0.0)
    g.solve()
for mode in ("iterate", "solve(cb)"):
    best = None
    for rep in range(6):
        t0 = time.perf_counter()
        g.set_cameras(p.models, p.intr, p.q_ck, p.p_ck)
        g.set_frames(p.T_wp, p.v_w, p.ftime)
        t1 = time.perf_counter()
        g.set_observations(p.obs_frame, p.obs_cam, p.p_w, p.p_c)
        g.set_imu(p.imu_t, p.imu_w, p.imu_a, synth.GYRO_SIGMA, synth.ACCEL_SIGMA)
        g.set_imu_params(p.g, p.b, p.sf, p.ts)
        t2 = time.perf_counter()
        s = g.iterate(K) if mode == "iterate" else g.solve(callback=lambda it: 0)
        t3 = time.perf_counter()
        g.state()
        t4 = time.perf_counter()
        row = (t1 - t0, t2 - t1, t3 - t2, s["device_seconds"], t4 - t3, t4 - t0, s["iterations"], s["kernel_launches"])
        if best is None or row[5] < best[5]:
            best = row
    print("%-9s best of 6: set_cam/frames %.2f ms  set_obs/imu %.2f  %d iterations %.2f (device %.2f, %d launches)  state %.2f  "
          "total %.2f ms -> %.0f it/s" % (mode, 1e3 * best[0], 1e3 * best[1], best[6], 1e3 * best[2], 1e3 * best[3], best[7],
                                          1e3 * best[4], 1e3 * best[5], best[6] / best[5]), flush=True)
