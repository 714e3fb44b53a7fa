"""End-to-end timing of one solve through the C-ABI from host buffers (upload / iterate / read-back)."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
p = synth.make_config("config2")
K = 20
g = Calibrator(); g.load(p); g.set_options(max_iters=K); g.iterate(3)
best = None
for rep in range(6):
    t0 = time.perf_counter(); g.set_cameras(p.models, p.intr, p.q_ck, p.p_ck); g.set_frames(p.T_wp, p.v_w, p.ftime); t1 = time.perf_counter()
    g.set_observations(p.obs_frame, p.obs_cam, p.p_w, p.p_c); t2 = time.perf_counter()
    s = g.iterate(K); t3 = time.perf_counter(); g.state(); t4 = time.perf_counter()
    row = (t1 - t0, t2 - t1, t3 - t2, s["device_seconds"], t4 - t3, t4 - t0)
    if best is None or row[-1] < best[-1]: best = row
print("best of 6: set_cam/frames %.2f ms  set_obs %.2f  iterate(%d) %.2f (device %.2f)  state %.2f  total %.2f ms -> %.0f it/s" %
      (1e3 * best[0], 1e3 * best[1], K, 1e3 * best[2], 1e3 * best[3], 1e3 * best[4], 1e3 * best[5], K / best[5]))
