import time, sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
p = synth.make_config("config2")
g = Calibrator(); g.load(p); g.set_options(max_iters=20); g.iterate(3)
for rep in range(3):
    t0=time.perf_counter(); g.set_cameras(p.models,p.intr,p.q_ck,p.p_ck); g.set_frames(p.T_wp,p.v_w,p.ftime); t1=time.perf_counter()
    g.set_observations(p.obs_frame,p.obs_cam,p.p_w,p.p_c); t2=time.perf_counter()
    g.set_imu(p.imu_t,p.imu_w,p.imu_a,1e-4,1e-3); g.set_imu_params(p.g,p.b,p.sf,p.ts); t3=time.perf_counter()
    s=g.iterate(20); t4=time.perf_counter(); st=g.state(); t5=time.perf_counter()
    print(f"set_cam/frames {1e3*(t1-t0):.2f} ms  set_obs {1e3*(t2-t1):.2f}  set_imu {1e3*(t3-t2):.2f}  iterate(20) {1e3*(t4-t3):.2f} (device {1e3*s['device_seconds']:.2f})  state {1e3*(t5-t4):.2f}")
