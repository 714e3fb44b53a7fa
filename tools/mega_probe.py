"""Per-phase device clocks of the persistent vision kernel (profile bit 8) and A/B against the multi-launch engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
K = 20
p = synth.make_config(wl)
g = Calibrator(); g.load(p); g.set_options(max_iters=K); g.iterate(3)
for mode, name in ((0, "persistent"), (4, "multi-launch"), (8, "persistent+clocks")):
    g.set_profiling(mode, False)
    best = 1e9
    for _ in range(5):
        g.load(p); s = g.iterate(K); best = min(best, s["device_seconds"])
    print(f"{name:20s} {1e6*best/K:8.2f} us/iter  launches {s['kernel_launches']}")
    if mode == 8:
        st = g.stage_times()
        for k, v in st.items():
            if v[1]: print(f"    {k:14s} {1e3*v[0]/v[1]:8.2f} us/iter")
g.set_profiling(0, False)
for rep in range(3):
    t0=time.perf_counter(); g.load(p); t1=time.perf_counter(); s=g.iterate(K); t2=time.perf_counter(); g.state(); t3=time.perf_counter()
    print(f"e2e: load {1e3*(t1-t0):.2f} ms  iterate({K}) {1e3*(t2-t1):.2f} ms (device {1e3*s['device_seconds']:.2f})  state {1e3*(t3-t2):.2f} ms")
