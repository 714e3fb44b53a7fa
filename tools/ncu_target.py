"""Target for ncu captures: warm-up iterations, then one solve of K iterations (persistent kernels: one or two launches
per iteration).   python tools/ncu_target.py [workload] [K] [profile-mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator

wl = sys.argv[1] if len(sys.argv) > 1 else "target"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = synth.make_config(wl)
g = Calibrator()
g.load(p)
if p.inertial:
    g.set_flags(inertial=1, bias_active=1, scale_active=1, optimize_ts=1)
g.set_options(max_iters=K)
g.set_profiling(mode, False)
g.iterate(3)
g.load(p)
s = g.iterate(K)
print(s)
