"""Target for ncu captures: warm-up launch, then one solve of K iterations (persistent kernel: one launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = synth.make_config(wl)
g = Calibrator(); g.load(p); g.set_options(max_iters=K); g.set_profiling(mode, False)
g.iterate(3)
g.load(p); s = g.iterate(K)
print(s)
