"""solve_arrow through the persistent chain solver vs the multi-launch engine vs the oracle, for growing chain lengths"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
from oracle.binding import Oracle
models, flags = ("poly3", "poly3"), dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1)
for nf in (int(os.environ.get("NF", "300")),):
    p = synth.make_problem(models=models, n_frames=nf, inertial=True, seed=15)
    o = Oracle(p, **flags); g = Calibrator(); g.load(p); g.set_flags(**flags)
    ne = o.normal_equations()
    diag = np.concatenate([np.einsum("fii->fi", ne["B"]).ravel(), np.diag(ne["C"])])
    scale = 1.0 / (1.0 + np.sqrt(diag))
    D2 = np.clip(diag * scale * scale, 1e-6, 1e32) / 1e4
    xo = o.solve_arrow(scale, D2)
    for prof, name in ((0, "persistent"), (4, "multi-launch")):
        g.set_profiling(prof, 0)
        try:
            xg = g.solve_arrow(scale, D2)
            err = np.abs(xg - xo)
            fr = err[:nf * 9].reshape(nf, 9).max(1)
            print(nf, name, "rel err", err.max() / np.abs(xo).max(), "worst frames", np.argsort(fr)[-4:], "globals err", err[nf * 9:].max() / np.abs(xo).max(), flush=True)
        except Exception as e:
            print(nf, name, "ERR", e, flush=True)
