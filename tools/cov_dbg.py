import sys; sys.path.insert(0,'.')
import numpy as np
from vicalib_b200 import synth
from vicalib_b200.capi import Calibrator
from oracle.binding import Oracle
models, flags = ("poly3","poly2"), dict(inertial=1,bias_active=1,scale_active=1,optimize_ts=1)
for nf in (18, 90):
    p = synth.make_problem(models=models, n_frames=nf, inertial=True, seed=15)
    o = Oracle(p, **flags); g = Calibrator(); g.load(p); g.set_flags(**flags)
    ne = o.normal_equations(); fd, G = o.fd, o.G
    diag = np.concatenate([np.einsum("fii->fi", ne["B"]).ravel(), np.diag(ne["C"])])
    scale = 1.0/(1.0+np.sqrt(diag))
    for dval in (1e-4, 1e-8, 1e-12, 0.0):
        D2 = np.full_like(diag, dval)
        try:
            xo = o.solve_arrow(scale, D2)
        except Exception as e:
            print(nf, dval, "oracle ERR", e); continue
        for prof, name in ((0, "persistent"), (4, "multi-launch")):
            g.set_profiling(prof, 0)
            try:
                xg = g.solve_arrow(scale, D2)
                print(nf, dval, name, "rel err", np.abs(xg-xo).max()/np.abs(xo).max())
            except Exception as e:
                print(nf, dval, name, "ERR", e)
