"""GPU: the DOGLEG strategy on the device (vc_dogleg.cuh; the reference's solver setting,
vicalibrator.h:151) against the CPU oracle's DoglegStrategy restatement on identical seeded inputs.

The per-iteration trace must match: same accept/reject sequence, cost after every iteration to 1e-9
relative, trust-region radius to 1e-6 relative (the dogleg point is a ratio of small inner products);
solved parameters to 1e-6 relative (north_star's tolerance).
"""
import numpy as np
import pytest

from vicalib_b200 import synth

pytestmark = pytest.mark.gpu


def _both(p, iters, inertial=False, **flags):
    from oracle.binding import Oracle
    from vicalib_b200.capi import Calibrator

    flags = dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1) if inertial else {}
    o = Oracle(p, **flags)
    o.set_options(max_iters=iters, function_tol=0.0, gradient_tol=0.0, param_tol=0.0, num_threads=8, strategy=1)
    so = o.solve()
    g = Calibrator()
    g.load(p)
    if inertial:
        g.set_flags(**flags)
    g.set_options(max_iters=iters, function_tol=0.0, gradient_tol=0.0, param_tol=0.0, strategy=1)
    sg = g.solve()
    return o, so, g, sg


@pytest.mark.parametrize("models,intr_init", [(("poly3",), "perturbed"), (("fov", "kb4"), "perturbed"), (("poly2",), "seed")])
def test_dogleg_trace_matches_oracle(models, intr_init):
    p = synth.make_problem(models=models, n_frames=30, grid=(14, 10), seed=13, intr_init=intr_init)
    o, so, g, sg = _both(p, 6)  # converged to rounding after ~7 iterations; compare the descent
    ro, rg = so["rows"], sg["rows"]
    assert len(ro) == len(rg)
    # columns: iteration, cost, cost_change, gmax, gnorm, step_norm, rho, radius, successful
    assert np.array_equal(ro[:, 8], rg[:, 8]), "accept/reject sequence differs"
    assert np.abs(rg[:, 1] - ro[:, 1]).max() <= 1e-9 * ro[:, 1].max()
    assert np.abs(rg[:, 7] - ro[:, 7]).max() <= 1e-6 * ro[:, 7].max()
    xo, xg = o.state(), g.state()
    for k in ("intr", "q_ck", "p_ck", "T_wp"):
        assert np.abs(xg[k] - xo[k]).max() <= 1e-6 * max(np.abs(xo[k]).max(), 1.0), k


def test_dogleg_converges_like_lm():
    from vicalib_b200.capi import Calibrator

    p = synth.make_problem(models=("poly3",), n_frames=60, grid=(14, 10), seed=4)
    out = {}
    for strategy in (0, 1):
        g = Calibrator()
        g.load(p)
        g.set_options(max_iters=60, strategy=strategy)
        s = g.solve()
        out[strategy] = (s, g.state())
    (s0, x0), (s1, x1) = out[0], out[1]
    assert s1["termination"] in (1, 2, 3, 4), s1  # converged, not out of iterations
    assert abs(s1["final_cost"] - s0["final_cost"]) <= 1e-5 * s0["final_cost"]
    assert np.abs(x0["intr"] - x1["intr"]).max() <= 1e-3 * np.abs(x0["intr"]).max()


def test_dogleg_inertial_matches_oracle():
    p = synth.make_problem(models=("poly3",), n_frames=24, grid=(14, 10), inertial=True, seed=17)
    o, so, g, sg = _both(p, 6, inertial=True)
    ro, rg = so["rows"], sg["rows"]
    assert len(ro) == len(rg)
    assert np.array_equal(ro[:, 8], rg[:, 8])
    assert np.abs(rg[:, 1] - ro[:, 1]).max() <= 1e-7 * ro[:, 1].max()
