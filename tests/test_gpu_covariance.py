"""GPU: solution covariance of the global parameters (GetSolutionCovariance, vicalibrator.h:802-857) from the
device's frame elimination vs a dense numpy inverse of the oracle's J^T J on small problems."""
import numpy as np
import pytest

from vicalib_b200 import synth

pytestmark = pytest.mark.gpu


def _dense_cov(o):
    """[globals, globals] block of inv(J^T J) from the oracle's block normal equations, constant columns removed"""
    ne = o.normal_equations()
    nf, fd, G = ne["B"].shape[0], o.fd, o.G
    n = nf * fd + G
    H = np.zeros((n, n))
    for f in range(nf):
        s = slice(f * fd, (f + 1) * fd)
        H[s, s] = ne["B"][f]
        H[s, nf * fd:] = ne["E"][f]
        H[nf * fd:, s] = ne["E"][f].T
        if f > 0:
            sp = slice((f - 1) * fd, f * fd)
            H[sp, s] = ne["U"][f]
            H[s, sp] = ne["U"][f].T
    H[nf * fd:, nf * fd:] = ne["C"]
    mask = o.global_mask() != 0
    keep = np.concatenate([np.ones(nf * fd, bool), mask])
    Hi = np.linalg.inv(H[np.ix_(keep, keep)])
    cov = np.zeros((G, G))
    idx = np.where(mask)[0]
    cov[np.ix_(idx, idx)] = Hi[nf * fd:, nf * fd:]
    return cov


@pytest.mark.parametrize("models,inertial,flags", [
    (("poly3",), False, {}),
    (("fov", "kb4"), False, {}),
    (("poly3", "poly2"), True, dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1)),
    (("poly3",), True, dict(inertial=1, bias_active=1, scale_active=0, optimize_ts=0)),
])
def test_covariance_matches_dense_inverse(models, inertial, flags):
    from oracle.binding import Oracle
    from vicalib_b200.capi import Calibrator

    # inertial problems need a few seconds of motion before biases / gravity are observable (else J^T J is singular to
    # working precision and neither inverse means anything)
    nf = 90 if inertial else 18
    p = synth.make_problem(models=models, n_frames=nf, inertial=inertial, seed=15)
    g = Calibrator()
    g.load(p)
    g.set_flags(**flags)
    g.set_options(function_tol=1e-12, max_iters=15)
    g.solve()
    st = g.state()
    # the oracle evaluates J^T J at the state the device converged to
    p2 = synth.make_problem(models=models, n_frames=nf, inertial=inertial, seed=15)
    p2.intr, p2.q_ck, p2.p_ck, p2.T_wp, p2.v_w = st["intr"], st["q_ck"], st["p_ck"], st["T_wp"], st["v_w"]
    p2.g, p2.b, p2.sf, p2.ts = st["g"], st["b"], st["sf"], st["ts"]
    o = Oracle(p2, **flags)
    if inertial:
        o.set_imu_weights(g.imu_weights())
    cov_o = _dense_cov(o)
    cov_g = g.covariance()
    assert cov_g.shape == cov_o.shape
    sd = np.sqrt(np.maximum(np.diag(cov_o), 1e-300))
    corr_err = np.abs(cov_g - cov_o) / np.outer(sd, sd).clip(1e-300)
    live = np.diag(cov_o) > 0
    # every entry relative to the two standard deviations; both inverses lose cond(J^T J) * eps digits, and the inertial
    # systems are conditioned ~1e10 even after Jacobi scaling
    assert corr_err[np.ix_(live, live)].max() <= (1e-4 if inertial else 1e-6)
    assert np.array_equal(cov_g[~live], np.zeros_like(cov_g[~live]))  # constant blocks: zero rows (ceres::Covariance)
    assert np.abs(cov_g - cov_g.T).max() <= 1e-9 * np.abs(cov_g).max()
    # a second solve after the covariance call is unaffected by it
    c0 = g.cost()
    assert abs(c0 - o.cost()) <= 1e-9 * c0
