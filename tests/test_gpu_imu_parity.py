"""GPU parity: inertial path (IMU residual/Jacobian by lane-parallel dual numbers, IMU blocks of the
normal equations, block-tridiagonal chain solve, LM solve, UpdateImuWeights) vs the CPU oracle.

Tolerances (FP64): IMU residuals are O(1e2) after the 500*I initial weight; Jacobian entries reach
O(1e5) (weight x lever arms), so comparisons are relative to the largest entry.
"""
import numpy as np
import pytest

from vicalib_b200 import synth

pytestmark = pytest.mark.gpu

ALL_ON = dict(inertial=1, rotation_only=0, bias_active=1, scale_active=1, optimize_ts=1)
STAGES = [
    dict(inertial=1, rotation_only=1, bias_active=0, scale_active=0, optimize_ts=1),  # stage 2 of SolveThread
    dict(inertial=1, rotation_only=0, bias_active=1, scale_active=0, optimize_ts=1),  # stage 3
    ALL_ON,                                                                           # stage 4 / has_initial_guess
    dict(inertial=1, rotation_only=0, bias_active=1, scale_active=1, optimize_ts=0),
]


def _pair(models=("poly3",), n_frames=14, flags=ALL_ON, seed=21, **kw):
    from oracle.binding import Oracle
    from vicalib_b200.capi import Calibrator

    p = synth.make_problem(models=models, n_frames=n_frames, grid=(14, 10), inertial=True, seed=seed, **kw)
    o = Oracle(p, **flags)
    g = Calibrator()
    g.load(p)
    g.set_flags(**flags)
    return p, o, g


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("flags", STAGES)
def test_imu_residuals_and_jacobians(flags):
    p, o, g = _pair(flags=flags, ts_truth=0.003)
    r_o, J_o = o.eval_imu()
    r_g, J_g = g.eval_imu()
    assert r_o.shape == (p.n_frames - 1, 9)
    assert np.abs(r_g - r_o).max() <= 1e-9 * max(np.abs(r_o).max(), 1.0)
    assert np.abs(J_g - J_o).max() <= 1e-9 * np.abs(J_o).max()


def test_imu_time_offset_boundary_crossing():
    """d/d(ts) across sample-boundary crossings: shift ts so interval endpoints straddle samples."""
    for ts in (-0.0049, 0.0, 0.00251, 0.0074):
        p, o, g = _pair(n_frames=8, seed=5)
        p.ts = ts
        o.set_imu_params(p.g, p.b, p.sf, ts)
        g.set_imu_params(p.g, p.b, p.sf, ts)
        r_o, J_o = o.eval_imu()
        r_g, J_g = g.eval_imu()
        assert np.abs(r_g - r_o).max() <= 1e-9 * max(np.abs(r_o).max(), 1.0)
        assert np.abs(J_g[:, :, 32] - J_o[:, :, 32]).max() <= 1e-9 * np.abs(J_o).max()


@pytest.mark.parametrize("flags", STAGES)
def test_normal_equations_with_imu(flags):
    p, o, g = _pair(models=("poly3", "fov"), flags=flags)
    ne_o, ne_g = o.normal_equations(), g.normal_equations()
    assert o.fd == g.fd == 9 and o.G == g.G
    assert abs(ne_g["cost"] - ne_o["cost"]) <= 1e-11 * ne_o["cost"]
    for k in ("B", "U", "E", "gf", "C", "gc"):
        assert _relerr(ne_g[k], ne_o[k]) <= 1e-10, k


@pytest.mark.parametrize("n_frames", [3, 4, 5, 9, 17, 70])
def test_chain_solve_matches_oracle(n_frames):
    """Block-tridiagonal + arrow solve (partitioned elimination on the device) vs the oracle's
    sequential block Cholesky, for chain lengths around the chunk / level boundaries."""
    p, o, g = _pair(models=("poly2",), n_frames=n_frames)
    ne = o.normal_equations()
    diag = np.concatenate([np.einsum("fii->fi", ne["B"]).ravel(), np.diag(ne["C"])])
    scale = 1.0 / (1.0 + np.sqrt(diag))
    D2 = np.clip(diag * scale * scale, 1e-6, 1e32) / 1e4
    x_o = o.solve_arrow(scale, D2)
    x_g = g.solve_arrow(scale, D2)
    assert _relerr(x_g, x_o) <= 1e-7


@pytest.mark.parametrize("models,flags", [(("poly3",), ALL_ON), (("fov", "fov"), STAGES[0]), (("kb4", "poly3"), STAGES[1])])
def test_lm_solve_with_imu_fixed_weights(models, flags):
    p, o, g = _pair(models=models, n_frames=40, flags=flags)
    o.set_options(function_tol=1e-13, max_iters=40, update_imu_weights=0)
    g.set_options(function_tol=1e-13, max_iters=40, update_imu_weights=0)
    s_o, s_g = o.solve(), g.solve()
    assert abs(s_g["final_cost"] - s_o["final_cost"]) <= 1e-8 * s_o["final_cost"]
    st_o, st_g = o.state(), g.state()
    for c, m in enumerate(p.models):
        K = synth.NUM_INTR[int(m)]
        rel = np.abs(st_g["intr"][c, :K] - st_o["intr"][c, :K]) / np.maximum(np.abs(st_o["intr"][c, :K]), 1e-3)
        assert rel.max() <= 1e-6
    for k in ("T_wp", "v_w", "q_ck", "p_ck", "g", "b", "sf"):
        assert np.abs(st_g[k] - st_o[k]).max() <= 1e-6 * max(1.0, np.abs(st_o[k]).max()), k
    assert abs(st_g["ts"] - st_o["ts"]) <= 1e-8


def test_update_imu_weights_matches_oracle():
    """UpdateImuWeights KAT: per-interval 9x9 weight_sqrt_ for fixed inputs (vicalibrator.h:723-799)."""
    p, o, g = _pair(models=("poly3",), n_frames=25, ts_truth=0.003)
    o.update_imu_weights()
    g.update_imu_weights()
    W_o, W_g = o.imu_weights(), g.imu_weights()
    assert np.abs(W_o - 500 * np.eye(9)).max() > 1.0  # it did change
    for k in range(W_o.shape[0]):
        assert np.abs(W_g[k] - W_o[k]).max() <= 1e-7 * np.abs(W_o[k]).max(), k
    # rotation-only: weights must stay untouched (vicalibrator.h:725)
    p, o, g = _pair(flags=STAGES[0])
    g.update_imu_weights()
    assert np.array_equal(g.imu_weights(), np.broadcast_to(500 * np.eye(9), (p.n_frames - 1, 9, 9)))


def test_lm_solve_with_weight_updates():
    p, o, g = _pair(models=("poly3", "poly3"), n_frames=40)
    o.set_options(function_tol=1e-13, max_iters=25, update_imu_weights=1)
    g.set_options(function_tol=1e-13, max_iters=25, update_imu_weights=1)
    s_o, s_g = o.solve(), g.solve()
    assert abs(s_g["final_cost"] - s_o["final_cost"]) <= 1e-6 * s_o["final_cost"]
    st_o, st_g = o.state(), g.state()
    for k in ("T_wp", "q_ck", "p_ck", "intr"):
        assert np.abs(st_g[k] - st_o[k]).max() <= 1e-6 * max(1.0, np.abs(st_o[k]).max()), k
    # the weights W = sqrtm((Jt C Jt^T)^-1) agree to ~1e-15 (previous test), but 25 unconverged iterations of
    # this short, weakly observable trajectory amplify rounding-level differences into the accelerometer
    # bias / gravity directions: those are compared at 2e-5 absolute
    for k in ("v_w", "g", "b", "sf"):
        assert np.abs(st_g[k] - st_o[k]).max() <= 2e-5, k
