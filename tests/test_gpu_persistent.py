"""GPU: the persistent cooperative kernel (vc_mega.cuh) against the multi-launch engine on the same
inputs, and against the CPU oracle.  Both device paths run the same algorithm with different summation
orders, so they agree to rounding:  cost 1e-10 relative, parameters 1e-8 relative after a fixed number
of iterations (the trust-region decisions must be identical: same accepted-step count).

Shapes are chosen to reach every branch of the persistent kernel: several cameras with different
models, cameras that miss frames, more frames per CTA than warps (frame loop inside a warp), and more
(frame, camera) pairs per CTA than the static group table holds (global look-up fallback).
"""
import numpy as np
import pytest

from vicalib_b200 import synth

pytestmark = pytest.mark.gpu

MULTI_LAUNCH = 4  # vcgpu_set_profiling bit 2


def _run(p, mode, iters):
    from vicalib_b200.capi import Calibrator

    g = Calibrator()
    g.load(p)
    g.set_options(max_iters=iters, function_tol=0.0, gradient_tol=0.0, param_tol=0.0)
    g.set_profiling(mode, False)
    s = g.solve()
    st = g.state()
    ne = g.normal_equations()
    return s, st, ne


def _agree(p, iters=6):
    s0, st0, ne0 = _run(p, 0, iters)
    s1, st1, ne1 = _run(p, MULTI_LAUNCH, iters)
    assert s0["kernel_launches"] < s1["kernel_launches"], "the persistent kernel did not run"
    assert s0["iterations"] == s1["iterations"] == iters
    assert s0["successful_steps"] == s1["successful_steps"]
    assert abs(s0["final_cost"] - s1["final_cost"]) <= 1e-10 * s1["final_cost"]
    for k in ("intr", "q_ck", "p_ck", "T_wp"):
        assert np.abs(st0[k] - st1[k]).max() <= 1e-8 * max(np.abs(st1[k]).max(), 1.0), k
    # the blocks the persistent kernel hands back (accepted point) are those of the multi-launch engine
    for k in ("B", "E", "C"):
        scale = max(np.abs(ne1[k]).max(), 1e-300)
        assert np.abs(ne0[k] - ne1[k]).max() <= 1e-7 * scale, k
    # gradients vanish at the optimum: they differ by H * (parameter difference), so the Hessian sets the scale
    assert np.abs(ne0["gf"] - ne1["gf"]).max() <= 1e-8 * np.abs(ne1["B"]).max()
    assert np.abs(ne0["gc"] - ne1["gc"]).max() <= 1e-8 * np.abs(ne1["C"]).max()
    return s0


@pytest.mark.parametrize("models", [("poly3",), ("fov", "kb4"), ("poly3", "poly2", "linear"),
                                    ("poly3", "poly3", "kb4", "fov")])  # 4 cameras: G = 51, fewer warps fit
def test_persistent_matches_multi_launch(models):
    p = synth.make_problem(models=models, n_frames=200 if len(models) < 4 else 80, grid=(14, 10), seed=11)
    _agree(p)


def test_cameras_missing_frames():
    p = synth.make_problem(models=("poly3", "fov"), n_frames=120, grid=(14, 10), seed=5)
    keep = ~((p.obs_cam == 1) & (p.obs_frame % 3 == 0)) & ~((p.obs_cam == 0) & (p.obs_frame % 7 == 2))
    keep &= ~(p.obs_frame == 50)  # and one frame nobody sees
    p.obs_frame, p.obs_cam, p.p_w, p.p_c = p.obs_frame[keep], p.obs_cam[keep], p.p_w[keep], p.p_c[keep]
    if hasattr(p, "grid_idx") and p.grid_idx is not None:
        p.grid_idx = p.grid_idx[keep]
    _agree(p)


def test_more_frames_than_warps():
    # 148 CTAs x 16 warps = 2368 frame slots: 6000 frames make every warp loop over 2-3 frames
    p = synth.make_problem(models=("fov",), n_frames=6000, grid=(4, 3), seed=3)
    _agree(p, iters=4)


def test_group_table_fallback():
    # > 256 (frame, camera) pairs per CTA: groups are looked up in global memory
    p = synth.make_problem(models=("linear",), n_frames=40000, grid=(3, 2), seed=9)
    _agree(p, iters=3)


def test_persistent_matches_oracle():
    from oracle.binding import Oracle

    p = synth.make_problem(models=("poly3", "fov"), n_frames=40, grid=(14, 10), seed=21)
    o = Oracle(p)
    o.set_options(max_iters=8, function_tol=0.0, gradient_tol=0.0, param_tol=0.0, num_threads=8)
    so = o.solve()
    sg, st, _ = _run(p, 0, 8)
    assert sg["successful_steps"] == so["successful_steps"]
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-9 * so["final_cost"]
    xo = o.state()
    for k in ("intr", "q_ck", "p_ck", "T_wp"):
        assert np.abs(st[k] - xo[k]).max() <= 1e-6 * max(np.abs(xo[k]).max(), 1.0), k


def test_callback_mode_uses_single_iteration_launches():
    from vicalib_b200.capi import Calibrator

    p = synth.make_problem(models=("poly3",), n_frames=60, grid=(14, 10), seed=2)
    g = Calibrator()
    g.load(p)
    g.set_options(max_iters=5, function_tol=0.0, gradient_tol=0.0, param_tol=0.0)
    seen = []
    s = g.solve(callback=lambda it: seen.append((it.iteration, it.cost)) or 0)
    assert [k for k, _ in seen] == list(range(0, 6))
    costs = [c for _, c in seen]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(costs, costs[1:]))
    assert abs(costs[-1] - s["final_cost"]) <= 1e-12 * s["final_cost"]
