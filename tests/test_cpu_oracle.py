"""Oracle self-consistency on CPU: Lie helpers, interpolation-buffer edge cases, block solver vs a
dense numpy solve, LM / dogleg reach the same optimum and recover the synthetic truth
(ViSimTest-style assertions, testing/vi_sim_test.cpp:80-92), outlier removal, IMU weights."""
import copy

import numpy as np
import pytest

from oracle import binding as ob
from oracle.binding import Oracle
from vicalib_b200 import synth

ALL_ON = dict(inertial=1, rotation_only=0, bias_active=1, scale_active=1, optimize_ts=1)


def test_se3_exp_log_roundtrip_and_small_angle_branches():
    rng = np.random.default_rng(0)
    for scale in (1.0, 1e-3, 1e-9, 0.0):
        d = np.concatenate([rng.normal(size=3), scale * rng.normal(size=3)])
        x = ob.se3_exp(d)
        assert abs(np.linalg.norm(x[:4]) - 1) < 1e-14
        # theta ~ 1e-9 sits just above Sophus' 1e-10 Taylor switch, where (1-cos theta)/theta^2 has
        # already lost its digits in double precision: a property of the restated formulas, kept as is
        tol = 1e-12 if scale in (1.0, 0.0) else 3e-9
        assert np.abs(ob.se3_log(x) - d).max() < tol * max(1.0, np.abs(d).max())
    x = ob.se3_exp(np.array([0.1, -0.2, 0.3, 0.4, 0.5, -0.6]))
    assert np.allclose(ob.se3_plus(x, np.zeros(6)), x, atol=1e-15)


def test_projection_models_against_numpy_generator():
    rng = np.random.default_rng(1)
    for name, m in synth.MODEL_IDS.items():
        k = np.zeros(10)
        k[:4] = [310, 305, 322, 238]
        k[4:4 + len(synth.TRUTH_DIST[m])] = synth.TRUTH_DIST[m]
        for _ in range(20):
            ray = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.4, 1.0)])
            assert np.allclose(ob.project(m, ray, k), synth.project(m, ray, k), rtol=1e-13, atol=1e-10)


def test_fov_small_radius_branch_and_linear_limit():
    k = np.array([300, 300, 320, 240, 0.9, 0, 0, 0, 0, 0.0])
    near = ob.project(synth.FOV, np.array([1e-4, -1e-4, 1.0]), k)  # rad^2 < 1e-5 -> factor 2tan(w/2)/w
    assert np.allclose(near - k[2:4], 300 * 2 * np.tan(0.45) / 0.9 * np.array([1e-4, -1e-4]), rtol=1e-12)
    k[4] = 1e-3  # w^2 <= 1e-5 -> factor 1
    assert np.allclose(ob.project(synth.FOV, np.array([0.1, 0.2, 1.0]), k), [350, 300])


def test_imu_range_edge_cases():
    p = synth.make_problem(models=("linear",), n_frames=6, inertial=True, seed=2)
    o = Oracle(p, **ALL_ON)
    t0, t1 = p.ftime[1], p.ftime[2]
    m = o.imu_get_range(t0, t1, 0.0)
    assert m[0, 0] == t0 and m[-1, 0] == t1 and np.all(np.diff(m[:, 0]) > 0)
    inner = p.imu_t[(p.imu_t > t0) & (p.imu_t <= t1)]
    assert np.allclose(m[1:-1, 0], inner)
    # shifted by a time offset: the same samples appear with shifted stamps
    ms = o.imu_get_range(t0, t1, 0.004)
    inner = p.imu_t[(p.imu_t + 0.004 > t0) & (p.imu_t + 0.004 <= t1)]
    assert np.allclose(ms[1:-1, 0], inner + 0.004)
    # before the first sample: HasElement fails -> empty range -> zero residual (ceres-cost-functions.h:452-455)
    assert len(o.imu_get_range(p.imu_t[0] - 1.0, p.imu_t[0] - 0.5, 0.0)) == 0


@pytest.mark.parametrize("inertial", [False, True])
def test_block_solver_matches_dense(inertial):
    p = synth.make_problem(models=("poly2", "fov"), n_frames=9, inertial=inertial, seed=5)
    o = Oracle(p, **(ALL_ON if inertial else {}))
    ne = o.normal_equations()
    nf, fd, G = p.n_frames, o.fd, o.G
    n = nf * fd + G
    H = np.zeros((n, n))
    for f in range(nf):
        s = slice(f * fd, (f + 1) * fd)
        H[s, s] = ne["B"][f]
        H[s, nf * fd:] = ne["E"][f]
        H[nf * fd:, s] = ne["E"][f].T
        if f > 0:
            H[(f - 1) * fd:f * fd, s] = ne["U"][f]
            H[s, (f - 1) * fd:f * fd] = ne["U"][f].T
    H[nf * fd:, nf * fd:] = ne["C"]
    g = np.concatenate([ne["gf"].ravel(), ne["gc"]])
    assert np.allclose(H, H.T, rtol=1e-12, atol=1e-9)
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    D2 = np.clip(np.diag(H) * scale ** 2, 1e-6, 1e32) / 1e4
    x = o.solve_arrow(scale, D2)
    Hs = H * scale[:, None] * scale[None, :] + np.diag(D2)
    x_ref = np.linalg.solve(Hs, -g * scale)
    assert np.abs(x - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_lm_recovers_truth_and_dogleg_agrees():
    p = synth.make_config("config1", intr_init="seed")
    o = Oracle(p)
    o.set_options(function_tol=1e-12, num_threads=4)
    s = o.solve()
    st = o.state()
    assert s["termination"] in (1, 2, 3, 4)
    assert np.linalg.norm(st["intr"][0, :4] - p.truth["intr"][0, :4]) < 5.0      # vi_sim_test.cpp:87
    assert np.sqrt(2 * o.evaluate_camera(0) / p.n_obs) < 0.1 * np.sqrt(2) + 0.02   # rmse ~ pixel sigma
    o2 = Oracle(p)
    o2.set_options(function_tol=1e-12, num_threads=4, strategy=1)                 # DOGLEG (vicalibrator.h:151)
    s2 = o2.solve()
    assert abs(s2["final_cost"] - s["final_cost"]) <= 1e-7 * s["final_cost"]
    assert np.abs(o2.state()["intr"][0, :7] - st["intr"][0, :7]).max() <= 1e-5 * np.abs(st["intr"][0, :7]).max() + 1e-7


def test_masks_follow_setup_problem():
    p = synth.make_problem(models=("poly3", "fov"), n_frames=4, inertial=True, seed=3)
    # vision only: cam0 extrinsics constant, cam1 free, intrinsics free (vicalibrator.h:572-592)
    m = Oracle(p).global_mask()
    assert list(m[:6]) == [0] * 6 and np.all(m[6:13] == 1) and np.all(m[13:] == 1)
    # rotation-only IMU stage: cam0 rotation free, translation constant, gravity/bias/scale constant, ts free
    o = Oracle(p, inertial=1, rotation_only=1, optimize_ts=1)
    m = o.global_mask()
    io = 13 + 11
    assert list(m[:6]) == [1, 1, 1, 0, 0, 0] and list(m[io:io + 2]) == [0, 0] and m[io + 14] == 1 and np.all(m[io + 2:io + 14] == 0)
    assert o.fd == 9 and o.G == io + 15
    m = Oracle(p, fix_intrinsics=1).global_mask()
    assert np.all(m[6:13] == 0) and np.all(m[19:24] == 0)


def test_outlier_removal_and_multiplicity():
    p = synth.make_problem(models=("poly3",), n_frames=6, seed=8, intr_init="truth", pose_noise=(1e-6, 1e-6))
    bad = [5, 77, 300]
    p.p_c[bad] += 15.0
    o = Oracle(p)
    rmse = np.array([np.sqrt(o.evaluate_camera(0) / p.n_obs)])
    assert o.remove_outliers(rmse, 2.0) >= 3
    assert np.all(o.obs_active()[bad] == 0)
    c1 = o.cost()
    o.set_flags(visual_mult=3.0)  # staged flow re-adds the visual blocks
    assert abs(o.cost() - 3 * c1) <= 1e-12 * c1
    assert o.num_residuals() == 3 * 2 * int(o.obs_active().sum())


def test_imu_weights_are_sqrt_information():
    p = synth.make_problem(models=("linear",), n_frames=8, inertial=True, seed=12)
    o = Oracle(p, **ALL_ON)
    o.update_imu_weights()
    W = o.imu_weights()
    assert np.abs(W - np.swapaxes(W, 1, 2)).max() <= 1e-6 * np.abs(W).max()   # principal root of an SPD matrix
    ev = np.linalg.eigvalsh(0.5 * (W[0] + W[0].T))
    assert ev.min() > 0
    # rotation-only: untouched (vicalibrator.h:725)
    o2 = Oracle(p, inertial=1, rotation_only=1)
    o2.update_imu_weights()
    assert np.array_equal(o2.imu_weights()[0], 500 * np.eye(9))


def test_extrinsic_columns_are_a_constant_map_of_the_pose_columns():
    """The identity behind the reduced Gram formulation of the CUDA build phase: for one camera the six
    T_ck columns of the reprojection Jacobian are  J_pose * A  with  A = [[0, -R_ck^T], [-I, 0]]  for every corner,
    because both perturbations are right perturbations acting on the same point.  Checked on the oracle's dual-number
    Jacobians (independent of the kernels), all five models."""
    for models in (("linear",), ("fov",), ("poly2",), ("poly3",), ("kb4",)):
        p = synth.make_problem(models=models, n_frames=5, grid=(14, 10), seed=8)
        o = Oracle(p)
        _, J = o.eval_reproj()          # [n_obs, 2, 12 + K]: pose (t, w), extrinsics (w_ck, p_ck), intrinsics
        R = synth.quat_to_mat(p.q_ck[0])
        A = np.zeros((6, 6))
        A[0:3, 3:6] = -R.T
        A[3:6, 0:3] = -np.eye(3)
        Jp, Je = J[:, :, 0:6], J[:, :, 6:12]
        assert np.abs(Jp @ A - Je).max() <= 1e-12 * np.abs(J).max(), models
