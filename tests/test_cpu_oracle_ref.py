"""CPU: the oracle restatement against oracle/_ref — the reference's OWN headers (types.h, vicalibrator-utils.h,
interpolation-buffer.h, ceres-cost-functions.h, local-param-se3.h) compiled unmodified from /root/reference/include
against stand-ins for Eigen / Sophus / ceres::Jet / glog (oracle/ref_shim, recipe: `make -C oracle _ref`).

This pins to reference TEXT: the IMU cost functor incl. its RK4 integrator and the interpolation buffer,
UpdateImuWeights' double integrator with covariance and the hand-derived derivative tables (a9), the SE3 / SO3 local
parameterisations (a4) and the pose chain of the reprojection functor (a1).  Still recalled, not pinned: Sophus'
exp / log (stand-in written from the published formulas), Calibu's Project bodies (a2) and the Ceres loop (a12).

oracle/_ref is built where /root/reference exists (this container) and travels to the GPU box as a binary."""
import ctypes as C
import os

import numpy as np
import pytest

from vicalib_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libvicalib_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref is built only where /root/reference exists")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(SO)
    for f in ("ref_get_range", "ref_imu_eval", "ref_update_weight", "ref_integrate", "ref_reproj"):
        getattr(L, f).restype = C.c_int
    return L


def _problem(**kw):
    args = dict(models=("poly3",), n_frames=12, grid=(14, 10), inertial=True, seed=77, ts_truth=0.003)
    args.update(kw)
    return synth.make_problem(**args)


def _imu(p):
    return C.c_int(len(p.imu_t)), _p(_c(p.imu_t)), _p(_c(p.imu_w)), _p(_c(p.imu_a))


def test_interpolation_buffer_get_range(ref):
    """InterpolationBufferT::GetRange (interpolation-buffer.h:208-226) incl. the edges: intervals that start before the
    first sample / end after the last one, time offsets that move interval ends across samples."""
    from oracle.binding import Oracle

    p = _problem()
    o = Oracle(p, inertial=1)
    t0s = [p.ftime[0], p.ftime[3], p.imu_t[0] - 0.01, p.imu_t[0], p.imu_t[5], p.imu_t[-1] - 0.004, p.imu_t[-1] + 0.01]
    for ts in (-0.0049, 0.0, 0.00251, 0.0074):
        for t0 in t0s:
            for dt in (1e-4, 1.0 / 30, 0.2):
                out = np.zeros((256, 7))
                n = ref.ref_get_range(*_imu(p), C.c_double(t0), C.c_double(t0 + dt), C.c_double(ts), _p(out), 256)
                mine = o.imu_get_range(t0, t0 + dt, ts)
                assert n == mine.shape[0]
                assert np.allclose(out[:n], mine, rtol=1e-14, atol=1e-15), (ts, t0, dt)  # same samples; FMA contraction differs


@pytest.mark.parametrize("rot_only", [0, 1])
def test_imu_cost_functor_and_jacobian(ref, rot_only):
    """SwitchedFullImuCostFunction (ceres-cost-functions.h:379-490) with ceres::Jet<double, 35> + LocalParamSe3's Jacobian,
    as AutoDiffCostFunction evaluates it, vs the oracle's EvalImu: residual and 9 x 33 tangent Jacobian."""
    from oracle.binding import Oracle

    p = _problem()
    rng = np.random.default_rng(3)
    p.b = 1e-2 * rng.standard_normal(6)
    p.sf = 1 + 1e-2 * rng.standard_normal(6)
    p.ts = 0.0021
    p.v_w = p.truth["v_w"] + 1e-2 * rng.standard_normal(p.v_w.shape)
    o = Oracle(p, inertial=1, rotation_only=rot_only, bias_active=1, scale_active=1, optimize_ts=1)
    W = rng.standard_normal((p.n_frames - 1, 9, 9)) + 30 * np.eye(9)
    W = 0.5 * (W + W.transpose(0, 2, 1))
    o.set_imu_weights(W)
    r_o, J_o = o.eval_imu()
    for k in range(p.n_frames - 1):
        r, J = np.zeros(9), np.zeros((9, 33))
        rc = ref.ref_imu_eval(*_imu(p), C.c_double(p.ftime[k]), C.c_double(p.ftime[k + 1]), _p(_c(W[k])), C.c_int(rot_only),
                              _p(_c(p.T_wp[k + 1])), _p(_c(p.T_wp[k])), _p(_c(p.v_w[k + 1])), _p(_c(p.v_w[k])), _p(_c(p.g)),
                              _p(_c(p.b)), _p(_c(p.sf)), C.c_double(p.ts), _p(r), _p(J))
        assert rc == 0
        assert np.abs(r - r_o[k]).max() <= 1e-11 * max(1.0, np.abs(r_o[k]).max()), k
        assert np.abs(J - J_o[k]).max() <= 1e-11 * np.abs(J_o[k]).max(), k


def test_update_imu_weights(ref):
    """The loop body of ViCalibrator::UpdateImuWeights (vicalibrator.h:726-796) on the reference's double integrator with
    Jacobians + covariance (types.h:330-687) and derivative tables (vicalibrator-utils.h:106-434) vs the oracle's
    restatement: the 9 x 9 weight_sqrt_ of every interval."""
    from oracle.binding import Oracle

    p = _problem(n_frames=16)
    rng = np.random.default_rng(5)
    p.b = 1e-2 * rng.standard_normal(6)
    p.sf = 1 + 1e-2 * rng.standard_normal(6)
    p.ts = 0.0013
    p.v_w = p.truth["v_w"].copy()
    o = Oracle(p, inertial=1, bias_active=1, scale_active=1, optimize_ts=1)
    o.update_imu_weights()
    W_o = o.imu_weights()
    for k in range(p.n_frames - 1):
        W = np.zeros((9, 9))
        m = C.c_double()
        rc = ref.ref_update_weight(*_imu(p), C.c_double(p.ftime[k]), C.c_double(p.ftime[k + 1]), _p(_c(p.T_wp[k])), _p(_c(p.v_w[k])),
                                   _p(_c(p.T_wp[k + 1])), _p(_c(p.v_w[k + 1])), _p(_c(p.g)), _p(_c(p.b)), _p(_c(p.sf)),
                                   C.c_double(p.ts), C.c_double(synth.GYRO_SIGMA), C.c_double(synth.ACCEL_SIGMA), _p(W), C.byref(m))
        assert rc == 1
        assert np.abs(W - W_o[k]).max() <= 1e-8 * np.abs(W_o[k]).max(), (k, np.abs(W - W_o[k]).max() / np.abs(W_o[k]).max())


def test_local_parameterisations(ref):
    """LocalParamSe3 / LocalParamSo3 Plus and ComputeJacobian (local-param-se3.h:14-91, 107-157) vs the oracle's."""
    from oracle import binding

    rng = np.random.default_rng(9)
    for _ in range(50):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        x = np.concatenate([q, rng.standard_normal(3)])
        for scale in (1e-12, 1e-3, 0.7):
            d = scale * rng.standard_normal(6)
            out = np.zeros(7)
            ref.ref_se3_plus(_p(x), _p(d), _p(out))
            assert np.abs(out - binding.se3_plus(x, d)).max() <= 1e-14
            out4 = np.zeros(4)
            ref.ref_so3_plus(_p(_c(q)), _p(_c(d[3:])), _p(out4))
            assert np.abs(out4 - binding.so3_plus(q, d[3:])).max() <= 1e-14
        # the Jacobians are the derivative of Plus at delta = 0 (central differences of the REFERENCE's Plus)
        J = np.zeros((7, 6))
        ref.ref_se3_jacobian(_p(x), _p(J))
        h = 1e-6
        for c in range(6):
            e = np.zeros(6)
            e[c] = h
            a, b = np.zeros(7), np.zeros(7)
            ref.ref_se3_plus(_p(x), _p(e), _p(a))
            ref.ref_se3_plus(_p(x), _p(-e), _p(b))
            assert np.abs((a - b) / (2 * h) - J[:, c]).max() <= 1e-8


@pytest.mark.parametrize("model", ["fov", "poly2", "poly3", "kb4", "linear"])
def test_reprojection_functor_pose_chain(ref, model):
    """ImuReprojectionCostFunctor (ceres-cost-functions.h:342-377) with Jet<double, 22> and the reference's local
    parameterisations vs the oracle's EvalReprojection: residual and 2 x (6 + 3 + 3 + K) tangent Jacobian.  (Project itself
    is the oracle's restatement of Calibu on both sides: this pins the pose chain and the tangent-space algebra.)"""
    from oracle.binding import Oracle

    p = synth.make_problem(models=(model, "poly3"), n_frames=4, seed=12, inertial=True)
    o = Oracle(p, inertial=1)
    r_o, J_o = o.eval_reproj()
    sel = np.where(p.obs_cam == 0)[0][::17]
    for i in sel:
        f = p.obs_frame[i]
        r, J = np.zeros(2), np.zeros((2, 22))
        rc = ref.ref_reproj(C.c_int(int(p.models[0])), _p(_c(p.T_wp[f])), _p(_c(p.q_ck[0])), _p(_c(p.p_ck[0])), _p(_c(p.intr[0])),
                            _p(_c(p.p_w[i])), _p(_c(p.p_c[i])), _p(r), _p(J))
        assert rc == 0
        assert np.abs(r - r_o[i]).max() <= 1e-10
        assert np.abs(J - J_o[i]).max() <= 1e-11 * np.abs(J_o[i]).max()
