"""Known-answer vector for UpdateImuWeights (vicalibrator.h:723-799) from an INDEPENDENT computation.

The reference (and oracle/imu_weights.h, vc_imu_weights.cuh after it) propagates the covariance of the
integrated state y = (p, q, v) through every RK4 step with hand-derived Jacobians
    C <- A C A^T + G R G^T,   A = dy/dy0 (10x10),  G = dy/db (10x6),  R = diag(sigma_g^2 x3, sigma_a^2 x3)
(types.h:427-595, vicalibrator-utils.h:187-274), maps it to the residual with dLog_dSE3 * dt1t2_dt1
(vicalibrator-utils.h:260-434), and takes  W = sqrtm((J C J^T)^-1).

Here none of those derivative formulas are used: A, G and J come from exact central differences
(h = 1e-25, 60-digit mpmath) of the integrator and of the residual map themselves, written on the raw
10-vector (quaternion components unconstrained, as the reference differentiates them).  With unit scale
factors the reference's formulas are exact derivatives up to its truncated series for d exp(w)/dw
(relative 1e-9 at |w| dt ~ 5e-3), so the two computations must agree closely; the fixture stores the exact
result and tests/test_cpu_oracle_golden.py states the tolerance.

    python tests/golden/make_weights_golden.py
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import GRAV, V, get_range, gravity_vec  # noqa: E402  (same sample-range and gravity restatement)

mp.mp.dps = 60
H = mp.mpf(10) ** -25


def qmul(a, b):  # (x, y, z, w), Hamilton product a (x) b
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return [aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz]


def qexp(w):  # SO3::exp as a quaternion
    th = mp.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2)
    if th == 0:
        return [w[0] / 2, w[1] / 2, w[2] / 2, mp.mpf(1)]
    s = mp.sin(th / 2) / th
    return [s * w[0], s * w[1], s * w[2], mp.cos(th / 2)]


def rot(q, v):  # the polynomial the reference differentiates (unit-quaternion form, no normalisation)
    x, y, z, w = q
    R = mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R * mp.matrix(v)


def deriv(y, z0, z1, dt_off, b, sf, gv):
    al = (z1[0] - (z0[0] + dt_off)) / (z1[0] - z0[0])
    zg = z0[1] * al + z1[1] * (1 - al)
    za = z0[2] * al + z1[2] * (1 - al)
    q = y[3:7]
    kw = rot(q, [zg[i] * sf[i] + b[i] for i in range(3)])
    ka = rot(q, [za[i] * sf[3 + i] + b[3 + i] for i in range(3)]) - gv
    return list(y[7:10]) + list(kw) + list(ka)


def integrate_pose(y, k, dt):  # types.h:330-378: p += k_v dt, q <- exp(k_w dt) (x) q, v += k_a dt
    p = [y[i] + k[i] * dt for i in range(3)]
    q = qmul(qexp([k[3 + i] * dt for i in range(3)]), y[3:7])
    v = [y[7 + i] + k[6 + i] * dt for i in range(3)]
    return p + q + v


def rk4_step(y, z0, z1, b, sf, gv):  # types.h:427-595 without the Jacobian bookkeeping
    dt = z1[0] - z0[0]
    k1 = deriv(y, z0, z1, 0, b, sf, gv)
    k2 = deriv(integrate_pose(y, k1, dt / 2), z0, z1, dt / 2, b, sf, gv)
    k3 = deriv(integrate_pose(y, k2, dt / 2), z0, z1, dt / 2, b, sf, gv)
    k4 = deriv(integrate_pose(y, k3, dt), z0, z1, dt, b, sf, gv)
    k = [k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i] for i in range(9)]
    return integrate_pose(y, k, dt / 6)


def residual(y, q2inv, t2inv, v2):  # [log(T(y) T_2^-1) (upsilon, omega); v - v2] on the raw 10-vector
    q12 = qmul(y[3:7], q2inv)
    t12 = mp.matrix(y[0:3]) + rot(y[3:7], t2inv)
    n = mp.sqrt(q12[0] ** 2 + q12[1] ** 2 + q12[2] ** 2)
    f = 2 * mp.atan(n / q12[3]) / n  # Sophus SO3::log: 2 atan(|v| / w) / |v|
    om = mp.matrix([f * q12[0], f * q12[1], f * q12[2]])
    th = mp.sqrt(om[0] ** 2 + om[1] ** 2 + om[2] ** 2)
    O = mp.matrix([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    c = (1 - th / (2 * mp.tan(th / 2))) / th ** 2
    ups = (mp.eye(3) - O / 2 + c * O * O) * t12
    return list(ups) + list(om) + [y[7 + i] - v2[i] for i in range(3)]


def fd(fun, x, n_out):
    J = mp.zeros(n_out, len(x))
    for j in range(len(x)):
        xp, xm = list(x), list(x)
        xp[j] += H
        xm[j] -= H
        fp, fm = fun(xp), fun(xm)
        for i in range(n_out):
            J[i, j] = (fp[i] - fm[i]) / (2 * H)
    return J


def main():
    from vicalib_b200.synth import ACCEL_SIGMA, GYRO_SIGMA, so3_exp as np_exp

    rng = np.random.default_rng(20260926)
    n = 24
    t = 0.1 + 0.005 * np.arange(n) + 0.0003 * rng.random(n)
    w = rng.normal(0, 0.6, (n, 3))
    a = rng.normal(0, 1.0, (n, 3)) + np.array([0.3, 9.7, 0.5])
    ftime = np.array([0.1231, 0.1568])
    q1 = np_exp(rng.normal(0, 0.5, 3))
    p1, v1 = rng.normal(0, 0.3, 3), rng.normal(0, 0.5, 3)
    g = np.array([0.05, -0.03])
    b = np.concatenate([rng.normal(0, 1e-2, 3), rng.normal(0, 1e-1, 3)])
    sf = np.ones(6)  # unit scale factors: the reference's d k / d x drops them (types.h:413-423)
    ts = 0.0023
    tm, wm, am = [mp.mpf(float(x)) for x in t], [V(x) for x in w], [V(x) for x in a]
    bm, sfm = [mp.mpf(float(x)) for x in b], [mp.mpf(float(x)) for x in sf]
    gv = gravity_vec([mp.mpf(float(x)) for x in g])
    meas = get_range(tm, wm, am, mp.mpf(ts), mp.mpf(float(ftime[0])), mp.mpf(float(ftime[1])))

    y = [mp.mpf(float(x)) for x in np.concatenate([p1, q1, v1])]
    R = mp.diag([mp.mpf(GYRO_SIGMA) ** 2] * 3 + [mp.mpf(ACCEL_SIGMA) ** 2] * 3)
    C = mp.zeros(10, 10)
    for z0, z1 in zip(meas[:-1], meas[1:]):
        if z1[0] - z0[0] == 0:
            continue
        A = fd(lambda x: rk4_step(x, z0, z1, bm, sfm, gv), y, 10)
        G = fd(lambda x: rk4_step(y, z0, z1, x, sfm, gv), bm, 10)
        C = A * C * A.T + G * R * G.T
        y = rk4_step(y, z0, z1, bm, sfm, gv)

    # frame 2: the integrated pose moved by a generic amount (so that log() is far from its small-angle branch)
    y_end = [float(x) for x in y]
    q2 = np.array(qmul([mp.mpf(float(x)) for x in np_exp(np.array([0.21, -0.13, 0.17]))], y[3:7]), dtype=float)
    q2 /= np.linalg.norm(q2)
    p2 = np.array(y_end[0:3]) + np.array([0.012, -0.02, 0.007])
    v2 = np.array(y_end[7:10]) + np.array([0.03, 0.01, -0.02])
    q2m = [mp.mpf(float(x)) for x in q2]
    q2inv = [-q2m[0], -q2m[1], -q2m[2], q2m[3]]
    t2inv = list(-rot(q2inv, [mp.mpf(float(x)) for x in p2]))
    v2m = [mp.mpf(float(x)) for x in v2]
    J = fd(lambda x: residual(x, q2inv, t2inv, v2m), y, 9)
    P = J * C * J.T
    info = mp.inverse(P)
    info = (info + info.T) / 2
    E, Q = mp.eigsy(info)
    Wm = Q * mp.diag([mp.sqrt(E[i]) for i in range(9)]) * Q.T

    def npm(m):
        return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])

    np.savez_compressed(os.path.join(HERE, "imu_weights_kat.npz"), imu_t=t, imu_w=w, imu_a=a, ftime=ftime,
                        T_wp=np.stack([np.concatenate([q1, p1]), np.concatenate([q2, p2])]), v_w=np.stack([v1, v2]),
                        g=g, b=b, sf=sf, ts=ts, sigma_g=GYRO_SIGMA, sigma_a=ACCEL_SIGMA,
                        C=npm(C), P=npm(P), info=npm(info), W=npm(Wm))
    print("cond(P) = %.3g" % float(mp.norm(P, 2) * mp.norm(info, 2)))


if __name__ == "__main__":
    main()
