"""Generates the known-answer fixtures in tests/golden/ (run once, here; fixtures are committed).

The reference pins nothing for this path (its only test needs a tarball fetched at configure time
and uses the linear model, testing/vi_sim_test.cpp:18-21), and none of Ceres / Calibu / Sophus /
Eigen can be imported in this container.  These vectors are therefore produced by an INDEPENDENT
restatement in 60-digit mpmath arithmetic — rotation matrices instead of quaternions, Rodrigues
formulae instead of Sophus' quaternion exp/log, exact central differences (h = 1e-25) instead of
dual numbers — of the same published formulas:
    reprojection  r = Project(R_ck R_wk^T (p_w - t_wk) + p_ck) - z       ceres-cost-functions.h:361-370
    IMU           r = W^T [log(T_end T_2^-1); v_end - v2], RK4 integration ceres-cost-functions.h:138-177,402-484
Both the oracle (C++, float64 dual numbers) and the CUDA kernels are checked against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

LINEAR, FOV, POLY2, POLY3, KB4 = range(5)
NUM_INTR = {LINEAR: 4, FOV: 5, POLY2: 6, POLY3: 7, KB4: 8}


def M(a):
    return mp.matrix([[mp.mpf(float(x)) for x in row] for row in np.atleast_2d(a)])


def V(a):
    return mp.matrix([mp.mpf(float(x)) for x in a])


def quat_R(q):
    x, y, z, w = [mp.mpf(float(v)) for v in q]
    n = mp.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def hat(w):
    return mp.matrix([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def so3_exp(w):
    th = mp.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2)
    W = hat(w)
    if th == 0:
        return mp.eye(3)
    return mp.eye(3) + mp.sin(th) / th * W + (1 - mp.cos(th)) / th ** 2 * W * W


def so3_V(w):
    th = mp.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2)
    W = hat(w)
    if th == 0:
        return mp.eye(3)
    return mp.eye(3) + (1 - mp.cos(th)) / th ** 2 * W + (th - mp.sin(th)) / th ** 3 * W * W


def so3_log(R):
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) / 2
    th = mp.acos(c)
    if th == 0:
        return mp.matrix([0, 0, 0])
    f = th / (2 * mp.sin(th))
    return mp.matrix([f * (R[2, 1] - R[1, 2]), f * (R[0, 2] - R[2, 0]), f * (R[1, 0] - R[0, 1])])


def se3_log(R, t):
    w = so3_log(R)
    return list(mp.inverse(so3_V(w)) * t) + list(w)


def se3_plus(R, t, d):  # T * exp(upsilon, omega)   (local-param-se3.h:14-26)
    ups, om = mp.matrix(d[:3]), mp.matrix(d[3:])
    return R * so3_exp(om), t + R * (so3_V(om) * ups)


def project(model, p, k):
    X, Y, Z = p
    if model == KB4:
        rho = mp.sqrt(X * X + Y * Y)
        th = mp.atan2(rho, Z)
        d = th + k[4] * th ** 3 + k[5] * th ** 5 + k[6] * th ** 7 + k[7] * th ** 9
        return [k[0] * d * X / rho + k[2], k[1] * d * Y / rho + k[3]]
    u, v = X / Z, Y / Z
    r2 = u * u + v * v
    if model == LINEAR:
        f = 1
    elif model == FOV:  # Calibu FovCamera::Factor: the two small-argument branches are part of the model
        w = k[4]
        if w * w > mp.mpf("1e-5"):
            if r2 < mp.mpf("1e-5"):
                f = 2 * mp.tan(w / 2) / w
            else:
                rad = mp.sqrt(r2)
                f = mp.atan(rad * 2 * mp.tan(w / 2)) / (rad * w)
        else:
            f = mp.mpf(1)
    elif model == POLY2:
        f = 1 + k[4] * r2 + k[5] * r2 ** 2
    else:
        f = 1 + k[4] * r2 + k[5] * r2 ** 2 + k[6] * r2 ** 3
    return [k[0] * f * u + k[2], k[1] * f * v + k[3]]


def reproj(model, Rwk, twk, Rck, pck, intr, pw, z):
    pc = Rck * (Rwk.T * (pw - twk)) + pck
    pr = project(model, pc, intr)
    return [pr[0] - z[0], pr[1] - z[1]]


def random_sample(rng, model, tgt=None, dist=None):
    """(T_wk, T_ck, intrinsics, world point, measurement); tgt = the point's camera-frame coordinates."""
    K = NUM_INTR[model]
    from vicalib_b200.synth import TRUTH_DIST, quat_to_mat, so3_exp as np_exp
    q = np_exp(rng.normal(0, 0.6, 3)); t = rng.normal(0, 0.2, 3) + np.array([0, 0, -0.6])
    qc = np_exp(rng.normal(0, 0.3, 3)); pc = rng.normal(0, 0.05, 3)
    intr = np.zeros(10)
    intr[:4] = [300 + 30 * rng.random(), 310 + 30 * rng.random(), 320 + 5 * rng.random(), 240 + 5 * rng.random()]
    intr[4:K] = np.array(TRUTH_DIST[model]) * (1 + 0.2 * rng.random(K - 4))
    if dist is not None:
        intr[4:K] = dist
    # a world point in front of the camera: p_w = R_wk (R_ck^T (pc_target - p_ck)) + t
    if tgt is None:
        tgt = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.2, 0.2), rng.uniform(0.4, 0.8)])
    pw = quat_to_mat(q) @ (quat_to_mat(qc).T @ (np.asarray(tgt, float) - pc)) + t
    z = rng.uniform(100, 500, 2)
    return q, t, qc, pc, intr, pw, z


def reproj_kat(rng, model, n=12, samples=None):
    K = NUM_INTR[model]
    rows = []
    if samples is None:
        samples = [random_sample(rng, model) for _ in range(n)]
    for q, t, qc, pc, intr, pw, z in samples:
        Rwk, Rck = quat_R(q), quat_R(qc)
        a = (model, Rwk, V(t), Rck, V(pc), [mp.mpf(float(x)) for x in intr], V(pw), V(z))
        r = reproj(*a)
        h = mp.mpf(10) ** -25
        J = np.zeros((2, 12 + K))
        col = 0

        def fd(fp, fm):
            return [(x - y) / (2 * h) for x, y in zip(fp, fm)]

        for k in range(6):  # frame pose, right perturbation
            d = [mp.mpf(0)] * 6
            d[k] = h
            Rp, tp = se3_plus(Rwk, V(t), d)
            d[k] = -h
            Rm, tm = se3_plus(Rwk, V(t), d)
            g = fd(reproj(model, Rp, tp, Rck, V(pc), a[5], V(pw), V(z)), reproj(model, Rm, tm, Rck, V(pc), a[5], V(pw), V(z)))
            J[:, col] = [float(x) for x in g]; col += 1
        for k in range(3):  # R_ck * exp(w)
            w = [mp.mpf(0)] * 3
            w[k] = h
            Rp = Rck * so3_exp(mp.matrix(w))
            w[k] = -h
            Rm = Rck * so3_exp(mp.matrix(w))
            g = fd(reproj(model, Rwk, V(t), Rp, V(pc), a[5], V(pw), V(z)), reproj(model, Rwk, V(t), Rm, V(pc), a[5], V(pw), V(z)))
            J[:, col] = [float(x) for x in g]; col += 1
        for k in range(3):
            e = mp.matrix([0, 0, 0]); e[k] = h
            g = fd(reproj(model, Rwk, V(t), Rck, V(pc) + e, a[5], V(pw), V(z)), reproj(model, Rwk, V(t), Rck, V(pc) - e, a[5], V(pw), V(z)))
            J[:, col] = [float(x) for x in g]; col += 1
        for k in range(K):
            ip, im = list(a[5]), list(a[5])
            ip[k] += h; im[k] -= h
            g = fd(reproj(model, Rwk, V(t), Rck, V(pc), ip, V(pw), V(z)), reproj(model, Rwk, V(t), Rck, V(pc), im, V(pw), V(z)))
            J[:, col] = [float(x) for x in g]; col += 1
        rows.append(dict(T_wk=np.concatenate([q, t]), q_ck=qc, p_ck=pc, intr=intr, p_w=pw, z=z,
                         r=np.array([float(x) for x in r]), J=J))
    return {k: np.stack([r_[k] for r_ in rows]) for k in rows[0]}


# ------------------------------------------------------------------ IMU
GRAV = mp.mpf("9.8007")


def gravity_vec(g):
    sp, cp, sq, cq = mp.sin(g[0]), mp.cos(g[0]), mp.sin(g[1]), mp.cos(g[1])
    return mp.matrix([-GRAV * cp * sq, GRAV * sp, -GRAV * cp * cq])


def get_range(t, w, a, ts, t0, t1):
    """[interp(t0), samples with t0 < t_i + ts <= t1, interp(t1)] (interpolation-buffer.h:208-226)."""
    tt = [x + ts for x in t]

    def interp(time):
        i = max(k for k in range(len(tt)) if tt[k] <= time)
        f = (time - tt[i]) / (tt[i + 1] - tt[i])
        return (time, w[i] * (1 - f) + w[i + 1] * f, a[i] * (1 - f) + a[i + 1] * f)

    out = [interp(t0)]
    i0 = max(k for k in range(len(tt)) if tt[k] <= t0)
    for k in range(i0 + 1, len(tt)):
        if tt[k] > t1:
            break
        out.append((tt[k], w[k], a[k]))
    out.append(interp(t1))
    return out


def imu_residual(t, w, a, t0, t1, R2, p2, v2, R1, p1, v1, g, b, sf, ts, W, switch):
    gv = gravity_vec(g)
    bg, ba = mp.matrix(b[:3]), mp.matrix(b[3:])
    meas = get_range(t, w, a, ts, t0, t1)
    p, R, v = p1, R1, v1

    def deriv(p_, R_, v_, z0, z1, dt):
        al = (z1[0] - (z0[0] + dt)) / (z1[0] - z0[0])
        zg = z0[1] * al + z1[1] * (1 - al)
        za = z0[2] * al + z1[2] * (1 - al)
        wv = R_ * mp.matrix([zg[i] * sf[i] + bg[i] for i in range(3)])
        av = R_ * mp.matrix([za[i] * sf[3 + i] + ba[i] for i in range(3)]) - gv
        return v_, wv, av

    def step(p_, R_, v_, k, dt):
        return p_ + k[0] * dt, so3_exp(k[1] * dt) * R_, v_ + k[2] * dt

    for z0, z1 in zip(meas[:-1], meas[1:]):
        dt = z1[0] - z0[0]
        if dt == 0:
            continue
        k1 = deriv(p, R, v, z0, z1, 0)
        y1 = step(p, R, v, k1, dt / 2)
        k2 = deriv(*y1, z0, z1, dt / 2)
        y2 = step(p, R, v, k2, dt / 2)
        k3 = deriv(*y2, z0, z1, dt / 2)
        y3 = step(p, R, v, k3, dt)
        k4 = deriv(*y3, z0, z1, dt)
        k = [k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i] for i in range(3)]
        p, R, v = step(p, R, v, k, dt / 6)
    Re = R * R2.T
    te = p - Re * p2
    raw = se3_log(Re, te) + list(v - v2)
    r = [sum(raw[i] * W[i, j] for i in range(9)) for j in range(9)]
    if switch:
        for j in (0, 1, 2, 6, 7, 8):
            r[j] = mp.mpf(0)
    return r


def imu_kat(rng):
    from vicalib_b200.synth import so3_exp as np_exp
    n = 40
    t = 0.1 + 0.005 * np.arange(n) + 0.0003 * rng.random(n)
    w = rng.normal(0, 0.4, (n, 3)); a = rng.normal(0, 1.0, (n, 3)) + np.array([0.3, 9.7, 0.5])
    ftime = np.array([0.123, 0.151, 0.1837, 0.2169])
    nf = len(ftime)
    q = np.stack([np_exp(rng.normal(0, 0.5, 3)) for _ in range(nf)])
    p = rng.normal(0, 0.3, (nf, 3)); v = rng.normal(0, 0.5, (nf, 3))
    g = np.array([0.05, -0.03]); b = np.concatenate([rng.normal(0, 1e-2, 3), rng.normal(0, 1e-1, 3)])
    sf = 1 + 0.02 * rng.normal(size=6); ts = 0.0023
    W = 500 * np.eye(9) + 20 * rng.normal(size=(9, 9))
    tm, wm, am = [mp.mpf(float(x)) for x in t], [V(x) for x in w], [V(x) for x in a]
    Wm = M(W)
    h = mp.mpf(10) ** -25
    out_r, out_J = [], []
    for sw in (0, 1):
        for k in range(nf - 1):
            base = dict(R2=quat_R(q[k + 1]), p2=V(p[k + 1]), v2=V(v[k + 1]), R1=quat_R(q[k]), p1=V(p[k]), v1=V(v[k]),
                        g=[mp.mpf(float(x)) for x in g], b=[mp.mpf(float(x)) for x in b], sf=[mp.mpf(float(x)) for x in sf],
                        ts=mp.mpf(ts))

            def ev(**kw):
                d = dict(base); d.update(kw)
                return imu_residual(tm, wm, am, mp.mpf(float(ftime[k])), mp.mpf(float(ftime[k + 1])), d["R2"], d["p2"], d["v2"],
                                    d["R1"], d["p1"], d["v1"], d["g"], d["b"], d["sf"], d["ts"], Wm, sw)

            r0 = ev()
            cols = []

            def fd(kp, km):
                rp, rm = ev(**kp), ev(**km)
                cols.append([float((x - y) / (2 * h)) for x, y in zip(rp, rm)])

            for which in ("2", "1"):
                R, pp = base["R" + which], base["p" + which]
                for j in range(6):
                    d = [mp.mpf(0)] * 6
                    d[j] = h
                    Rp, tp = se3_plus(R, pp, d)
                    d[j] = -h
                    Rm, tm_ = se3_plus(R, pp, d)
                    fd({"R" + which: Rp, "p" + which: tp}, {"R" + which: Rm, "p" + which: tm_})
            for which in ("v2", "v1"):
                for j in range(3):
                    e = mp.matrix([0, 0, 0]); e[j] = h
                    fd({which: base[which] + e}, {which: base[which] - e})
            for name, cnt in (("g", 2), ("b", 6), ("sf", 6)):
                for j in range(cnt):
                    lp, lm = list(base[name]), list(base[name])
                    lp[j] += h; lm[j] -= h
                    fd({name: lp}, {name: lm})
            fd({"ts": base["ts"] + h}, {"ts": base["ts"] - h})
            out_r.append([float(x) for x in r0])
            out_J.append(np.array(cols).T)
    return dict(imu_t=t, imu_w=w, imu_a=a, ftime=ftime, T_wp=np.concatenate([q, p], 1), v_w=v, g=g, b=b, sf=sf, ts=np.array(ts),
                W=W, r=np.array(out_r).reshape(2, nf - 1, 9), J=np.array(out_J).reshape(2, nf - 1, 9, 33))


def edge_kat():
    """Branch and boundary cases of the camera models (each model's own small-argument branches, the image
    centre, the image border).  Written to reproj_edge_kat.npz with the keys of reproj_kat.npz."""
    rng = np.random.default_rng(20260925)
    out = {}
    cases = {
        "fov": [dict(tgt=[1e-4, -2e-4, 0.5]),                      # rad^2 < 1e-5:  2 tan(w/2) / w
                dict(dist=[2e-3]),                                   # w^2 <= 1e-5:  factor 1
                dict(tgt=[-3e-5, 1e-5, 0.7], dist=[1e-3]),          # both
                dict(tgt=[0.45, -0.3, 0.4])],                        # wide angle (rad ~ 1.35)
        "kb4": [dict(tgt=[1e-5, -2e-5, 0.5]), dict(tgt=[0.5, 0.35, 0.35])],
        "poly3": [dict(tgt=[1e-6, 2e-6, 0.6]), dict(tgt=[0.35, 0.25, 0.4])],
        "poly2": [dict(tgt=[0.0, 0.0, 0.5]), dict(tgt=[-0.35, 0.25, 0.4])],
        "linear": [dict(tgt=[0.0, 0.0, 0.5]), dict(tgt=[0.4, -0.3, 0.4])],
    }
    ids = dict(linear=LINEAR, fov=FOV, poly2=POLY2, poly3=POLY3, kb4=KB4)
    for name, specs in cases.items():
        samples = [random_sample(rng, ids[name], **sp) for sp in specs]
        kat = reproj_kat(rng, ids[name], samples=samples)
        for k, v in kat.items():
            out[f"{name}_{k}"] = v
        print(name, "edge done", flush=True)
    np.savez_compressed(os.path.join(HERE, "reproj_edge_kat.npz"), **out)


if __name__ == "__main__":
    if "--edge" in sys.argv:  # only the branch / boundary fixture (the others are unchanged)
        edge_kat()
        sys.exit(0)
    rng = np.random.default_rng(20260924)
    out = {}
    for name, m in (("linear", LINEAR), ("fov", FOV), ("poly2", POLY2), ("poly3", POLY3), ("kb4", KB4)):
        kat = reproj_kat(rng, m)
        for k, v in kat.items():
            out[f"{name}_{k}"] = v
        print(name, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "reproj_kat.npz"), **out)
    kat = imu_kat(rng)
    np.savez_compressed(os.path.join(HERE, "imu_kat.npz"), **kat)
    print("imu done")
