"""ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned for this file: Calibu's PosePnPRansac and OpenCV's solvePnP are
un-vendored dependencies of the reference (call site src/vicalib-task.cc:322-325, `PosePnPRansac(camera, ellipses,
target.Circles3D(), ellipse_target_map, 0, 0, &t_cw)`); the published algorithm is restated here and anchored on
synthetic truth poses (tests/test_cpu_pnp_oracle.py).

Pose of the PLANAR calibration target (p_w = spacing * (gx, gy, 0), vicalib-task.cc:357-358) in one camera from 2-D / 3-D
correspondences, the way Calibu + OpenCV's iterative solvePnP do it for coplanar points:

  1. unproject every detected centre through the camera model to normalised coordinates (ray / ray_z) — Calibu passes
     unprojected points and an identity camera matrix to OpenCV;
  2. homography target plane -> normalised image (Hartley-normalised DLT; smallest eigenvector of A^T A);
  3. pose from the homography columns (r1, r2 normalised, r3 = r1 x r2, re-orthogonalised), target in front of the camera;
  4. Levenberg-Marquardt on the 6-DoF pose (left perturbation), reprojection error in normalised coordinates;
  robust_its > 0: RANSAC over 4-point homographies first (deterministic sample sequence), steps 2-4 on the inliers.

Returns T_cw as (qx, qy, qz, qw, tx, ty, tz): p_c = R p_w + t.  The frame pose the calibrator is seeded with is
T_wp = T_cw^-1 * T_ck (vicalib-task.cc:341-349).
"""
from __future__ import annotations

import numpy as np

LINEAR, FOV, POLY2, POLY3, KB4 = 0, 1, 2, 3, 4
LM_ITERS = 12
NEWTON_ITERS = 8


def unproject(model: int, pix, p):
    """pixel -> normalised coordinates (x, y) = ray.xy / ray.z; inverse of synth.project / SURVEY App. A.2"""
    pix = np.asarray(pix, dtype=np.float64)
    xd = (pix[..., 0] - p[2]) / p[0]
    yd = (pix[..., 1] - p[3]) / p[1]
    rd = np.sqrt(xd * xd + yd * yd)
    if model == LINEAR:
        return np.stack([xd, yd], -1)
    if model == FOV:
        w = p[4]
        m = 2.0 * np.tan(w / 2.0)
        ru = np.tan(np.minimum(rd * w, 1.5)) / m
    elif model in (POLY2, POLY3):
        k = (p[4], p[5], p[6] if model == POLY3 else 0.0)
        ru = rd.copy()
        for _ in range(NEWTON_ITERS):  # Newton on r_u * f(r_u) = r_d
            r2 = ru * ru
            f = 1 + r2 * (k[0] + r2 * (k[1] + r2 * k[2]))
            df = f + ru * ru * (2 * k[0] + r2 * (4 * k[1] + r2 * 6 * k[2]))
            ru = ru - (ru * f - rd) / df
    elif model == KB4:
        th = rd.copy()
        for _ in range(NEWTON_ITERS):  # Newton on theta + k0 th^3 + ... = r_d
            t2 = th * th
            d = th * (1 + t2 * (p[4] + t2 * (p[5] + t2 * (p[6] + t2 * p[7]))))
            dd = 1 + t2 * (3 * p[4] + t2 * (5 * p[5] + t2 * (7 * p[6] + t2 * 9 * p[7])))
            th = th - (d - rd) / dd
        ru = np.tan(np.minimum(th, 1.5))
    else:
        raise ValueError(model)
    small = rd <= 1e-12  # on the optical axis: the limit of r_u / r_d
    limit = p[4] / (2.0 * np.tan(p[4] / 2.0)) if model == FOV else 1.0
    s = np.where(small, limit, ru / np.where(small, 1.0, rd))
    return np.stack([xd * s, yd * s], -1)


def _hartley(pts):
    c = pts.mean(0)
    d = np.sqrt(((pts - c) ** 2).sum(1)).mean()
    s = np.sqrt(2.0) / d if d > 0 else 1.0
    return np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1.0]])


def jacobi_eigh(A, sweeps=30):
    """cyclic Jacobi eigen-decomposition (the device uses the same sweep order): returns (eigenvalues, eigenvectors)"""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    V = np.eye(n)
    for _ in range(sweeps):
        off = sum(A[i, j] ** 2 for i in range(n) for j in range(i + 1, n))
        if off <= 1e-30 * (np.diag(A) ** 2).sum():
            break
        for p in range(n):
            for q in range(p + 1, n):
                if A[p, q] == 0.0:
                    continue
                tau = (A[q, q] - A[p, p]) / (2.0 * A[p, q])
                t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1.0 + tau * tau))
                c = 1.0 / np.sqrt(1.0 + t * t)
                s = t * c
                Ap, Aq = A[:, p].copy(), A[:, q].copy()
                A[:, p], A[:, q] = c * Ap - s * Aq, s * Ap + c * Aq
                Ap, Aq = A[p, :].copy(), A[q, :].copy()
                A[p, :], A[q, :] = c * Ap - s * Aq, s * Ap + c * Aq
                Vp, Vq = V[:, p].copy(), V[:, q].copy()
                V[:, p], V[:, q] = c * Vp - s * Vq, s * Vp + c * Vq
    return np.diag(A).copy(), V


def homography(XY, xy):
    """H with xy ~ H (X, Y, 1): normalised DLT, null vector = eigenvector of the smallest eigenvalue of A^T A"""
    Ta, Tb = _hartley(XY), _hartley(xy)
    a = (Ta @ np.c_[XY, np.ones(len(XY))].T).T
    b = (Tb @ np.c_[xy, np.ones(len(xy))].T).T
    M = np.zeros((9, 9))
    for (X, Y, _), (x, y, _) in zip(a, b):
        r1 = np.array([X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x])
        r2 = np.array([0, 0, 0, X, Y, 1, -y * X, -y * Y, -y])
        M += np.outer(r1, r1) + np.outer(r2, r2)
    lam, V = jacobi_eigh(M)
    h = V[:, int(np.argmin(lam))]
    return np.linalg.inv(Tb) @ h.reshape(3, 3) @ Ta


def pose_from_homography(H):
    h1, h2, h3 = H[:, 0], H[:, 1], H[:, 2]
    n1, n2 = np.linalg.norm(h1), np.linalg.norm(h2)
    t = h3 * (2.0 / (n1 + n2))
    r1, r2 = h1 / n1, h2 / n2
    if t[2] < 0:  # the target is in front of the camera
        t, r1, r2 = -t, -r1, -r2
    r3 = np.cross(r1, r2)
    r3 /= np.linalg.norm(r3)
    r2 = np.cross(r3, r1)
    return np.stack([r1, r2, r3], 1), t


def _exp_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])
    if th < 1e-10:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / (th * th) * K @ K


def _cost_and_system(R, t, XYZ, xy):
    p = XYZ @ R.T + t
    iz = 1.0 / p[:, 2]
    r = np.stack([p[:, 0] * iz - xy[:, 0], p[:, 1] * iz - xy[:, 1]], 1)
    JtJ, Jtr = np.zeros((6, 6)), np.zeros(6)
    for (px, py, pz), izz, ri in zip(p, iz, r):
        dpi = np.array([[izz, 0, -px * izz * izz], [0, izz, -py * izz * izz]])
        dp = np.concatenate([np.eye(3), -np.array([[0, -pz, py], [pz, 0, -px], [-py, px, 0.0]])], 1)  # [I | -[p]x]
        J = dpi @ dp
        JtJ += J.T @ J
        Jtr += J.T @ ri
    return float((r * r).sum()), JtJ, Jtr


def refine(R, t, XYZ, xy, iters=LM_ITERS):
    lam = 1e-3
    cost, JtJ, Jtr = _cost_and_system(R, t, XYZ, xy)
    for _ in range(iters):
        A = JtJ + lam * np.diag(np.diag(JtJ))
        d = np.linalg.solve(A, -Jtr)
        dR = _exp_so3(d[3:])
        R2, t2 = dR @ R, dR @ t + d[:3]
        c2, JtJ2, Jtr2 = _cost_and_system(R2, t2, XYZ, xy)
        if c2 < cost:
            R, t, cost, JtJ, Jtr = R2, t2, c2, JtJ2, Jtr2
            lam = max(lam * 0.1, 1e-9)
        else:
            lam = min(lam * 10.0, 1e6)
    return R, t, cost


def mat_to_quat(R):
    from vicalib_b200.synth import mat_to_quat as m2q

    return m2q(R)


def _sample4(view, it, n):
    """deterministic 4 distinct indices for RANSAC draw `it` of `view` (the device hashes the same way)"""
    idx = []
    s = (view * 2654435761 + it * 40503 + 12345) & 0xFFFFFFFF
    while len(idx) < 4:
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        k = (s >> 8) % n
        if k not in idx:
            idx.append(k)
    return idx


def pnp_planar(model, intr, pix, pw, robust_its=0, robust_tol=0.0, view=0):
    """-> (T_cw[7], rmse in normalised coordinates, number of points used) or None with fewer than 4 points"""
    pix, pw = np.asarray(pix, dtype=np.float64), np.asarray(pw, dtype=np.float64)
    n = len(pix)
    if n < 4:
        return None
    xy = unproject(model, pix, intr)
    XY = pw[:, :2]
    use = np.ones(n, dtype=bool)
    if robust_its > 0:
        best = -1
        for it in range(robust_its):
            s = _sample4(view, it, n)
            H = homography(XY[s], xy[s])
            q = (H @ np.c_[XY, np.ones(n)].T).T
            with np.errstate(divide="ignore", invalid="ignore"):  # a degenerate sample maps points to infinity
                e = np.sqrt(((q[:, :2] / q[:, 2:3] - xy) ** 2).sum(1))
            inl = e < robust_tol
            if inl.sum() > best:
                best, use = int(inl.sum()), inl
        if use.sum() < 4:
            return None
    R, t = pose_from_homography(homography(XY[use], xy[use]))
    R, t, cost = refine(R, t, pw[use], xy[use])
    return np.concatenate([mat_to_quat(R[None])[0], t]), float(np.sqrt(cost / use.sum())), int(use.sum())
