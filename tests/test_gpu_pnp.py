"""GPU: batched pose initialisation (PosePnPRansac at src/vicalib-task.cc:322-325) — one warp per (frame, camera)
view — vs the CPU restatement (oracle/pnp.py) on the same correspondences, and vs the synthetic truth poses."""
import numpy as np
import pytest

from vicalib_b200 import synth

pytestmark = pytest.mark.gpu


def _views(p):
    """(cam, start, count) of every (camera, frame) group; observations are sorted by (camera, frame)"""
    key = p.obs_cam.astype(np.int64) * p.n_frames + p.obs_frame
    start = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    count = np.diff(np.r_[start, len(key)])
    return p.obs_cam[start].astype(np.int32), start.astype(np.int64), count.astype(np.int32), p.obs_frame[start]


def _truth_T_cw(p, c, f):
    Rck = synth.quat_to_mat(p.truth["q_ck"][c])
    Rwk = synth.quat_to_mat(p.truth["T_wp"][f, :4])
    R = Rck @ Rwk.T
    return R, p.truth["p_ck"][c] - R @ p.truth["T_wp"][f, 4:]


@pytest.mark.parametrize("models", [("poly3", "fov"), ("kb4", "poly2"), ("linear",)])
def test_batched_pnp_matches_oracle_and_truth(models):
    from oracle import pnp
    from vicalib_b200.capi import Calibrator

    p = synth.make_problem(models=models, n_frames=12, seed=21, pixel_sigma=0.05)
    p.intr = p.truth["intr"].copy()  # the pose initialisation runs with the current intrinsics guess; here: the truth
    g = Calibrator()
    g.set_cameras(p.models, p.intr, p.q_ck, p.p_ck)
    cam, start, count, frame = _views(p)
    T, rmse, used = g.pose_pnp_ransac(cam, start, count, p.p_c, p.p_w)
    assert np.array_equal(used, count)
    for v in range(len(cam)):
        sl = slice(start[v], start[v] + count[v])
        To, ro, no = pnp.pnp_planar(int(p.models[cam[v]]), p.intr[cam[v]], p.p_c[sl], p.p_w[sl], view=v)
        assert np.abs(T[v] - To).max() <= 1e-8, (v, np.abs(T[v] - To).max())  # same minimum of the same cost
        assert abs(rmse[v] - ro) <= 1e-10
        R, t = _truth_T_cw(p, cam[v], frame[v])
        Re = synth.quat_to_mat(T[v, :4])
        assert np.arccos(np.clip((np.trace(Re.T @ R) - 1) / 2, -1, 1)) < 2e-3 and np.abs(T[v, 4:] - t).max() < 2e-3


def test_batched_pnp_ransac_and_degenerate_views():
    from oracle import pnp
    from vicalib_b200.capi import Calibrator

    p = synth.make_problem(models=("poly3",), n_frames=8, seed=22, pixel_sigma=0.05)
    p.intr = p.truth["intr"].copy()
    pix = p.p_c.copy()
    pix[::9] += 25.0  # wrong associations
    g = Calibrator()
    g.set_cameras(p.models, p.intr, p.q_ck, p.p_ck)
    cam, start, count, frame = _views(p)
    count = count.copy()
    count[3] = 3  # a view with too few points: no pose, n_used = 0
    T, rmse, used = g.pose_pnp_ransac(cam, start, count, pix, p.p_w, robust_its=40, robust_tol=3.0 / 300)
    assert used[3] == 0 and np.array_equal(T[3], [0, 0, 0, 1, 0, 0, 0])
    for v in range(len(cam)):
        if v == 3:
            continue
        sl = slice(start[v], start[v] + count[v])
        To, ro, no = pnp.pnp_planar(synth.POLY3, p.intr[0], pix[sl], p.p_w[sl], robust_its=40, robust_tol=3.0 / 300, view=v)
        assert used[v] == no
        assert np.abs(T[v] - To).max() <= 1e-8
        R, t = _truth_T_cw(p, 0, frame[v])
        Re = synth.quat_to_mat(T[v, :4])
        assert np.arccos(np.clip((np.trace(Re.T @ R) - 1) / 2, -1, 1)) < 2e-3 and np.abs(T[v, 4:] - t).max() < 2e-3
