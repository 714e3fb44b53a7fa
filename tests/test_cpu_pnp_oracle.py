"""CPU: the PnP restatement (oracle/pnp.py; Calibu PosePnPRansac + OpenCV solvePnP are un-vendored) recovers the synthetic
truth poses for every camera model, with and without gross outliers."""
import numpy as np
import pytest

from oracle import pnp
from vicalib_b200 import synth


def _views(model, n_frames=6, seed=3, pixel_sigma=0.05):
    p = synth.make_problem(models=(model,), n_frames=n_frames, seed=seed, pixel_sigma=pixel_sigma)
    M = p.n_obs // p.n_frames
    for f in range(p.n_frames):
        sl = slice(f * M, (f + 1) * M)
        # truth T_cw = T_ck * T_wk^-1  (vision-only problems: T_ck of camera 0 is the identity)
        R = synth.quat_to_mat(p.truth["T_wp"][f, :4]).T
        t = -R @ p.truth["T_wp"][f, 4:]
        yield f, p.truth["intr"][0], p.p_c[sl], p.p_w[sl], R, t


@pytest.mark.parametrize("model", ["linear", "fov", "poly2", "poly3", "kb4"])
def test_unproject_inverts_project(model):
    m = synth.MODEL_IDS[model]
    rng = np.random.default_rng(1)
    intr = np.zeros(10)
    intr[:4] = [320, 321, 318, 243]
    intr[4:4 + len(synth.TRUTH_DIST[m])] = synth.TRUTH_DIST[m]
    xy = rng.uniform(-0.6, 0.6, (200, 2))
    xy[0] = 0.0  # the optical axis
    pix = synth.project(m, np.c_[xy, np.ones(200)], intr)
    back = pnp.unproject(m, pix, intr)
    assert np.abs(back - xy).max() < 1e-9


@pytest.mark.parametrize("model", ["fov", "poly3", "kb4"])
def test_pnp_recovers_truth_pose(model):
    for f, intr, pix, pw, R, t in _views(model):
        T, rmse, n = pnp.pnp_planar(synth.MODEL_IDS[model], intr, pix, pw, view=f)
        Re = synth.quat_to_mat(T[:4])
        ang = np.arccos(np.clip((np.trace(Re.T @ R) - 1) / 2, -1, 1))
        assert ang < 2e-3 and np.abs(T[4:] - t).max() < 2e-3 and n == len(pix)
        assert rmse < 5 * 0.05 / 300  # pixel noise / focal length


def test_pnp_ransac_rejects_gross_outliers():
    for f, intr, pix, pw, R, t in _views("poly3", n_frames=3):
        pix = pix.copy()
        pix[::9] += 25.0  # wrong associations
        T0, rmse0, _ = pnp.pnp_planar(synth.POLY3, intr, pix, pw, view=f)
        T, rmse, n = pnp.pnp_planar(synth.POLY3, intr, pix, pw, robust_its=40, robust_tol=3.0 / 300, view=f)
        Re = synth.quat_to_mat(T[:4])
        ang = np.arccos(np.clip((np.trace(Re.T @ R) - 1) / 2, -1, 1))
        assert n == len(pix) - len(pix[::9])
        assert ang < 2e-3 and np.abs(T[4:] - t).max() < 2e-3
        assert rmse < rmse0 / 10


def test_too_few_points():
    assert pnp.pnp_planar(synth.POLY3, np.array([300, 300, 320, 240, 0, 0, 0, 0, 0, 0.0]), np.zeros((3, 2)), np.zeros((3, 3))) is None
