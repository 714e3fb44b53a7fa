"""ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).

ctypes wrapper over oracle/liboracle.so (the CPU restatement of the reference's calibration
solve).  Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.vo_create.restype = C.c_void_p
        L.vo_evaluate.restype = C.c_double
        L.vo_evaluate_camera.restype = C.c_double
        for name in ("vo_frame_dim", "vo_num_globals", "vo_num_residuals", "vo_remove_outliers", "vo_solve",
                     "vo_solve_arrow", "vo_imu_get_range"):
            getattr(L, name).restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """One CPU problem instance; mirrors the C-ABI call sequence of the product library."""

    def __init__(self, prob=None, **flags):
        self.L = lib()
        self.h = C.c_void_p(self.L.vo_create())
        if prob is not None:
            self.load(prob)
            self.set_flags(**flags)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vo_destroy(self.h)
            self.h = None

    # ---- uploads
    def load(self, p):
        self.models = _c(p.models, np.int32)
        self.n_cams, self.n_frames, self.n_obs = p.n_cams, p.n_frames, p.n_obs
        self._obs_cam = np.asarray(p.obs_cam)
        self.L.vo_set_cameras(self.h, C.c_int(p.n_cams), _p(self.models), _p(_c(p.intr)), _p(_c(p.q_ck)),
                              _p(_c(p.p_ck)))
        self.L.vo_set_frames(self.h, C.c_int(p.n_frames), _p(_c(p.T_wp)), _p(_c(p.v_w)), _p(_c(p.ftime)))
        self.L.vo_set_observations(self.h, C.c_int64(p.n_obs), _p(_c(p.obs_frame, np.int32)),
                                   _p(_c(p.obs_cam, np.int32)), _p(_c(p.p_w)), _p(_c(p.p_c)))
        from vicalib_b200.synth import ACCEL_SIGMA, GYRO_SIGMA

        self.L.vo_set_imu(self.h, C.c_int(len(p.imu_t)), _p(_c(p.imu_t)), _p(_c(p.imu_w)), _p(_c(p.imu_a)),
                          C.c_double(GYRO_SIGMA), C.c_double(ACCEL_SIGMA))
        self.L.vo_set_imu_params(self.h, _p(_c(p.g)), _p(_c(p.b)), _p(_c(p.sf)), C.c_double(p.ts))

    def set_frames(self, T_wp, v_w, ftime):
        self.L.vo_set_frames(self.h, C.c_int(len(ftime)), _p(_c(T_wp)), _p(_c(v_w)), _p(_c(ftime)))

    def set_cameras(self, models, intr, q_ck, p_ck):
        self.L.vo_set_cameras(self.h, C.c_int(len(models)), _p(_c(models, np.int32)), _p(_c(intr)),
                              _p(_c(q_ck)), _p(_c(p_ck)))

    def set_imu_params(self, g, b, sf, ts):
        self.L.vo_set_imu_params(self.h, _p(_c(g)), _p(_c(b)), _p(_c(sf)), C.c_double(ts))

    def set_flags(self, inertial=0, rotation_only=0, bias_active=0, scale_active=0, optimize_ts=0,
                  fix_intrinsics=0, visual=1, visual_mult=1.0, imu_mult=1.0):
        self.L.vo_set_flags(self.h, int(inertial), int(rotation_only), int(bias_active), int(scale_active),
                            int(optimize_ts), int(fix_intrinsics), int(visual), C.c_double(visual_mult),
                            C.c_double(imu_mult))

    def set_options(self, max_iters=200, function_tol=1e-6, gradient_tol=1e-10, param_tol=1e-8,
                    init_radius=1e4, strategy=0, jacobi_scaling=1, num_threads=1, update_imu_weights=1):
        self.L.vo_set_options(self.h, int(max_iters), C.c_double(function_tol), C.c_double(gradient_tol),
                              C.c_double(param_tol), C.c_double(init_radius), int(strategy),
                              int(jacobi_scaling), int(num_threads), int(update_imu_weights))

    # ---- queries
    @property
    def fd(self):
        return self.L.vo_frame_dim(self.h)

    @property
    def G(self):
        return self.L.vo_num_globals(self.h)

    def num_residuals(self):
        return self.L.vo_num_residuals(self.h)

    def global_mask(self):
        m = np.zeros(self.G)
        self.L.vo_global_mask(self.h, _p(m))
        return m

    def eval_reproj(self, i0=0, n=None, jac=True):
        n = self.n_obs - i0 if n is None else n
        r = np.zeros((n, 2))
        J = np.zeros((n, 2, 22)) if jac else None
        self.L.vo_eval_reproj(self.h, C.c_int64(i0), C.c_int64(n), _p(r), _p(J))
        return r, J

    def eval_imu(self, k0=0, n=None, jac=True):
        n = self.n_frames - 1 - k0 if n is None else n
        r = np.zeros((n, 9))
        J = np.zeros((n, 9, 33)) if jac else None
        self.L.vo_eval_imu(self.h, int(k0), int(n), _p(r), _p(J))
        return r, J

    def cost(self):
        return self.L.vo_evaluate(self.h, None, None, None, None, None, None)

    def normal_equations(self):
        nf, fd, G = self.n_frames, self.fd, self.G
        out = dict(B=np.zeros((nf, fd, fd)), U=np.zeros((nf, fd, fd)), E=np.zeros((nf, fd, G)),
                   gf=np.zeros((nf, fd)), C=np.zeros((G, G)), gc=np.zeros(G))
        out["cost"] = self.L.vo_evaluate(self.h, _p(out["B"]), _p(out["U"]), _p(out["E"]), _p(out["gf"]),
                                         _p(out["C"]), _p(out["gc"]))
        return out

    def evaluate_camera(self, cam, residuals=False):
        n = int((self.obs_cam_count(cam)))
        res = np.zeros(2 * n) if residuals else None
        c = self.L.vo_evaluate_camera(self.h, int(cam), _p(res))
        return (c, res) if residuals else c

    def obs_cam_count(self, cam):
        act = np.zeros(self.n_obs, dtype=np.uint8)
        self.L.vo_get_obs_active(self.h, _p(act))
        return int(((self._obs_cam == cam) & (act > 0)).sum()) if hasattr(self, "_obs_cam") else self.n_obs

    def remove_outliers(self, rmse, threshold):
        return self.L.vo_remove_outliers(self.h, _p(_c(rmse)), C.c_double(threshold))

    def obs_active(self):
        act = np.zeros(self.n_obs, dtype=np.uint8)
        self.L.vo_get_obs_active(self.h, _p(act))
        return act

    def update_imu_weights(self):
        self.L.vo_update_imu_weights(self.h)

    def imu_weights(self):
        w = np.zeros((max(self.n_frames - 1, 0), 9, 9))
        self.L.vo_get_imu_weights(self.h, _p(w))
        return w

    def set_imu_weights(self, w):
        self.L.vo_set_imu_weights(self.h, _p(_c(w)))

    def solve_arrow(self, scale, D2):
        x = np.zeros(self.n_frames * self.fd + self.G)
        rc = self.L.vo_solve_arrow(self.h, _p(_c(scale)), _p(_c(D2)), _p(x))
        if rc != 0:
            raise RuntimeError("arrow system not positive definite")
        return x

    def plus(self, delta):
        self.L.vo_plus(self.h, _p(_c(delta)))

    def solve(self, max_rows=1024):
        summ = np.zeros(7)
        rows = np.zeros((max_rows, 9))
        nr = self.L.vo_solve(self.h, _p(summ), _p(rows), int(max_rows))
        keys = ("iterations", "successful_steps", "initial_cost", "final_cost", "termination",
                "num_residuals", "seconds")
        out = dict(zip(keys, summ.tolist()))
        out["rows"] = rows[: min(nr, max_rows)]
        return out

    def state(self):
        s = dict(intr=np.zeros((self.n_cams, 10)), q_ck=np.zeros((self.n_cams, 4)),
                 p_ck=np.zeros((self.n_cams, 3)), T_wp=np.zeros((self.n_frames, 7)),
                 v_w=np.zeros((self.n_frames, 3)), g=np.zeros(2), b=np.zeros(6), sf=np.zeros(6))
        ts = C.c_double(0)
        self.L.vo_get_state(self.h, _p(s["intr"]), _p(s["q_ck"]), _p(s["p_ck"]), _p(s["T_wp"]), _p(s["v_w"]),
                            _p(s["g"]), _p(s["b"]), _p(s["sf"]), C.byref(ts))
        s["ts"] = ts.value
        return s

    def imu_get_range(self, t0, t1, ts, max_n=256):
        out = np.zeros((max_n, 7))
        n = self.L.vo_imu_get_range(self.h, C.c_double(t0), C.c_double(t1), C.c_double(ts), _p(out), int(max_n))
        return out[:n]


def se3_exp(d):
    out = np.zeros(7)
    lib().vo_se3_exp(_p(_c(d)), _p(out))
    return out


def se3_log(x):
    out = np.zeros(6)
    lib().vo_se3_log(_p(_c(x)), _p(out))
    return out


def se3_plus(x, d):
    out = np.zeros(7)
    lib().vo_se3_plus(_p(_c(x)), _p(_c(d)), _p(out))
    return out


def so3_plus(x, d):
    out = np.zeros(4)
    lib().vo_so3_plus(_p(_c(x)), _p(_c(d)), _p(out))
    return out


def project(model, ray, params):
    out = np.zeros(2)
    lib().vo_project(int(model), _p(_c(ray)), _p(_c(params)), _p(out))
    return out
