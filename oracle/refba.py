"""ctypes binding of the CPU oracle (oracle/refba.cpp).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librefba.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "refba.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, fp, ip, up = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8))
        L.refba_create.restype = C.c_void_p
        L.refba_create.argtypes = [C.c_int, C.c_int, C.c_int, dp, up, dp, ip, ip, fp, dp]
        L.refba_destroy.argtypes = [C.c_void_p]
        L.refba_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.refba_solve_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.refba_solve_global.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.refba_solve_seconds.restype = C.c_double
        L.refba_solve_seconds.argtypes = [C.c_void_p]
        L.refba_get_poses.argtypes = [C.c_void_p, dp]
        L.refba_get_points.argtypes = [C.c_void_p, dp]
        L.refba_get_outliers.argtypes = [C.c_void_p, up]
        L.refba_trace_len.argtypes = [C.c_void_p]
        L.refba_get_trace.argtypes = [C.c_void_p, dp]
        L.refba_linearize_all.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]
        L.refba_schur_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, dp, dp, dp, dp, ip, ip, dp]
        L.refba_pose_oplus.argtypes = [dp, dp, dp]
        L.refba_se3_exp.argtypes = [dp, dp]
        L.refba_cam_project_mono.argtypes = [dp, C.c_double, C.c_double, C.c_double, C.c_double, dp]
        L.refba_cam_project_stereo.argtypes = [dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, dp]
        L.refba_huber.argtypes = [C.c_double, C.c_double, dp]
        L.refba_set_lidar_edges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int]
        L.refba_set_lidar.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp, C.c_int, fp, C.c_int64, fp, ip, C.c_int64, fp, ip,
                                      C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.refba_lidar_num_matches.argtypes = [C.c_void_p]
        L.refba_get_lidar_matches.argtypes = [C.c_void_p, ip]
        L.refba_num_lidar_edges.argtypes = [C.c_void_p]
        L.refba_debug_phase.argtypes = [C.c_void_p, ip, C.c_int, ip, ip, dp, dp, dp, dp, dp]
        L.refba_debug_system.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.refba_sim3_exp.argtypes = [dp, dp]
        L.refba_sim3_log.argtypes = [dp, dp]
        L.refba_sim3_oplus.argtypes = [dp, dp, C.c_int, dp]
        L.refba_sim3_edge_error.argtypes = [dp, dp, dp, dp]
        L.refba_pose_graph.argtypes = [C.c_int, dp, up, C.c_int, C.c_int, ip, dp, C.c_int, C.c_double, dp, C.c_int, ip]
        L.refba_pose_opt.argtypes = [dp, dp, C.c_int, dp, fp, up, dp, C.c_int, ip]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


TRACE_COLS = ("pass", "iter", "trial", "lambda", "chi_before", "chi_trial", "rho", "accepted")


class RefBA:
    """One g2o-style optimiser instance over a flat SoA problem (same layout as include/sqrtba.h)."""

    def __init__(self, prob, threads: int = 1):
        L = lib()
        self.n_pose, self.n_point, self.n_obs = prob.n_pose, prob.n_point, prob.n_obs
        self._keep = [np.ascontiguousarray(prob.pose_qt, np.float64), np.ascontiguousarray(prob.pose_fixed, np.uint8),
                      np.ascontiguousarray(prob.point_xyz, np.float64), np.ascontiguousarray(prob.obs_pose, np.int32),
                      np.ascontiguousarray(prob.obs_point, np.int32), np.ascontiguousarray(prob.obs_meas, np.float32),
                      np.ascontiguousarray(prob.cam, np.float64)]
        k = self._keep
        self.h = L.refba_create(self.n_pose, self.n_point, self.n_obs, _p(k[0], C.c_double), _p(k[1], C.c_uint8),
                                _p(k[2], C.c_double), _p(k[3], C.c_int32), _p(k[4], C.c_int32),
                                _p(k[5], C.c_float), _p(k[6], C.c_double))
        L.refba_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            lib().refba_destroy(self.h)
            self.h = None

    def solve_local(self, third_pass_iters: int = 0, stop=None):
        return lib().refba_solve_local(self.h, stop, third_pass_iters)

    def solve_global(self, iters: int, robust: bool, stop=None):
        return lib().refba_solve_global(self.h, iters, int(robust), stop)

    def debug_phase(self, levels, robust, update=None):
        """initializeOptimization(0) + computeActiveErrors + chi2 [+ update] with the given edge levels (refba_debug_phase)."""
        lv = np.ascontiguousarray(levels, np.int32)
        pi, li = np.zeros(self.n_pose, np.int32), np.zeros(self.n_point, np.int32)
        err, chi = np.zeros((self.n_obs, 3)), np.zeros(2)
        P, X = np.zeros((self.n_pose, 7)), np.zeros((self.n_point, 3))
        up = None if update is None else np.ascontiguousarray(update, np.float64)
        lib().refba_debug_phase(self.h, _p(lv, C.c_int32), int(robust), _p(pi, C.c_int32), _p(li, C.c_int32),
                                _p(err, C.c_double), _p(chi, C.c_double), None if up is None else _p(up, C.c_double),
                                _p(P, C.c_double), _p(X, C.c_double))
        return dict(pose_index=pi, point_index=li, err=err, chi2=chi[0], robust_chi2=chi[1], poses=P, points=X)

    def debug_system(self):
        """Normal equations of the active set of the last debug_phase (refba_debug_system): Hpp per pose, Hll per
        landmark, Hpl per edge (6x3), b per vertex."""
        Hpp, Hll = np.zeros((self.n_pose, 6, 6)), np.zeros((self.n_point, 3, 3))
        Hpl, b = np.zeros((self.n_obs, 6, 3)), np.zeros(self.n_pose * 6 + self.n_point * 3)
        lib().refba_debug_system(self.h, _p(Hpp, C.c_double), _p(Hll, C.c_double), _p(Hpl, C.c_double), _p(b, C.c_double))
        return dict(Hpp=Hpp, Hll=Hll, Hpl=Hpl, b_pose=b[:self.n_pose * 6].reshape(-1, 6),
                    b_point=b[self.n_pose * 6:].reshape(-1, 3))

    def set_lidar_edges(self, cur_pose, pc, qw, normal, w, n_flat, numeric_jacobian=True):
        """Explicit lidar correspondences for the third pass (flat edges first; w == 0 -> no edge)."""
        pc, qw, normal, w = (np.ascontiguousarray(a, np.float64) for a in (pc, qw, normal, w))
        n = len(w)
        lib().refba_set_lidar_edges(self.h, cur_pose, n_flat, n - n_flat, _p(pc, C.c_double), _p(qw, C.c_double),
                                    _p(normal, C.c_double), _p(w, C.c_double), int(numeric_jacobian))

    def set_lidar(self, ld, numeric_jacobian=True):
        """Clouds of the lidar pass (synth.LidarData); the association runs inside solve_local before the third pass."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        a = [f32(ld.flat_xyz), f32(ld.flat_normal), f32(ld.corner_xyz), f32(ld.map_flat_xyz), i32(ld.map_flat_pose),
             f32(ld.map_corner_xyz), i32(ld.map_corner_pose)]
        lib().refba_set_lidar(self.h, ld.cur_pose, len(a[0]), _p(a[0], C.c_float), _p(a[1], C.c_float), len(a[2]),
                              _p(a[2], C.c_float), len(a[3]), _p(a[3], C.c_float), _p(a[4], C.c_int32), len(a[5]),
                              _p(a[5], C.c_float), _p(a[6], C.c_int32), ld.distance_sq_threshold, ld.flat_weight,
                              ld.corner_weight, int(ld.use_flat), int(ld.use_corner), int(numeric_jacobian))

    def lidar_matches(self):
        n = lib().refba_lidar_num_matches(self.h)
        out = np.full(n, -1, np.int32)
        if n:
            lib().refba_get_lidar_matches(self.h, _p(out, C.c_int32))
        return out

    def num_lidar_edges(self) -> int:
        return lib().refba_num_lidar_edges(self.h)

    def solve_seconds(self) -> float:
        return lib().refba_solve_seconds(self.h)

    def poses(self):
        out = np.zeros((self.n_pose, 7))
        lib().refba_get_poses(self.h, _p(out, C.c_double))
        return out

    def points(self):
        out = np.zeros((self.n_point, 3))
        lib().refba_get_points(self.h, _p(out, C.c_double))
        return out

    def outliers(self):
        out = np.zeros(self.n_obs, np.uint8)
        lib().refba_get_outliers(self.h, _p(out, C.c_uint8))
        return out

    def trace(self):
        n = lib().refba_trace_len(self.h)
        out = np.zeros((n, 8))
        if n:
            lib().refba_get_trace(self.h, _p(out, C.c_double))
        return out

    def linearize_all(self, huber: int = 1):
        n = self.n_obs
        err, Jp, Jl, w, rho0 = np.zeros((n, 3)), np.zeros((n, 18)), np.zeros((n, 9)), np.zeros(n), np.zeros(n)
        lib().refba_linearize_all(self.h, huber, _p(err, C.c_double), _p(Jp, C.c_double), _p(Jl, C.c_double),
                                  _p(w, C.c_double), _p(rho0, C.c_double))
        return dict(err=err, Jp=Jp.reshape(n, 3, 6), Jl=Jl.reshape(n, 3, 3), w=w, rho0=rho0)

    def schur_solve(self, lam: float, huber: int = 1):
        npz, nl = self.n_pose, self.n_point
        S = np.zeros((6 * npz, 6 * npz))
        bs = np.zeros(6 * npz)
        b = np.zeros(6 * npz + 3 * nl)
        x = np.zeros(6 * npz + 3 * nl)
        sp = np.zeros(npz, np.int32)
        sl = np.zeros(nl, np.int32)
        md = C.c_double(0)
        # the C side writes S with leading dimension 6*Np, so hand it a scratch buffer and reshape afterwards
        Np = lib().refba_schur_solve(self.h, huber, lam, _p(S, C.c_double), _p(bs, C.c_double), _p(b, C.c_double),
                                     _p(x, C.c_double), _p(sp, C.c_int32), _p(sl, C.c_int32), C.byref(md))
        if Np < 0:
            raise RuntimeError("oracle LDLT failed")
        n = 6 * Np
        Sd = S.ravel()[: n * n].reshape(n, n).copy()
        Nl = int((self.n_point))
        # number of active landmarks = all (every landmark has >= 1 edge in generated problems)
        return dict(Np=Np, S=Sd, bschur=bs[:n].copy(), b=b[: n + 3 * Nl].copy(), x=x[: n + 3 * Nl].copy(),
                    slot_pose=sp[:Np].copy(), slot_point=sl[:Nl].copy(), max_diag=md.value)


def pose_oplus(pose7, upd6):
    a = np.ascontiguousarray(pose7, np.float64)
    u = np.ascontiguousarray(upd6, np.float64)
    out = np.zeros(7)
    lib().refba_pose_oplus(_p(a, C.c_double), _p(u, C.c_double), _p(out, C.c_double))
    return out


def se3_exp(upd6):
    u = np.ascontiguousarray(upd6, np.float64)
    out = np.zeros(7)
    lib().refba_se3_exp(_p(u, C.c_double), _p(out, C.c_double))
    return out


def pose_opt(pose7, cam5, xyz, meas):
    """g2oOptimizer::PoseOptimization restated (oracle/refba.cpp: refba_pose_opt).
    Returns (pose7_out, outlier flags, inliers, trace[rows, 8])."""
    L = lib()
    pose = np.ascontiguousarray(pose7, np.float64).copy()
    cam = np.ascontiguousarray(cam5, np.float64)
    xyz = np.ascontiguousarray(xyz, np.float64)
    meas = np.ascontiguousarray(meas, np.float32)
    n = xyz.shape[0]
    out = np.zeros(max(n, 1), np.uint8)
    tr = np.zeros((400, 8))
    nt = C.c_int32(0)
    inl = L.refba_pose_opt(_p(pose, C.c_double), _p(cam, C.c_double), n, _p(xyz, C.c_double), _p(meas, C.c_float),
                           _p(out, C.c_uint8), _p(tr, C.c_double), 400, C.byref(nt))
    return pose, out[:n], int(inl), tr[:nt.value]


# ---- essential-graph (Sim3 pose-graph) optimisation; 8-vectors are qx,qy,qz,qw,tx,ty,tz,s
def sim3_exp(upd7):
    u, out = np.ascontiguousarray(upd7, np.float64), np.zeros(8)
    lib().refba_sim3_exp(_p(u, C.c_double), _p(out, C.c_double))
    return out


def sim3_log(s8):
    a, out = np.ascontiguousarray(s8, np.float64), np.zeros(7)
    lib().refba_sim3_log(_p(a, C.c_double), _p(out, C.c_double))
    return out


def sim3_oplus(s8, upd7, fix_scale):
    a, u, out = np.ascontiguousarray(s8, np.float64), np.ascontiguousarray(upd7, np.float64), np.zeros(8)
    lib().refba_sim3_oplus(_p(a, C.c_double), _p(u, C.c_double), int(fix_scale), _p(out, C.c_double))
    return out


def sim3_edge_error(meas8, v1, v2):
    m, a, b, out = (np.ascontiguousarray(x, np.float64) for x in (meas8, v1, v2, np.zeros(7)))
    lib().refba_sim3_edge_error(_p(m, C.c_double), _p(a, C.c_double), _p(b, C.c_double), _p(out, C.c_double))
    return out


def pose_graph(vert8, fixed, fix_scale, edge_ij, meas8, iters=20, lambda_init=1e-16):
    """g2oOptimizer::OptimizeEssentialGraph's optimisation restated (refba_pose_graph).
    Returns (vertices n x 8, trace rows x 8, iterations performed)."""
    V = np.ascontiguousarray(vert8, np.float64).copy()
    fx = np.ascontiguousarray(fixed, np.uint8)
    E = np.ascontiguousarray(edge_ij, np.int32)
    M = np.ascontiguousarray(meas8, np.float64)
    tr = np.zeros((iters * 10 + 1, 8))
    nt = C.c_int32(0)
    done = lib().refba_pose_graph(len(V), _p(V, C.c_double), _p(fx, C.c_uint8), int(fix_scale), len(E), _p(E, C.c_int32),
                                  _p(M, C.c_double), iters, lambda_init, _p(tr, C.c_double), len(tr), C.byref(nt))
    return V, tr[:nt.value], done


def optimize_sim3(s12, cam8, p1c, p2c, meas6, th2=10.0, fix_scale=False):
    """g2oOptimizer::OptimizeSim3 restated (refba_optimize_sim3).
    Returns (s12_out 8, keep flags n, nIn, trace[rows, 8])."""
    L = lib()
    S = np.ascontiguousarray(s12, np.float64).copy()
    cam = np.ascontiguousarray(cam8, np.float64)
    a, b = np.ascontiguousarray(p1c, np.float64), np.ascontiguousarray(p2c, np.float64)
    m = np.ascontiguousarray(meas6, np.float32)
    n = a.shape[0]
    keep = np.zeros(max(n, 1), np.uint8)
    tr = np.zeros((150, 8))
    nt = C.c_int32(0)
    L.refba_optimize_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.refba_optimize_sim3.restype = C.c_int
    n_in = L.refba_optimize_sim3(S.ctypes.data, cam.ctypes.data, n, a.ctypes.data, b.ctypes.data, m.ctypes.data,
                                 C.c_float(th2), int(fix_scale), keep.ctypes.data, tr.ctypes.data, len(tr), C.byref(nt))
    return S, keep[:n], int(n_in), tr[:nt.value]


def sim3_match_linearize(s12, cam8, p1c, p2c, meas6, fix_scale):
    """(e12[n,2], e21[n,2], J12[n,2,7], J21[n,2,7]) of every match at s12 (refba_sim3_match_linearize)."""
    L = lib()
    S = np.ascontiguousarray(s12, np.float64)
    cam = np.ascontiguousarray(cam8, np.float64)
    a, b = np.ascontiguousarray(p1c, np.float64), np.ascontiguousarray(p2c, np.float64)
    m = np.ascontiguousarray(meas6, np.float32)
    n = a.shape[0]
    e12, e21, J12, J21 = np.zeros((n, 2)), np.zeros((n, 2)), np.zeros((n, 2, 7)), np.zeros((n, 2, 7))
    L.refba_sim3_match_linearize.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4
    L.refba_sim3_match_linearize.restype = None
    for k in range(n):
        L.refba_sim3_match_linearize(S.ctypes.data, cam.ctypes.data, a[k].ctypes.data, b[k].ctypes.data, m[k].ctypes.data,
                                     int(fix_scale), e12[k].ctypes.data, e21[k].ctypes.data, J12[k].ctypes.data, J21[k].ctypes.data)
    return e12, e21, J12, J21


class FrameLidarC(C.Structure):
    _fields_ = [("n_flat", C.c_int32), ("flat_xyz", C.c_void_p), ("flat_normal", C.c_void_p), ("n_corner", C.c_int32),
                ("corner_xyz", C.c_void_p), ("n_map", C.c_int64), ("map_xyz", C.c_void_p),
                ("distance_sq_threshold", C.c_double), ("flat_weight", C.c_double), ("corner_weight", C.c_double),
                ("use_flat", C.c_int32), ("use_corner", C.c_int32)]


def frame_lidar_struct(ld):
    """(ctypes struct, keep-alive list) for a synth.FrameLidar -- same layout as sqrtba_frame_lidar (include/sqrtba.h)."""
    keep = [np.ascontiguousarray(a, np.float32) for a in (ld.flat_xyz, ld.flat_normal, ld.corner_xyz, ld.map_xyz)]
    s = FrameLidarC(len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, len(keep[2]), keep[2].ctypes.data,
                    len(keep[3]), keep[3].ctypes.data, ld.distance_sq_threshold, ld.flat_weight, ld.corner_weight,
                    int(ld.use_flat), int(ld.use_corner))
    return s, keep


def pose_opt_lidar(pose7, cam5, xyz, meas, ld):
    """g2oOptimizer::PoseOptimization with this fork's lidar block restated (refba_pose_opt_lidar).
    Returns (pose7_out, outlier flags, inliers, trace[rows, 8], (flat matches, corner matches))."""
    L = lib()
    pose = np.ascontiguousarray(pose7, np.float64).copy()
    cam = np.ascontiguousarray(cam5, np.float64)
    xyz = np.ascontiguousarray(xyz, np.float64)
    meas = np.ascontiguousarray(meas, np.float32)
    n = xyz.shape[0]
    out = np.zeros(max(n, 1), np.uint8)
    tr = np.zeros((500, 8))
    nt = C.c_int32(0)
    nm = np.zeros(2, np.int32)
    st, keep = frame_lidar_struct(ld)
    L.refba_pose_opt_lidar.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
    L.refba_pose_opt_lidar.restype = C.c_int
    inl = L.refba_pose_opt_lidar(pose.ctypes.data, cam.ctypes.data, n, xyz.ctypes.data, meas.ctypes.data, out.ctypes.data,
                                 tr.ctypes.data, 500, C.addressof(nt), C.addressof(st), nm.ctypes.data)
    return pose, out[:n], int(inl), tr[:nt.value], (int(nm[0]), int(nm[1]))
