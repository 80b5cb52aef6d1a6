"""ctypes mirror of include/sqrtba.h -- the call a C++ host makes, exposed to Python tests and bench.py.

No compute happens here and there is NO fallback: if libsqrtba.so is missing the import of `lib()` raises,
and every solve goes through the C ABI to the CUDA device."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SQRTBA_LIB") or os.path.join(HERE, "libsqrtba.so")  # SQRTBA_LIB: instrumented builds
_lib = None


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("pcg_rtol", C.c_double), ("pcg_max_iters", C.c_int32),
                ("third_pass_iters", C.c_int32), ("pcg_mode", C.c_int32), ("pcg_check_every", C.c_int32),
                ("reserved", C.c_int32 * 8)]


class Stats(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("lm_trials", C.c_int32), ("cg_iters_total", C.c_int32),
                ("kernel_launches", C.c_int32), ("ms_total", C.c_double), ("ms_linearize", C.c_double),
                ("ms_qr", C.c_double), ("ms_pcg", C.c_double), ("ms_backsub", C.c_double), ("ms_cost", C.c_double),
                ("ms_matvec", C.c_double), ("reserved", C.c_double * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["persistent_pcg"] = int(self.reserved[0])
        d["peer_exchange"] = int(self.reserved[1])
        d["reproducible"] = int(self.reserved[5])
        d["pcg_rtol"] = float(self.reserved[4])
        d["chunk_precond"] = int(self.reserved[6])
        d["coarse_level"] = int(self.reserved[7])
        return d


class FrameLidar(C.Structure):
    """sqrtba_frame_lidar (include/sqrtba.h)"""
    _fields_ = [("n_flat", C.c_int32), ("flat_xyz", C.c_void_p), ("flat_normal", C.c_void_p), ("n_corner", C.c_int32),
                ("corner_xyz", C.c_void_p), ("n_map", C.c_int64), ("map_xyz", C.c_void_p),
                ("distance_sq_threshold", C.c_double), ("flat_weight", C.c_double), ("corner_weight", C.c_double),
                ("use_flat", C.c_int32), ("use_corner", C.c_int32)]


class Lidar(C.Structure):  # include/sqrtba.h: sqrtba_lidar
    _fields_ = [("cur_pose", C.c_int32), ("n_flat", C.c_int32), ("flat_xyz", C.POINTER(C.c_float)),
                ("flat_normal", C.POINTER(C.c_float)), ("n_corner", C.c_int32), ("numeric_jacobian", C.c_int32),
                ("corner_xyz", C.POINTER(C.c_float)), ("n_map_flat", C.c_int64), ("map_flat_xyz", C.POINTER(C.c_float)),
                ("map_flat_pose", C.POINTER(C.c_int32)), ("n_map_corner", C.c_int64),
                ("map_corner_xyz", C.POINTER(C.c_float)), ("map_corner_pose", C.POINTER(C.c_int32)),
                ("distance_sq_threshold", C.c_double), ("flat_weight", C.c_double), ("corner_weight", C.c_double),
                ("use_flat", C.c_int32), ("use_corner", C.c_int32)]


TRACE_COLS = ("pass", "iter", "trial", "lambda", "chi_before", "chi_trial", "rho", "accepted", "cg_iters", "cg_relres")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with sqrtlm-slam_b200/build.py (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, dp, fp, ip, up, lp = (C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_uint8), C.POINTER(C.c_int64))
        L.sqrtba_version.restype = C.c_char_p
        L.sqrtba_last_error.restype = C.c_char_p
        L.sqrtba_last_error.argtypes = [vp]
        L.sqrtba_default_config.argtypes = [C.POINTER(Config)]
        L.sqrtba_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.sqrtba_destroy.argtypes = [vp]
        L.sqrtba_set_problem.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, dp, up, dp, dp, ip, ip, fp]
        L.sqrtba_set_problem_batch.argtypes = [vp, C.c_int32, lp, lp, lp, dp, up, dp, dp, ip, ip, fp]
        L.sqrtba_reset_state.argtypes = [vp]
        L.sqrtba_solve_local.argtypes = [vp, vp, C.POINTER(Stats)]
        L.sqrtba_solve_global.argtypes = [vp, C.c_int32, C.c_int32, vp, C.POINTER(Stats)]
        L.sqrtba_get_poses.argtypes = [vp, dp]
        L.sqrtba_get_points.argtypes = [vp, dp]
        L.sqrtba_get_outliers.argtypes = [vp, up]
        L.sqrtba_get_trace_len.argtypes = [vp, C.c_int32]
        L.sqrtba_get_trace.argtypes = [vp, C.c_int32, dp, C.c_int32]
        L.sqrtba_debug_linearize.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp]
        L.sqrtba_debug_step.argtypes = [vp, C.c_double, dp, dp, dp, ip]
        L.sqrtba_debug_matvec.argtypes = [vp, dp, dp]
        L.sqrtba_num_free_poses.argtypes = [vp]
        L.sqrtba_debug_plan.argtypes = [C.c_int32, lp, lp, lp, C.c_int32, C.c_int32, C.c_int32, up, ip, ip, C.c_int32, ip, ip,
                                        C.c_int32, C.POINTER(C.c_uint32), ip, ip, C.c_int64, ip]
        L.sqrtba_time_stage.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, dp]
        L.sqrtba_pose_opt.argtypes = [vp, C.c_int32, lp, dp, dp, dp, fp, up, ip, C.POINTER(Stats)]
        L.sqrtba_pose_opt_trace.argtypes = [vp, C.c_int32, dp, C.c_int32]
        L.sqrtba_pose_opt_lidar.argtypes = [vp, dp, dp, C.c_int32, dp, fp, up, ip, C.c_void_p, ip, C.POINTER(Stats)]
        L.sqrtba_pose_graph.argtypes = [vp, C.c_int32, dp, up, C.c_int32, C.c_int32, ip, dp, C.c_int32, C.c_double, C.POINTER(Stats)]
        L.sqrtba_pose_graph_trace.argtypes = [vp, dp, C.c_int32]
        L.sqrtba_optimize_sim3.argtypes = [vp, C.c_int32, lp, dp, dp, dp, dp, fp, C.c_float, C.c_int32, up, ip, C.POINTER(Stats)]
        L.sqrtba_optimize_sim3_trace.argtypes = [vp, C.c_int32, dp, C.c_int32]
        L.sqrtba_set_lidar.argtypes = [vp, C.POINTER(Lidar)]
        L.sqrtba_set_lidar_edges.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp, dp, C.c_int32]
        L.sqrtba_get_lidar_matches.argtypes = [vp, ip]
        L.sqrtba_num_lidar_edges.argtypes = [vp]
        L.sqrtba_comm_unique_id.argtypes = [up]
        L.sqrtba_comm_init.argtypes = [vp, C.c_int32, C.c_int32, up]
        L.sqrtba_comm_destroy.argtypes = [vp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class SqrtBAError(RuntimeError):
    pass


class SqrtBA:
    """Thin owner of one sqrtba_handle.  Method names follow the C ABI."""

    def __init__(self, device: int = 0, pcg_rtol: float = 0.0, pcg_max_iters: int = 2000, third_pass_iters: int = 0,
                 pcg_mode: int = 0, pcg_check_every: int = 4, stage_timing: bool = False,
                 general_matvec: bool = False, pipe_stages: int = 0, host_threads: int = 0,
                 no_reorder: bool = False, pipe_slots: int = 0, plain_qr: bool = False, qr_variant: int = 0):
        L = lib()
        cfg = Config()
        L.sqrtba_default_config(C.byref(cfg))
        cfg.device, cfg.pcg_rtol, cfg.pcg_max_iters = device, pcg_rtol, pcg_max_iters
        cfg.third_pass_iters, cfg.pcg_mode, cfg.pcg_check_every = third_pass_iters, pcg_mode, pcg_check_every
        cfg.reserved[0] = 1 if stage_timing else 0
        cfg.reserved[1] = 1 if general_matvec else 0   # force the non-pipelined tile kernel (A/B profiling)
        cfg.reserved[2] = pipe_stages                  # 0 = auto, 2/3 = forced TMA ring depth
        cfg.reserved[3] = host_threads                 # 0 = all host cores for set_problem's preprocessing
        cfg.reserved[4] = 1 if no_reorder else 0       # keep the caller's landmark order in big windows (A/B profiling)
        cfg.reserved[5] = pipe_slots                   # big windows: slots in the matvec's shared accumulator window
        cfg.reserved[7] = 1 if plain_qr else qr_variant  # 1: one-tile-per-CTA kernels; 2-6: A/B variants; 7: fused linearise + QR
        self.h = C.c_void_p()
        rc = L.sqrtba_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            msg = L.sqrtba_last_error(None)
            self.h = None
            raise SqrtBAError(f"sqrtba_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n_pose = self.n_point = self.n_obs = 0
        self.n_win = 1

    def close(self):
        if getattr(self, "h", None):
            lib().sqrtba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc < 0:
            msg = lib().sqrtba_last_error(self.h)
            raise SqrtBAError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    def _arrays(self, prob):
        return (np.ascontiguousarray(prob.pose_qt, np.float64), np.ascontiguousarray(prob.pose_fixed, np.uint8),
                np.ascontiguousarray(prob.cam, np.float64), np.ascontiguousarray(prob.point_xyz, np.float64),
                np.ascontiguousarray(prob.obs_pose, np.int32), np.ascontiguousarray(prob.obs_point, np.int32),
                np.ascontiguousarray(prob.obs_meas, np.float32))

    def set_problem(self, prob):
        a = self._arrays(prob)
        self.n_pose, self.n_point, self.n_obs, self.n_win = prob.n_pose, prob.n_point, prob.n_obs, 1
        self._chk(lib().sqrtba_set_problem(self.h, prob.n_pose, prob.n_point, prob.n_obs, _p(a[0], C.c_double),
                                           _p(a[1], C.c_uint8), _p(a[2], C.c_double), _p(a[3], C.c_double),
                                           _p(a[4], C.c_int32), _p(a[5], C.c_int32), _p(a[6], C.c_float)),
                  "sqrtba_set_problem")

    def set_problem_batch(self, prob, pose_ptr, point_ptr, obs_ptr):
        a = self._arrays(prob)
        pp, tp, op = (np.ascontiguousarray(x, np.int64) for x in (pose_ptr, point_ptr, obs_ptr))
        self.n_pose, self.n_point, self.n_obs, self.n_win = prob.n_pose, prob.n_point, prob.n_obs, len(pp) - 1
        self._chk(lib().sqrtba_set_problem_batch(self.h, self.n_win, _p(pp, C.c_int64), _p(tp, C.c_int64),
                                                 _p(op, C.c_int64), _p(a[0], C.c_double), _p(a[1], C.c_uint8),
                                                 _p(a[2], C.c_double), _p(a[3], C.c_double), _p(a[4], C.c_int32),
                                                 _p(a[5], C.c_int32), _p(a[6], C.c_float)), "sqrtba_set_problem_batch")

    def set_lidar(self, ld, numeric_jacobian: bool = True):
        """Clouds of the lidar tight-coupling pass (synth.LidarData layout); needs third_pass_iters > 0 to take part."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        a = [f32(ld.flat_xyz), f32(ld.flat_normal), f32(ld.corner_xyz), f32(ld.map_flat_xyz), i32(ld.map_flat_pose),
             f32(ld.map_corner_xyz), i32(ld.map_corner_pose)]
        c = Lidar(cur_pose=ld.cur_pose, n_flat=len(a[0]), flat_xyz=_p(a[0], C.c_float), flat_normal=_p(a[1], C.c_float),
                  n_corner=len(a[2]), numeric_jacobian=int(numeric_jacobian), corner_xyz=_p(a[2], C.c_float),
                  n_map_flat=len(a[3]), map_flat_xyz=_p(a[3], C.c_float), map_flat_pose=_p(a[4], C.c_int32),
                  n_map_corner=len(a[5]), map_corner_xyz=_p(a[5], C.c_float), map_corner_pose=_p(a[6], C.c_int32),
                  distance_sq_threshold=ld.distance_sq_threshold, flat_weight=ld.flat_weight,
                  corner_weight=ld.corner_weight, use_flat=int(ld.use_flat), use_corner=int(ld.use_corner))
        self._chk(lib().sqrtba_set_lidar(self.h, C.byref(c)), "sqrtba_set_lidar")
        self._n_lidar = len(a[0]) + len(a[2])

    def set_lidar_edges(self, cur_pose, pc, qw, normal, w, n_flat, numeric_jacobian: bool = True):
        pc, qw, normal, w = (np.ascontiguousarray(a, np.float64) for a in (pc, qw, normal, w))
        self._chk(lib().sqrtba_set_lidar_edges(self.h, cur_pose, n_flat, len(w) - n_flat, _p(pc, C.c_double),
                                               _p(qw, C.c_double), _p(normal, C.c_double), _p(w, C.c_double),
                                               int(numeric_jacobian)), "sqrtba_set_lidar_edges")

    def lidar_matches(self) -> np.ndarray:
        out = np.full(max(getattr(self, "_n_lidar", 0), 1), -1, np.int32)
        n = self._chk(lib().sqrtba_get_lidar_matches(self.h, _p(out, C.c_int32)), "sqrtba_get_lidar_matches")
        return out[:n]

    def num_lidar_edges(self) -> int:
        return self._chk(lib().sqrtba_num_lidar_edges(self.h), "sqrtba_num_lidar_edges")

    def reset_state(self):
        self._chk(lib().sqrtba_reset_state(self.h), "sqrtba_reset_state")

    def solve_local(self, stop=None) -> dict:
        st = Stats()
        self._chk(lib().sqrtba_solve_local(self.h, stop, C.byref(st)), "sqrtba_solve_local")
        return st.as_dict()

    def solve_global(self, iters: int, robust: bool, stop=None) -> dict:
        st = Stats()
        self._chk(lib().sqrtba_solve_global(self.h, iters, int(robust), stop, C.byref(st)), "sqrtba_solve_global")
        return st.as_dict()

    # the getters copy into a caller buffer (`out`: C-contiguous, right dtype and size; pinned host memory makes the
    # device-to-host copy a plain DMA) or allocate a fresh pageable array
    def poses(self, out=None):
        if out is None:
            out = np.empty((self.n_pose, 7))
        assert out.dtype == np.float64 and out.size == self.n_pose * 7 and out.flags.c_contiguous
        self._chk(lib().sqrtba_get_poses(self.h, _p(out, C.c_double)), "sqrtba_get_poses")
        return out

    def points(self, out=None):
        if out is None:
            out = np.empty((self.n_point, 3))
        assert out.dtype == np.float64 and out.size == self.n_point * 3 and out.flags.c_contiguous
        self._chk(lib().sqrtba_get_points(self.h, _p(out, C.c_double)), "sqrtba_get_points")
        return out

    def outliers(self, out=None):
        if out is None:
            out = np.empty(self.n_obs, np.uint8)
        assert out.dtype == np.uint8 and out.size == self.n_obs and out.flags.c_contiguous
        self._chk(lib().sqrtba_get_outliers(self.h, _p(out, C.c_uint8)), "sqrtba_get_outliers")
        return out

    def trace(self, window: int = 0):
        n = self._chk(lib().sqrtba_get_trace_len(self.h, window), "sqrtba_get_trace_len")
        out = np.zeros((n, len(TRACE_COLS)))
        if n:
            self._chk(lib().sqrtba_get_trace(self.h, window, _p(out, C.c_double), n), "sqrtba_get_trace")
        return out

    def num_free_poses(self) -> int:
        return self._chk(lib().sqrtba_num_free_poses(self.h), "sqrtba_num_free_poses")

    def debug_linearize(self, huber: int = 1):
        n = self.n_obs
        err, Jp, Jl, r = np.zeros((n, 3)), np.zeros((n, 18)), np.zeros((n, 9)), np.zeros((n, 3))
        chi = np.zeros(self.n_win)
        self._chk(lib().sqrtba_debug_linearize(self.h, huber, _p(err, C.c_double), _p(Jp, C.c_double),
                                               _p(Jl, C.c_double), _p(r, C.c_double), _p(chi, C.c_double)),
                  "sqrtba_debug_linearize")
        return dict(err=err, Jp=Jp.reshape(n, 3, 6), Jl=Jl.reshape(n, 3, 3), r=r, chi2=chi)

    def debug_step(self, lam: float):
        ns = self.num_free_poses()
        dp, dl, bs = np.zeros((max(ns, 1), 6)), np.zeros((self.n_point, 3)), np.zeros((max(ns, 1), 6))
        it = C.c_int32(0)
        self._chk(lib().sqrtba_debug_step(self.h, lam, _p(dp, C.c_double), _p(dl, C.c_double), _p(bs, C.c_double),
                                          C.byref(it)), "sqrtba_debug_step")
        return dict(dp=dp[:ns], dl=dl, bs=bs[:ns], cg_iters=it.value)

    def debug_matvec(self, p):
        p = np.ascontiguousarray(p, np.float64)
        y = np.zeros_like(p)
        self._chk(lib().sqrtba_debug_matvec(self.h, _p(p, C.c_double), _p(y, C.c_double)), "sqrtba_debug_matvec")
        return y

    def pose_opt(self, frame_ptr, pose_qt, cam, obs_xyz, obs_meas):
        """sqrtba_pose_opt on a batch of frames.  Returns (poses, outlier flags, inliers per frame, stats)."""
        fp_ = np.ascontiguousarray(frame_ptr, np.int64)
        nf = len(fp_) - 1
        pose = np.ascontiguousarray(pose_qt, np.float64).reshape(nf, 7).copy()
        cam = np.ascontiguousarray(cam, np.float64).reshape(nf, 5)
        xyz = np.ascontiguousarray(obs_xyz, np.float64).reshape(-1, 3)
        meas = np.ascontiguousarray(obs_meas, np.float32).reshape(-1, 4)
        out = np.zeros(max(int(fp_[-1]), 1), np.uint8)
        inl = np.zeros(nf, np.int32)
        st = Stats()
        self._chk(lib().sqrtba_pose_opt(self.h, nf, _p(fp_, C.c_int64), _p(pose, C.c_double), _p(cam, C.c_double),
                                        _p(xyz, C.c_double), _p(meas, C.c_float), _p(out, C.c_uint8), _p(inl, C.c_int32),
                                        C.byref(st)), "sqrtba_pose_opt")
        return pose, out[:int(fp_[-1])], inl, st.as_dict()

    def pose_opt_trace(self, frame: int = 0):
        rows = np.zeros((500, 8))
        n = self._chk(lib().sqrtba_pose_opt_trace(self.h, frame, _p(rows, C.c_double), 500), "sqrtba_pose_opt_trace")
        return rows[:n]

    def pose_opt_lidar(self, pose_qt, cam, obs_xyz, obs_meas, ld):
        """sqrtba_pose_opt_lidar on one frame (ld: synth.FrameLidar).  Returns (pose 7, outlier flags, inliers,
        (flat matches, corner matches), stats)."""
        pose = np.ascontiguousarray(pose_qt, np.float64).reshape(7).copy()
        cam = np.ascontiguousarray(cam, np.float64).reshape(5)
        xyz = np.ascontiguousarray(obs_xyz, np.float64).reshape(-1, 3)
        meas = np.ascontiguousarray(obs_meas, np.float32).reshape(-1, 4)
        keep = [np.ascontiguousarray(a, np.float32) for a in (ld.flat_xyz, ld.flat_normal, ld.corner_xyz, ld.map_xyz)]
        fl = FrameLidar(len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, len(keep[2]), keep[2].ctypes.data,
                        len(keep[3]), keep[3].ctypes.data, ld.distance_sq_threshold, ld.flat_weight, ld.corner_weight,
                        int(ld.use_flat), int(ld.use_corner))
        out = np.zeros(max(len(xyz), 1), np.uint8)
        inl, nm = np.zeros(1, np.int32), np.zeros(2, np.int32)
        st = Stats()
        self._chk(lib().sqrtba_pose_opt_lidar(self.h, _p(pose, C.c_double), _p(cam, C.c_double), len(xyz), _p(xyz, C.c_double),
                                              _p(meas, C.c_float), _p(out, C.c_uint8), _p(inl, C.c_int32), C.addressof(fl),
                                              _p(nm, C.c_int32), C.byref(st)), "sqrtba_pose_opt_lidar")
        return pose, out[:len(xyz)], int(inl[0]), (int(nm[0]), int(nm[1])), st.as_dict()

    def pose_graph(self, vert8, fixed, fix_scale, edge_ij, meas8, iters: int = 20, lambda_init: float = 1e-16):
        """sqrtba_pose_graph: returns (vertices n x 8 after the optimisation, LM trace rows x 8, stats)."""
        V = np.ascontiguousarray(vert8, np.float64).copy()
        fx = np.ascontiguousarray(fixed, np.uint8)
        E = np.ascontiguousarray(edge_ij, np.int32)
        M = np.ascontiguousarray(meas8, np.float64)
        st = Stats()
        self._chk(lib().sqrtba_pose_graph(self.h, len(V), _p(V, C.c_double), _p(fx, C.c_uint8), int(fix_scale), len(E),
                                          _p(E, C.c_int32), _p(M, C.c_double), iters, lambda_init, C.byref(st)), "sqrtba_pose_graph")
        n = lib().sqrtba_pose_graph_trace(self.h, None, 0)
        tr = np.zeros((max(n, 0), 8))
        if n > 0:
            lib().sqrtba_pose_graph_trace(self.h, _p(tr, C.c_double), n)
        return V, tr, st.as_dict()

    def optimize_sim3(self, pair_ptr, s12, cam8, p1c, p2c, meas6, th2: float = 10.0, fix_scale: bool = False):
        """sqrtba_optimize_sim3 over a batch of keyframe pairs: returns (s12 n_pairs x 8, keep flags per match,
        nIn per pair, stats)."""
        ptr = np.ascontiguousarray(pair_ptr, np.int64)
        S = np.ascontiguousarray(s12, np.float64).reshape(-1, 8).copy()
        cam = np.ascontiguousarray(cam8, np.float64).reshape(-1, 8)
        a, b = np.ascontiguousarray(p1c, np.float64), np.ascontiguousarray(p2c, np.float64)
        m = np.ascontiguousarray(meas6, np.float32)
        n_pairs = len(ptr) - 1
        keep = np.zeros(max(int(ptr[-1]), 1), np.uint8)
        n_in = np.zeros(n_pairs, np.int32)
        st = Stats()
        self._chk(lib().sqrtba_optimize_sim3(self.h, n_pairs, _p(ptr, C.c_int64), _p(S, C.c_double), _p(cam, C.c_double),
                                             _p(a, C.c_double), _p(b, C.c_double), _p(m, C.c_float), C.c_float(th2),
                                             int(fix_scale), _p(keep, C.c_uint8), _p(n_in, C.c_int32), C.byref(st)),
                  "sqrtba_optimize_sim3")
        return S, keep[:int(ptr[-1])], n_in, st.as_dict()

    def optimize_sim3_trace(self, pair: int) -> np.ndarray:
        rows = np.zeros((150, 8))
        n = self._chk(lib().sqrtba_optimize_sim3_trace(self.h, pair, _p(rows, C.c_double), 150), "sqrtba_optimize_sim3_trace")
        return rows[:n]

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._chk(lib().sqrtba_comm_init(self.h, nranks, rank, buf), "sqrtba_comm_init")

    def comm_destroy(self):
        self._chk(lib().sqrtba_comm_destroy(self.h), "sqrtba_comm_destroy")

    def time_stage(self, stage: int, warmup: int = 3, reps: int = 20) -> float:
        ms = C.c_double(0)
        self._chk(lib().sqrtba_time_stage(self.h, stage, warmup, reps, C.byref(ms)), "sqrtba_time_stage")
        return ms.value


TILE_COLS = ("item0", "nitem", "o0", "o1", "win", "nfree", "nt", "is_long", "nrun",
             "cnt0", "cnt1", "cnt2", "cnt3", "fcnt0", "fcnt1", "fcnt2", "fcnt3", "blk_doubles", "jq_off_lo", "jq_off_hi")


def debug_plan(prob, pose_ptr=None, point_ptr=None, obs_ptr=None, host_threads: int = 0) -> dict:
    """Host-side plan of set_problem[_batch] (items, tiles, run tables, landmark order) -- needs no GPU.

    Two calls: the first sizes the outputs from the summary, the second fills them.
    """
    if pose_ptr is None:
        pose_ptr, point_ptr, obs_ptr = [0, prob.n_pose], [0, prob.n_point], [0, prob.n_obs]
    pp, tp, op = (np.ascontiguousarray(x, np.int64) for x in (pose_ptr, point_ptr, obs_ptr))
    fx = np.ascontiguousarray(prob.pose_fixed, np.uint8)
    opose = np.ascontiguousarray(prob.obs_pose, np.int32)
    opoint = np.ascontiguousarray(prob.obs_point, np.int32)
    n_win = len(pp) - 1
    summ = np.zeros(8, np.int32)

    def call(tiles, lp_, rptr, runs, perm):
        rc = lib().sqrtba_debug_plan(
            n_win, _p(pp, C.c_int64), _p(tp, C.c_int64), _p(op, C.c_int64), prob.n_pose, prob.n_point, prob.n_obs,
            _p(fx, C.c_uint8), _p(opose, C.c_int32), _p(opoint, C.c_int32), host_threads, _p(summ, C.c_int32),
            None if tiles is None else _p(tiles, C.c_int32), 0 if tiles is None else len(tiles),
            None if lp_ is None else _p(lp_, C.c_uint32), None if rptr is None else _p(rptr, C.c_int32),
            None if runs is None else _p(runs, C.c_int32), 0 if runs is None else len(runs),
            None if perm is None else _p(perm, C.c_int32))
        if rc != 0:
            raise SqrtBAError(f"sqrtba_debug_plan failed with code {rc}")

    call(None, None, None, None, None)
    n_item, n_tile, n_runs = int(summ[0]), int(summ[1]), int(summ[2])
    tiles = np.zeros((n_tile, 20), np.int32)
    obs_lp = np.zeros(prob.n_obs, np.uint32)
    run_ptr = np.zeros(n_tile + 1, np.int32)
    runs = np.zeros(max(n_runs, 1), np.int32)
    perm = np.zeros(prob.n_point, np.int32)
    call(tiles, obs_lp, run_ptr, runs, perm)
    return dict(n_item=n_item, n_tile=n_tile, n_run_ints=n_runs, smallwin=bool(summ[3]), pq_shared=bool(summ[4]),
                reordered=bool(summ[5]), jq_doubles=int(summ[6]) + (int(summ[7]) << 31), tiles=tiles, obs_lp=obs_lp,
                tile_run_ptr=run_ptr, tile_runs=runs[:n_runs], landmark_order=perm)
