"""ctypes wrapper of host/harness.cc: drives the C++ adapter (reference Optimizer API) from Python tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build, synth

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build.build_host())
        fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.hh_build.restype = C.c_void_p
        L.hh_build.argtypes = [C.c_int, fp, fp, fp, C.c_int, C.c_int, fp, C.c_int, ip, ip, fp, ip]
        L.hh_destroy.argtypes = [C.c_void_p]
        L.hh_set_covisible.argtypes = [C.c_void_p, C.c_int, ip, C.c_int]
        L.hh_set_bad.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_set_lidar_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp, C.c_int, fp]
        L.hh_set_lidar_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.hh_gather.argtypes = [C.c_void_p, C.c_int, ip]
        L.hh_gather_get.argtypes = [C.POINTER(C.c_double), up, C.POINTER(C.c_double), C.POINTER(C.c_double), ip, ip, fp,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.hh_apply_local.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                     C.c_int, up]
        L.hh_local_ba.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hh_set_options.argtypes = [C.c_int, C.c_int]
        dpp = C.POINTER(C.c_double)
        L.hh_forget_gba.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_apply_gba.argtypes = [C.c_void_p, C.c_ulong]
        L.hh_set_origin.argtypes = [C.c_void_p, C.c_int]
        L.hh_set_parent.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_add_loop_edge.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_set_weight.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hh_set_ref_kf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_long]
        L.hh_essential_graph.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip, dpp, dpp, C.c_int, ip, C.c_int, ip]
        L.hh_essential_graph_get.argtypes = [dpp, up, up, ip, dpp]
        L.hh_global_ba.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_ulong, C.c_void_p]
        L.hh_set_K.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.hh_gather_sim3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ip, dpp, dpp, dpp, fp, ip]
        L.hh_optimize_sim3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ip, dpp, C.c_float, C.c_int]
        L.hh_get_pose.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.hh_get_point.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.hh_has_observation.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_keyframe_sees.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_point_updates.argtypes = [C.c_void_p, C.c_int]
        L.hh_gba_marker.restype = C.c_ulong
        L.hh_gba_marker.argtypes = [C.c_void_p, C.c_int]
        L.hh_last_error.restype = C.c_char_p
        L.hh_frame_build.restype = C.c_void_p
        L.hh_frame_build.argtypes = [fp, fp, fp, C.c_int, C.c_int, fp, fp, ip]
        L.hh_frame_destroy.argtypes = [C.c_void_p]
        L.hh_frame_unmatch.argtypes = [C.c_void_p, C.c_int]
        L.hh_pose_opt.argtypes = [C.c_void_p]
        L.hh_pose_opt_lidar.argtypes = [C.c_void_p, C.c_int, fp, fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_double]
        L.hh_pose_opt_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, ip]
        L.hh_frame_get.argtypes = [C.c_void_p, fp, up]
        L.hh_mirror_attach.argtypes = [C.c_void_p]
        L.hh_reset_markers.argtypes = [C.c_void_p]
        L.hh_add_observation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
        L.hh_erase_observation.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_set_point_bad.argtypes = [C.c_void_p, C.c_int]
        L.hh_mirror_mismatches.argtypes = [C.c_void_p]
        L.hh_mirror_stress.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hh_time_gather.restype = C.c_double
        L.hh_time_gather.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def set_options(stereo_edges: bool, two_pass: bool):
    """sqrtbaOptimizer::options() (host/Optimizer.h).  The C++ defaults are the reference fork's behaviour (monocular
    edges only, always a third optimize(20)); the methods below default to upstream ORB-SLAM2's (stereo edges, two
    passes -- the schedule BASELINE.json's configs name) and say so at every call."""
    lib().hh_set_options(int(stereo_edges), int(two_pass))


class MockMap:
    """A map of header-compatible KeyFrame/MapPoint objects built from a synthetic Problem (keyframe i has mnId i)."""

    def __init__(self, prob):
        self.prob = prob
        R = synth.quat_to_rotmat(prob.pose_qt[:, 3:])
        T = np.tile(np.eye(4, dtype=np.float32), (prob.n_pose, 1, 1))
        T[:, :3, :3] = R.astype(np.float32)
        T[:, :3, 3] = prob.pose_qt[:, :3].astype(np.float32)
        inv = synth.inv_level_sigma2()
        octave = np.array([int(np.argmin(np.abs(inv - s))) for s in prob.obs_meas[:, 3]], np.int32)
        assert np.all(inv[octave] == prob.obs_meas[:, 3])
        self._keep = [np.ascontiguousarray(T), np.ascontiguousarray(prob.cam[0], np.float32), inv,
                      np.ascontiguousarray(prob.point_xyz, np.float32), np.ascontiguousarray(prob.obs_pose, np.int32),
                      np.ascontiguousarray(prob.obs_point, np.int32), np.ascontiguousarray(prob.obs_meas[:, :3], np.float32),
                      octave]
        k = self._keep
        self.h = lib().hh_build(prob.n_pose, _f(k[0]), _f(k[1]), _f(k[2]), len(inv), prob.n_point, _f(k[3]), prob.n_obs,
                                _i(k[4]), _i(k[5]), _f(k[6]), _i(k[7]))

    @classmethod
    def from_poses(cls, Tcw, points):
        """A map of keyframes (float 4x4 Tcw each, mnId = index) and map points without observations: what the
        essential-graph adapter needs (spanning tree / loop edges / weights are set with the methods below)."""
        self = cls.__new__(cls)
        self.prob = None
        T = np.ascontiguousarray(Tcw, np.float32)
        X = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        cam = np.array([synth.FX, synth.FY, synth.CX, synth.CY, synth.BF], np.float32)
        inv = synth.inv_level_sigma2()
        e_i, e_f = np.zeros(1, np.int32), np.zeros(3, np.float32)
        self._keep = [T, cam, inv, X, e_i, e_f]
        self.h = lib().hh_build(len(T), _f(T), _f(cam), _f(inv), len(inv), len(X), _f(X), 0, _i(e_i), _i(e_i), _f(e_f), _i(e_i))
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().hh_destroy(self.h)
            self.h = None

    # ---- the caller's use of the staged global-BA result (LoopClosing::RunGlobalBundleAdjustment, LoopClosing.cc:1014-1103)
    def set_origin(self, kf):
        lib().hh_set_origin(self.h, kf)

    def forget_gba(self, kf=-1, mp=-1):
        """Reset the global-BA marker of a keyframe / point: as if it was created while the optimisation ran."""
        lib().hh_forget_gba(self.h, kf, mp)

    def apply_gba(self, n_loop_kf):
        lib().hh_apply_gba(self.h, n_loop_kf)

    # ---- essential graph (Optimizer::OptimizeEssentialGraph)
    def set_parent(self, kf, parent):
        lib().hh_set_parent(self.h, kf, parent)

    def add_loop_edge(self, a, b):
        lib().hh_add_loop_edge(self.h, a, b)

    def set_weight(self, a, b, w):
        lib().hh_set_weight(self.h, a, b, w)

    def set_ref_kf(self, mp, kf, corrected_by=-1, corrected_ref=0):
        lib().hh_set_ref_kf(self.h, mp, kf, corrected_by, corrected_ref)

    def essential_graph(self, loop_kf, cur_kf, corrected, non_corrected, connections, fix_scale, optimise):
        """corrected / non_corrected: {kf: Sim3 as 8 doubles} (LoopClosing::KeyFrameAndPose, same keys); connections:
        [(kf, kf)] (LoopConnections).  optimise=False returns the graph the adapter would hand to sqrtba_pose_graph
        (no GPU needed); optimise=True runs Optimizer::OptimizeEssentialGraph on the map."""
        L = lib()
        ids = np.array(sorted(corrected), np.int32)
        c8 = np.ascontiguousarray([corrected[int(k)] for k in ids], np.float64).reshape(-1, 8)
        n8 = np.ascontiguousarray([non_corrected[int(k)] for k in ids], np.float64).reshape(-1, 8)
        conn = np.ascontiguousarray(connections, np.int32).reshape(-1, 2)
        sz = np.zeros(3, np.int32)
        dp = C.POINTER(C.c_double)
        L.hh_essential_graph(self.h, 1 if optimise else 0, loop_kf, cur_kf, len(ids), _i(ids), c8.ctypes.data_as(dp),
                             n8.ctypes.data_as(dp), len(conn), _i(conn), int(fix_scale), _i(sz))
        if optimise:
            return None
        nv, ne = int(sz[0]), int(sz[1])
        out = dict(vert8=np.zeros((nv, 8)), fixed=np.zeros(nv, np.uint8), present=np.zeros(nv, np.uint8),
                   edge_ij=np.zeros((ne, 2), np.int32), meas8=np.zeros((ne, 8)))
        up = C.POINTER(C.c_uint8)
        L.hh_essential_graph_get(out["vert8"].ctypes.data_as(dp), out["fixed"].ctypes.data_as(up), out["present"].ctypes.data_as(up),
                                 _i(out["edge_ij"]), out["meas8"].ctypes.data_as(dp))
        return out

    # ---- Sim3 of a loop candidate (Optimizer::OptimizeSim3)
    @classmethod
    def sim3_candidates(cls, case, seed=0, cam2=None):
        """Two keyframes and two sets of map points out of a synth.sim3_pair case: keyframe 0 sees map points 0..n-1
        at its keypoints 0..n-1, keyframe 1 sees map points n..2n-1 at SHUFFLED keypoints; map point n+i is the match
        of keypoint i of keyframe 0.  Returns (map, match_mp)."""
        s0, cam8, p1c, p2c, meas, _ = case
        rng = np.random.default_rng(seed)
        n = len(p1c)
        T = np.tile(np.eye(4), (2, 1, 1))
        for k in range(2):
            T[k, :3, :3] = synth.so3_exp(rng.normal(0, 0.4, 3))
            T[k, :3, 3] = rng.normal(0, 5.0, 3)
        T = T.astype(np.float32)
        Xw = np.concatenate([(p1c - T[0, :3, 3].astype(float)) @ T[0, :3, :3].astype(float),
                             (p2c - T[1, :3, 3].astype(float)) @ T[1, :3, :3].astype(float)]).astype(np.float32)
        inv = synth.inv_level_sigma2()
        oct1 = np.array([int(np.argmin(np.abs(inv - v))) for v in meas[:, 2]], np.int32)
        oct2 = np.array([int(np.argmin(np.abs(inv - v))) for v in meas[:, 5]], np.int32)
        order = rng.permutation(n)                                   # keypoint j of keyframe 1 belongs to map point n + order[j]
        obs_kf = np.concatenate([np.zeros(n, np.int32), np.ones(n, np.int32)])
        obs_mp = np.concatenate([np.arange(n), n + order]).astype(np.int32)
        uvr = np.zeros((2 * n, 3), np.float32)
        uvr[:n, :2], uvr[n:, :2], uvr[:, 2] = meas[:, 0:2], meas[order, 3:5], -1.0
        octave = np.concatenate([oct1, oct2[order]]).astype(np.int32)
        cam = np.array([cam8[0], cam8[1], cam8[2], cam8[3], synth.BF], np.float32)
        self = cls.__new__(cls)
        self.prob = None
        self._keep = [np.ascontiguousarray(T), cam, inv, np.ascontiguousarray(Xw), obs_kf, obs_mp, np.ascontiguousarray(uvr), octave]
        k = self._keep
        self.h = lib().hh_build(2, _f(k[0]), _f(k[1]), _f(k[2]), len(inv), 2 * n, _f(k[3]), 2 * n, _i(k[4]), _i(k[5]), _f(k[6]), _i(k[7]))
        if cam2 is not None:
            lib().hh_set_K(self.h, 1, *[float(v) for v in cam2])
        self.Tcw, self.Xw = T, Xw
        return self, (n + np.arange(n)).astype(np.int32)

    def gather_sim3(self, kf1, kf2, match_mp):
        """The arrays OptimizeSim3 would hand to sqrtba_optimize_sim3 (no GPU needed)."""
        mm = np.ascontiguousarray(match_mp, np.int32)
        n = len(mm)
        cam8, p1c, p2c = np.zeros(8), np.zeros((n, 3)), np.zeros((n, 3))
        meas, index = np.zeros((n, 6), np.float32), np.zeros(n, np.int32)
        dp = C.POINTER(C.c_double)
        m = lib().hh_gather_sim3(self.h, kf1, kf2, n, _i(mm), cam8.ctypes.data_as(dp), p1c.ctypes.data_as(dp),
                                 p2c.ctypes.data_as(dp), _f(meas), _i(index))
        return dict(cam8=cam8, p1c=p1c[:m], p2c=p2c[:m], meas6=meas[:m], index=index[:m])

    def optimize_sim3(self, kf1, kf2, match_mp, s12, th2=10.0, fix_scale=False):
        """Optimizer::OptimizeSim3(pKF1, pKF2, vpMatches1, g2oS12, th2, bFixScale): returns (nIn, matches after the
        call with -1 where the reference writes NULL, S12 after the call)."""
        mm = np.ascontiguousarray(match_mp, np.int32).copy()
        S = np.ascontiguousarray(s12, np.float64).copy()
        n_in = lib().hh_optimize_sim3(self.h, kf1, kf2, len(mm), _i(mm), S.ctypes.data_as(C.POINTER(C.c_double)),
                                      C.c_float(th2), int(fix_scale))
        return int(n_in), mm, S

    def set_covisible(self, kf, others):
        a = np.ascontiguousarray(others, np.int32)
        lib().hh_set_covisible(self.h, kf, _i(a), len(a))

    def set_bad(self, kf=-1, mp=-1):
        lib().hh_set_bad(self.h, kf, mp)

    def set_lidar(self, ld):
        """Lidar clouds of synth.LidarData put on the keyframes (KeyFrame.h:437-442) + the lidarConfig of the pass."""
        L = lib()
        c32 = lambda a: np.ascontiguousarray(a, np.float32)
        f, n, c = c32(ld.flat_xyz), c32(ld.flat_normal), c32(ld.corner_xyz)
        L.hh_set_lidar_cloud(self.h, ld.cur_pose, len(f), _f(f), _f(n), len(c), _f(c))
        for k in np.unique(np.concatenate([ld.map_flat_pose, ld.map_corner_pose])):
            f = c32(ld.map_flat_xyz[ld.map_flat_pose == k])
            c = c32(ld.map_corner_xyz[ld.map_corner_pose == k])
            L.hh_set_lidar_cloud(self.h, int(k), len(f), _f(f), None, len(c), _f(c))
        L.hh_set_lidar_config(self.h, int(ld.use_flat), int(ld.use_corner), ld.distance_sq_threshold, ld.flat_weight,
                              ld.corner_weight)

    # ---- incremental observation mirror (host/map_mirror.h)
    def mirror_attach(self):
        lib().hh_mirror_attach(self.h)

    @staticmethod
    def mirror_detach():
        lib().hh_mirror_detach()

    @staticmethod
    def mirror_points() -> int:
        return int(lib().hh_mirror_points())

    def add_observation(self, kf, mp, u, v, ur=-1.0, octave=0):
        lib().hh_add_observation(self.h, int(kf), int(mp), float(u), float(v), float(ur), int(octave))

    def erase_observation(self, kf, mp):
        lib().hh_erase_observation(self.h, int(kf), int(mp))

    def set_point_bad(self, mp):
        lib().hh_set_point_bad(self.h, int(mp))

    def mirror_mismatches(self) -> int:
        return int(lib().hh_mirror_mismatches(self.h))

    def mirror_stress(self, n_writers=4, rounds=3) -> int:
        return int(lib().hh_mirror_stress(self.h, int(n_writers), int(rounds)))

    def time_gather(self, kf, reps=5) -> float:
        L = lib()
        set_options(True, True)
        return float(L.hh_time_gather(self.h, int(kf), int(reps)))

    def reset_markers(self):
        """Clear the per-call window-selection stamps so that the same local window can be selected again."""
        lib().hh_reset_markers(self.h)

    def gather(self, kf=-1, stereo_edges=True):
        """The flat problem the adapter builds for LocalBundleAdjustment(kf) (kf >= 0) or for the whole map (-1),
        without solving it -- needs no GPU.  Returns a dict of arrays in the layout of sqrtba_set_problem + the ids."""
        L = lib()
        set_options(stereo_edges, True)
        if kf >= 0:
            self.reset_markers()
        sz = np.zeros(3, np.int32)
        L.hh_gather(self.h, kf, _i(sz))
        nk, nm, no = (int(v) for v in sz)
        out = dict(pose_qt=np.zeros((nk, 7)), pose_fixed=np.zeros(nk, np.uint8), cam=np.zeros((nk, 5)),
                   point_xyz=np.zeros((nm, 3)), obs_pose=np.zeros(no, np.int32), obs_point=np.zeros(no, np.int32),
                   obs_meas=np.zeros((no, 4), np.float32), kf_ids=np.zeros(nk, np.int64), mp_ids=np.zeros(nm, np.int64))
        dp = C.POINTER(C.c_double)
        L.hh_gather_get(out["pose_qt"].ctypes.data_as(dp), out["pose_fixed"].ctypes.data_as(C.POINTER(C.c_uint8)),
                        out["cam"].ctypes.data_as(dp), out["point_xyz"].ctypes.data_as(dp), _i(out["obs_pose"]),
                        _i(out["obs_point"]), _f(out["obs_meas"]), out["kf_ids"].ctypes.data_as(C.POINTER(C.c_int64)),
                        out["mp_ids"].ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def apply_local(self, kf, pose_qt, point_xyz, outlier, stereo_edges=True):
        """The write-back half of LocalBundleAdjustment(kf) with a result given in the layout of gather(kf)."""
        set_options(stereo_edges, True)
        P = np.ascontiguousarray(pose_qt, np.float64)
        X = np.ascontiguousarray(point_xyz, np.float64)
        F = np.ascontiguousarray(outlier, np.uint8)
        dp = C.POINTER(C.c_double)
        lib().hh_apply_local(self.h, kf, len(P), P.ctypes.data_as(dp), len(X), X.ctypes.data_as(dp), len(F),
                             F.ctypes.data_as(C.POINTER(C.c_uint8)))

    def local_ba(self, kf, stop=None, stereo_edges=True, two_pass=True):
        """Optimizer::LocalBundleAdjustment(kf, stop, map, lidarconfig).  stereo_edges=False, two_pass=False is the
        reference fork's behaviour and the adapter's own default (g2oOptimizer.cc:914-916, 1113-1114)."""
        set_options(stereo_edges, two_pass)
        lib().hh_local_ba(self.h, kf, stop)

    def global_ba(self, iters, robust, n_loop_kf=0, stop=None):
        lib().hh_global_ba(self.h, iters, int(robust), n_loop_kf, stop)

    def pose(self, kf, gba=False):
        out = np.zeros(16, np.float32)
        lib().hh_get_pose(self.h, kf, int(gba), _f(out))
        return out.reshape(4, 4)

    def point(self, mp, gba=False):
        out = np.zeros(3, np.float32)
        lib().hh_get_point(self.h, mp, int(gba), _f(out))
        return out

    def has_observation(self, kf, mp):
        return bool(lib().hh_has_observation(self.h, kf, mp))

    def keyframe_sees(self, kf, mp):
        return bool(lib().hh_keyframe_sees(self.h, kf, mp))

    def point_updates(self, mp):
        return lib().hh_point_updates(self.h, mp)

    def gba_marker(self, kf):
        return lib().hh_gba_marker(self.h, kf)

    def last_error(self):
        return lib().hh_last_error().decode()


class MockFrame:
    """A header-compatible Frame (every keypoint matched to its own MapPoint) for Optimizer::PoseOptimization."""

    def __init__(self, pose7, cam5, xyz, meas):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = synth.quat_to_rotmat(np.asarray(pose7)[3:]).astype(np.float32)
        T[:3, 3] = np.asarray(pose7)[:3].astype(np.float32)
        inv = synth.inv_level_sigma2()
        octave = np.array([int(np.argmin(np.abs(inv - s))) for s in meas[:, 3]], np.int32)
        assert np.all(inv[octave] == meas[:, 3])
        self.n = len(xyz)
        self._keep = [np.ascontiguousarray(T), np.ascontiguousarray(cam5, np.float32), inv,
                      np.ascontiguousarray(xyz, np.float32), np.ascontiguousarray(meas[:, :3], np.float32), octave]
        k = self._keep
        self.h = lib().hh_frame_build(_f(k[0]), _f(k[1]), _f(k[2]), len(inv), self.n, _f(k[3]), _f(k[4]), _i(k[5]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().hh_frame_destroy(self.h)
            self.h = None

    def unmatch(self, i):
        lib().hh_frame_unmatch(self.h, i)

    def pose_optimization(self) -> int:
        return lib().hh_pose_opt(self.h)

    def pose_optimization_lidar(self, ld) -> int:
        """Optimizer::PoseOptimization(pFrame, local_lidarmap_cloud_ptr, kdtree_local_map, lidarconfig) with the clouds
        of a synth.FrameLidar put on the frame (Frame.h:273-277) and into a PointICloud."""
        c32 = lambda a: np.ascontiguousarray(a, np.float32)
        f, n, c, m = c32(ld.flat_xyz), c32(ld.flat_normal), c32(ld.corner_xyz), c32(ld.map_xyz)
        return lib().hh_pose_opt_lidar(self.h, len(f), _f(f), _f(n), len(c), _f(c), len(m), _f(m), int(ld.use_flat),
                                       int(ld.use_corner), ld.distance_sq_threshold, ld.flat_weight, ld.corner_weight)

    def state(self):
        T = np.zeros(16, np.float32)
        out = np.zeros(max(self.n, 1), np.uint8)
        lib().hh_frame_get(self.h, _f(T), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return T.reshape(4, 4), out[:self.n]


def pose_optimization_batch(frames):
    arr = (C.c_void_p * len(frames))(*[f.h for f in frames])
    inl = np.zeros(len(frames), np.int32)
    lib().hh_pose_opt_batch(arr, len(frames), _i(inl))
    return inl
