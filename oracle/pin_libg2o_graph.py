"""Graph-level pinning: a REAL g2o::SparseOptimizer of the reference's prebuilt libg2o.so.

pin_libg2o.py pins the per-edge functions and pin_libg2o_edges.py whole edge objects; this script goes one level up and
pins the graph semantics the BA control flow relies on against the binary itself:

  * index mapping of initializeOptimization(level) (sparse_optimizer.cpp:166-190, 199-267, 482-487): free poses first,
    marginalised landmarks second, each in ascending vertex id whatever the insertion order; fixed vertices and
    vertices whose edges are all at another level get no index;
  * computeActiveErrors (sparse_optimizer.cpp:61-88) rewrites `_error` of ACTIVE edges only -- an edge moved to level 1
    keeps the error it had (the stale-_error rule behind the local-BA outlier test, g2oOptimizer.cc:947-976,1119-1142);
  * activeChi2 / activeRobustChi2 (sparse_optimizer.cpp:90-114) over the active edges, Huber with the float dsqr;
  * update(double*) (sparse_optimizer.cpp:422-435) consumes the increment in index order and applies oplusImpl;
    push()/pop() restore the estimates;
  * the normal equations (row A7): with the vertices' Hessian blocks and every edge's off-diagonal block mapped onto
    our own memory (mapHessianMemory, as BlockSolver::buildStructure does, block_solver.hpp:143-295), the binary's
    linearizeOplus + constructQuadraticForm of every active edge (base_binary_edge.hpp:55-120; Huber weighting through
    robustInformation, base_edge.h:96-102) give Hpp, Hll, the 6x3 pose-landmark blocks and the gradient b.

Objects are built with the binary's own constructors in raw 64-byte aligned blocks; the few inline setters (setId,
setFixed, setMarginalized, setLevel, setRobustKernel, setInformation) are replaced by writes at member offsets that are
CHECKED against constructor defaults first (id -1 at byte 8, hessianIndex -1 at 80, dimension at 88 for vertices;
id -1 at 32, dimension at 36, level at 40, robust kernel at 48 for edges; `_information` between `_measurement` and
`_error`, both located by pin_libg2o_edges.Edges.layout).  Huber kernels come from the binary's
RobustKernelCreator<RobustKernelHuber>::construct().

Beyond the graph semantics the script runs the binary's own optimisation loops -- SparseOptimizer::optimize +
OptimizationAlgorithmLevenberg::solve -- on a g2o::Solver implemented here (FakeSolver: a vtable of ctypes callbacks):
  * make_lm      two problems far from their optimum, optimize(30): the lambda of every trial, iterations, estimates;
  * make_lba     the two-pass schedule of LocalBundleAdjustment (5 + 10 iterations, chi2 / depth classification between);
  * make_poseopt the four-round schedule of PoseOptimization over real EdgeSE3ProjectXYZOnlyPose /
                 EdgeStereoSE3ProjectXYZOnlyPose objects (inline constructors: laid out around the exported vtables);
  * make_sim3 / make_posegraph   g2o::Sim3 arithmetic and the optimisation of OptimizeEssentialGraph (VertexSim3Expmap,
                 EdgeSim3 with numeric Jacobians, lambda_0 = 1e-16) -- the oracle of SURVEY row N3;
  * make_sim3opt the schedule of OptimizeSim3 over real EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ objects.

TEST INFRASTRUCTURE ONLY.  Run as a script in a clean interpreter; appends `graph_*`, `lm<k>_*`, `lba_*`, `po_*`, `sim3_*`,
`pg<k>_*` and `s3o<k>_*` arrays to tests/golden/libg2o_vectors.npz (tests/test_pin_libg2o.py compares the oracle with them)."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pin_libg2o as P  # noqa: E402
import pin_libg2o_edges as PE  # noqa: E402

SYM = {
    "opt_ctor": ("_ZN3g2o15SparseOptimizerC1Ev", None, 1),
    "add_vertex": ("_ZN3g2o16OptimizableGraph9addVertexEPNS_10HyperGraph6VertexEPNS0_4DataE", C.c_bool, 3),
    "add_edge": ("_ZN3g2o16OptimizableGraph7addEdgeEPNS_10HyperGraph4EdgeE", C.c_bool, 2),
    "init": ("_ZN3g2o15SparseOptimizer22initializeOptimizationEi", C.c_bool, None),
    "errors": ("_ZN3g2o15SparseOptimizer19computeActiveErrorsEv", None, 1),
    "chi2": ("_ZNK3g2o15SparseOptimizer10activeChi2Ev", C.c_double, 1),
    "rchi2": ("_ZNK3g2o15SparseOptimizer16activeRobustChi2Ev", C.c_double, 1),
    "update": ("_ZN3g2o15SparseOptimizer6updateEPKd", None, 2),
    "push": ("_ZN3g2o15SparseOptimizer4pushEv", None, 1),
    "pop": ("_ZN3g2o15SparseOptimizer3popEv", None, 1),
    "huber_new": ("_ZN3g2o19RobustKernelCreatorINS_17RobustKernelHuberEE9constructEv", C.c_void_p, 1),
    "pose_map": ("_ZN3g2o10BaseVertexILi6ENS_7SE3QuatEE16mapHessianMemoryEPd", None, 2),
    "pose_clear": ("_ZN3g2o10BaseVertexILi6ENS_7SE3QuatEE18clearQuadraticFormEv", None, 1),
    "pt_map": ("_ZN3g2o10BaseVertexILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEEE16mapHessianMemoryEPd", None, 2),
    "pt_clear": ("_ZN3g2o10BaseVertexILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEEE18clearQuadraticFormEv", None, 1),
    "e3_map": ("_ZN3g2o14BaseBinaryEdgeILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE16mapHessianMemoryEPdiib", None, -1),
    "e2_map": ("_ZN3g2o14BaseBinaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE16mapHessianMemoryEPdiib", None, -1),
    "e3_quad": ("_ZN3g2o14BaseBinaryEdgeILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE22constructQuadraticFormEv", None, 1),
    "e2_quad": ("_ZN3g2o14BaseBinaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE22constructQuadraticFormEv", None, 1),
}
V_ID, V_HIDX, V_FIXED, V_MARG, V_DIM = 8, 80, 84, 85, 88          # byte offsets inside a vertex
E_ID, E_DIM, E_LEVEL, E_KERNEL = 32, 36, 40, 48                    # byte offsets inside an edge
POINT_ID0 = 100                                                    # vertex id of landmark j is POINT_ID0 + j


class Graph:
    def __init__(self):
        self.ed = PE.Edges()
        L = self.ed.g.L
        self.f = {}
        for k, (name, res, nargs) in SYM.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = ([C.c_void_p, C.c_int] if nargs is None else
                           [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_bool] if nargs == -1 else [C.c_void_p] * nargs)
            self.f[k] = fn
        self.lay = {True: self.ed.layout(True), False: self.ed.layout(False)}
        self.opt = P._aligned(8192)
        self.f["opt_ctor"](self.opt.ctypes.data)
        self.keep, self.poses, self.points, self.edges = [], {}, {}, []
        self.pose_est_off = None

    # ---- vertices
    def _check_vertex(self, v, dim):
        i32 = v.view(np.int32)
        assert i32[V_ID // 4] == -1 and i32[V_HIDX // 4] == -1 and i32[V_DIM // 4] == dim, "unexpected vertex layout"
        assert v.view(np.uint8)[V_FIXED] == 0 and v.view(np.uint8)[V_MARG] == 0

    def add_pose(self, idx, upd, fixed):
        v = self.ed._pose_vertex(np.zeros(6))
        self._check_vertex(v, 6)
        if self.pose_est_off is None:  # _estimate: identity quaternion (x y z w) followed by t = 0 after setToOriginImpl
            for i in range(0, 500, 2):
                if v[i + 3] == 1.0 and not v[i:i + 3].any() and not v[i + 4:i + 7].any():
                    self.pose_est_off = i
                    break
            assert self.pose_est_off is not None
        u = P._aligned(6)
        u[:] = upd
        self.ed.g.f["vtx_oplus"](v.ctypes.data, u.ctypes.data)
        v.view(np.int32)[V_ID // 4] = idx
        v.view(np.uint8)[V_FIXED] = 1 if fixed else 0
        assert self.f["add_vertex"](self.opt.ctypes.data, v.ctypes.data, None)
        self.poses[idx] = v

    def add_point(self, j, X):
        v = self.ed._point_vertex(np.asarray(X, float))
        i32 = v.view(np.int32)
        assert i32[V_ID // 4] == -1 and i32[V_HIDX // 4] == -1 and i32[V_DIM // 4] == 3, "unexpected vertex layout"
        assert np.array_equal(v[19:22], X), "VertexSBAPointXYZ::_estimate is not where it was found"
        i32[V_ID // 4] = POINT_ID0 + j
        v.view(np.uint8)[V_MARG] = 1                                 # setMarginalized(true), g2oOptimizer.cc:868
        assert self.f["add_vertex"](self.opt.ctypes.data, v.ctypes.data, None)
        self.points[j] = v

    def pose_estimate(self, idx):
        e = self.poses[idx][self.pose_est_off:self.pose_est_off + 7]
        return np.array([e[4], e[5], e[6], e[0], e[1], e[2], e[3]])  # (t, q) like SE3Quat::toVector

    def point_estimate(self, j):
        return self.points[j][19:22].copy()

    # ---- edges
    def add_edge(self, pose, point, meas, info, cam, delta):
        stereo = not (meas[2] < 0)
        lay, d = self.lay[stereo], 3 if stereo else 2
        off = self.ed.g._stereo_off if stereo else self.ed.g._mono_off
        e, vbeg = self.ed._edge(stereo)
        i32 = e.view(np.int32)
        assert i32[E_ID // 4] == -1 and i32[E_DIM // 4] == d and i32[E_LEVEL // 4] == 0 and e.view(np.uint64)[E_KERNEL // 8] == 0
        ptrs = (C.c_uint64 * 2).from_address(vbeg)
        ptrs[0], ptrs[1] = self.points[point].ctypes.data, self.poses[pose].ctypes.data   # vertex 0 = landmark, 1 = pose
        for k, val in zip(("fx", "fy", "cx", "cy"), cam[:4]):
            e[off[k]] = val
        if stereo:
            e[lay["bf"]] = cam[4]
        e[lay["meas"]:lay["meas"] + d] = meas[:d]
        io = lay["err"] - d * d                                      # _information sits right before _error
        assert io >= lay["meas"] + d
        e[io:io + d * d] = (np.eye(d) * info).ravel()
        rk = self.f["huber_new"](None)                               # new RobustKernelHuber (with its vtable)
        self.ed.g.f["set_delta"](rk, float(delta))
        e.view(np.uint64)[E_KERNEL // 8] = rk
        assert self.f["add_edge"](self.opt.ctypes.data, e.ctypes.data)
        self.edges.append((e, stereo, rk))

    def set_levels(self, levels, robust):
        for (e, stereo, rk), lv in zip(self.edges, levels):
            e.view(np.int32)[E_LEVEL // 4] = int(lv)
            e.view(np.uint64)[E_KERNEL // 8] = rk if robust else 0   # e->setRobustKernel(0), g2oOptimizer.cc:969

    def errors(self):
        out = np.zeros((len(self.edges), 3))
        for k, (e, stereo, _) in enumerate(self.edges):
            lay = self.lay[stereo]
            out[k, :lay["d"]] = e[lay["err"]:lay["err"] + lay["d"]]
        return out

    def system(self, levels, fixed, obs):
        """Hpp / Hll / per-edge Hpl (6x3) / b of the active edges at the current estimates, assembled by the binary."""
        n_pose, n_point = len(self.poses), len(self.points)
        Hpp = [P._aligned(36) for _ in range(n_pose)]
        Hll = [P._aligned(12) for _ in range(n_point)]
        Hpl = [P._aligned(20) for _ in self.edges]
        for i in range(n_pose):
            self.f["pose_map"](self.poses[i].ctypes.data, Hpp[i].ctypes.data)
        for j in range(n_point):
            self.f["pt_map"](self.points[j].ctypes.data, Hll[j].ctypes.data)
        snap = {("p", i): self.poses[i].copy() for i in range(n_pose)}
        snap.update({("l", j): self.points[j].copy() for j in range(n_point)})
        for i in range(n_pose):
            self.f["pose_clear"](self.poses[i].ctypes.data)
        for j in range(n_point):
            self.f["pt_clear"](self.points[j].ctypes.data)
        jws = []
        for k, (e, stereo, _) in enumerate(self.edges):
            if levels[k] != 0:
                continue
            tag = "e3" if stereo else "e2"
            if not fixed[obs[k][0]]:   # BlockSolver maps the pose-landmark block transposed: pose index < landmark index
                self.f[tag + "_map"](e.ctypes.data, Hpl[k].ctypes.data, 0, 1, True)
            jw = P._aligned(64)
            self.ed.f["jw_ctor"](jw.ctypes.data)
            self.ed.f["jw_size"](jw.ctypes.data, e.ctypes.data)
            assert self.ed.f["jw_alloc"](jw.ctypes.data)
            self.ed.f["stereo_lin" if stereo else "mono_lin"](e.ctypes.data, jw.ctypes.data)
            self.f[tag + "_quad"](e.ctypes.data)
            jws.append(jw)
        # _b of a vertex: the D consecutive doubles that constructQuadraticForm moved (found by diffing the object)
        def b_of(kind, idx, D):
            v = self.poses[idx] if kind == "p" else self.points[idx]
            ch = np.flatnonzero(v.view(np.uint64) != snap[(kind, idx)].view(np.uint64))   # bitwise: some slots hold NaN patterns
            if len(ch) == 0:
                return np.zeros(D)
            assert ch.max() - ch.min() < D, (kind, idx, ch)
            return ch.min()
        boff = {}
        for kind, n, D in (("p", n_pose, 6), ("l", n_point, 3)):
            offs = [b_of(kind, i, D) for i in range(n)]
            offs = [o for o in offs if not isinstance(o, np.ndarray)]
            # the first changed slot may not be _b[0] if _b[0] stayed 0: take the smallest over all vertices of the kind
            boff[kind] = min(offs)
        self.boff = boff
        bp = np.stack([self.poses[i][boff["p"]:boff["p"] + 6] for i in range(n_pose)])
        bl = np.stack([self.points[j][boff["l"]:boff["l"] + 3] for j in range(n_point)])
        H6 = np.stack([h[:36].reshape(6, 6).T for h in Hpp])       # Eigen blocks are column-major
        H3 = np.stack([h[:9].reshape(3, 3).T for h in Hll])
        Hx = np.stack([h[:18].reshape(3, 6).T for h in Hpl])       # 6x3 column-major -> (6, 3)
        return dict(Hpp=H6, Hll=H3, Hpl=Hx, b_pose=bp, b_point=bl), jws

    def phase(self, levels, robust, update, n_pose, n_point):
        """set levels -> initializeOptimization(0) -> computeActiveErrors -> chi2s -> push, update, pop check -> update."""
        o = self.opt.ctypes.data
        self.set_levels(levels, robust)
        assert self.f["init"](o, 0)
        self.f["errors"](o)
        pidx = np.array([self.poses[i].view(np.int32)[V_HIDX // 4] for i in range(n_pose)], np.int32)
        lidx = np.array([self.points[j].view(np.int32)[V_HIDX // 4] for j in range(n_point)], np.int32)
        err = self.errors()
        chi, rchi = self.f["chi2"](o), self.f["rchi2"](o)
        before = [self.pose_estimate(i) for i in range(n_pose)] + [self.point_estimate(j) for j in range(n_point)]
        u = P._aligned(len(update) + 8)
        u[:len(update)] = update
        self.f["push"](o)
        self.f["update"](o, u.ctypes.data)
        moved = [self.pose_estimate(i) for i in range(n_pose)] + [self.point_estimate(j) for j in range(n_point)]
        self.f["pop"](o)
        after = [self.pose_estimate(i) for i in range(n_pose)] + [self.point_estimate(j) for j in range(n_point)]
        assert all(np.array_equal(a, b) for a, b in zip(before, after)), "pop() did not restore the estimates"
        assert any(not np.array_equal(a, b) for a, b in zip(before, moved))
        self.f["update"](o, u.ctypes.data)                            # keep the update for the next phase
        poses = np.stack([self.pose_estimate(i) for i in range(n_pose)])
        points = np.stack([self.point_estimate(j) for j in range(n_point)])
        return dict(pose_index=pidx, point_index=lidx, err=err, chi2=chi, robust_chi2=rchi, poses=poses, points=points)


class FakeSolver:
    """A g2o::Solver (core/solver.h:43-150) implemented HERE and handed to the binary's own
    OptimizationAlgorithmLevenberg: a hand-made vtable of ctypes callbacks in declaration order (two destructor
    slots, init, buildStructure, updateStructure, buildSystem, solve, computeMarginals, setLambda, restoreDiagonal,
    supportsSchur, schur, setSchur, setWriteDebug, writeDebug, saveHessian) in front of the members the inline accessors
    read (_optimizer, _x, _b, _xSize, _maxXSize, _isLevenberg, _additionalVectorSpace).  The linear algebra is the
    binary's linearizeOplus + constructQuadraticForm into blocks mapped here (what BlockSolver::buildSystem does,
    block_solver.hpp:502-560), lambda on every diagonal (setLambda, :564-589) and ONE dense solve of the full system
    (exactly what Schur complement + LDLT + back-substitution compute, :354-486).  Everything else -- lambda_0, the gain
    ratio, the lambda/nu policy, trial loop, push/pop/discardTop, stop rules -- is the reference binary's code."""

    def __init__(self, G, fixed, obs, levels):
        self.G, self.fixed, self.obs, self.levels = G, fixed, obs, levels
        self.log = []          # ("lambda", value) / ("solve", |x|) in call order
        vp, b, d = C.c_void_p, C.c_bool, C.c_double
        sig = [(None, [vp]), (None, [vp]), (b, [vp, vp, b]), (b, [vp, b]), (b, [vp, vp, vp]), (b, [vp]), (b, [vp]),
               (b, [vp, vp, vp]), (b, [vp, d, b]), (None, [vp]), (b, [vp]), (b, [vp]), (None, [vp, b]), (None, [vp, b]),
               (b, [vp]), (b, [vp, vp])]
        impl = [self._noop, self._noop, self._init, self._build_structure, self._false3, self._build_system, self._solve,
                self._false3, self._set_lambda, self._restore, self._false1, self._false1, self._noop2, self._noop2,
                self._false1, self._false2]
        self.cbs = [C.CFUNCTYPE(r, *a)(f) for (r, a), f in zip(sig, impl)]
        self.vtable = (C.c_void_p * (len(self.cbs) + 2))()
        for i, cb in enumerate(self.cbs):
            self.vtable[i + 2] = C.cast(cb, C.c_void_p).value           # [0] offset-to-top, [1] RTTI stay 0
        self.obj = P._aligned(16)
        u = self.obj.view(np.uint64)
        u[0] = C.addressof(self.vtable) + 16
        u[1] = G.opt.ctypes.data                                          # _optimizer
        u[6] = 1                                                          # _isLevenberg

    # ---- trivial slots
    def _noop(self, this): return None
    def _noop2(self, this, flag): return None
    def _false1(self, this): return False
    def _false2(self, this, a): return False
    def _false3(self, this, a, b): return False
    def _init(self, this, opt, online): return True

    def _build_structure(self, this, zero):
        G = self.G
        n_pose, n_point = len(G.poses), len(G.points)
        ent = [(int(G.poses[i].view(np.int32)[V_HIDX // 4]), "p", i) for i in range(n_pose)]
        ent += [(int(G.points[j].view(np.int32)[V_HIDX // 4]), "l", j) for j in range(n_point)]
        ent = sorted(e for e in ent if e[0] >= 0)
        self.off, n = {}, 0
        for _, kind, i in ent:
            self.off[(kind, i)] = n
            n += 6 if kind == "p" else 3
        self.n = n
        self.Hpp = {i: P._aligned(36) for (k, i) in self.off if k == "p"}
        self.Hll = {j: P._aligned(12) for (k, j) in self.off if k == "l"}
        for i, h in self.Hpp.items():
            G.f["pose_map"](G.poses[i].ctypes.data, h.ctypes.data)
        for j, h in self.Hll.items():
            G.f["pt_map"](G.points[j].ctypes.data, h.ctypes.data)
        self.Hpl, self.jw = {}, {}
        for k, (e, stereo, _) in enumerate(G.edges):
            if self.levels[k] != 0:
                continue
            i, j = int(self.obs[k][0]), int(self.obs[k][1])
            if ("p", i) in self.off and ("l", j) in self.off:
                self.Hpl[k] = P._aligned(20)
                G.f[("e3" if stereo else "e2") + "_map"](e.ctypes.data, self.Hpl[k].ctypes.data, 0, 1, True)
            jw = P._aligned(64)
            G.ed.f["jw_ctor"](jw.ctypes.data)
            G.ed.f["jw_size"](jw.ctypes.data, e.ctypes.data)
            assert G.ed.f["jw_alloc"](jw.ctypes.data)
            self.jw[k] = jw
        self.x, self.b = P._aligned(n + 8), P._aligned(n + 8)
        u = self.obj.view(np.uint64)
        u[2], u[3], u[4], u[5] = self.x.ctypes.data, self.b.ctypes.data, n, n
        return True

    def _build_system(self, this):
        G = self.G
        for h in list(self.Hpp.values()) + list(self.Hll.values()) + list(self.Hpl.values()):
            h[:] = 0.0
        for (kind, i) in self.off:
            G.f["pose_clear" if kind == "p" else "pt_clear"]((G.poses if kind == "p" else G.points)[i].ctypes.data)
        for k, jw in self.jw.items():
            e, stereo, _ = G.edges[k]
            G.ed.f["stereo_lin" if stereo else "mono_lin"](e.ctypes.data, jw.ctypes.data)
            G.f[("e3" if stereo else "e2") + "_quad"](e.ctypes.data)
        for (kind, i), o in self.off.items():                            # v->copyB(_b + colInHessian)
            D = 6 if kind == "p" else 3
            v = (G.poses if kind == "p" else G.points)[i]
            self.b[o:o + D] = v[G.boff[kind]:G.boff[kind] + D]
        return True

    def _diag(self):
        for i, h in self.Hpp.items():
            yield h, [r * 6 + r for r in range(6)]
        for j, h in self.Hll.items():
            yield h, [r * 3 + r for r in range(3)]

    def _set_lambda(self, this, lam, backup):
        self.log.append(("lambda", float(lam)))
        if backup:
            self.backup = [(h, idx, h[idx].copy()) for h, idx in self._diag()]
        for h, idx in self._diag():
            h[idx] += lam
        return True

    def _restore(self, this):
        for h, idx, val in self.backup:
            h[idx] = val
        return None

    def _solve(self, this):
        H = np.zeros((self.n, self.n))
        for i, h in self.Hpp.items():
            o = self.off[("p", i)]
            H[o:o + 6, o:o + 6] = h[:36].reshape(6, 6).T
        for j, h in self.Hll.items():
            o = self.off[("l", j)]
            H[o:o + 3, o:o + 3] = h[:9].reshape(3, 3).T
        for k, h in self.Hpl.items():
            op, ol = self.off[("p", int(self.obs[k][0]))], self.off[("l", int(self.obs[k][1]))]
            blk = h[:18].reshape(3, 6).T                                 # 6x3
            H[op:op + 6, ol:ol + 3] += blk
            H[ol:ol + 3, op:op + 6] += blk.T
        x = np.linalg.solve(H, self.b[:self.n])
        self.x[:self.n] = x
        self.log.append(("solve", float(np.linalg.norm(x))))
        return True


def run_lm(G, fixed, obs, levels, iters):
    """SparseOptimizer::optimize(iters) of the binary with its own OptimizationAlgorithmLevenberg on top of FakeSolver."""
    L = G.ed.g.L
    lm_ctor = L._ZN3g2o30OptimizationAlgorithmLevenbergC1EPNS_6SolverE
    lm_ctor.restype, lm_ctor.argtypes = None, [C.c_void_p, C.c_void_p]
    set_alg = L._ZN3g2o15SparseOptimizer12setAlgorithmEPNS_21OptimizationAlgorithmE
    set_alg.restype, set_alg.argtypes = None, [C.c_void_p, C.c_void_p]
    optimize = L._ZN3g2o15SparseOptimizer8optimizeEib
    optimize.restype, optimize.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_bool]
    solver = FakeSolver(G, fixed, obs, levels)
    alg = P._aligned(1024)
    lm_ctor(alg.ctypes.data, solver.obj.ctypes.data)
    set_alg(G.opt.ctypes.data, alg.ctypes.data)
    n_it = optimize(G.opt.ctypes.data, iters, False)
    return n_it, solver, alg


def scenario(seed=7):
    """A small local-BA-shaped graph: 6 poses (0 and 3 fixed), 8 landmarks, 30 mono/stereo observations."""
    rng = np.random.default_rng(seed)
    n_pose, n_point = 6, 8
    cam = np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448]).astype(np.float32).astype(np.float64)
    upd = rng.normal(0, 1, (n_pose, 6)) * np.array([0.03, 0.03, 0.03, 0.5, 0.2, 0.5])
    fixed = np.array([1, 0, 0, 1, 0, 0], np.uint8)
    X = np.stack([rng.uniform(-6, 6, n_point), rng.uniform(-2, 2, n_point), rng.uniform(8, 30, n_point)], 1)
    obs = []
    for j in range(n_point):
        for i in rng.choice(n_pose, size=int(rng.integers(3, 5)), replace=False):
            obs.append((int(i), j))
    obs.sort(key=lambda t: (t[1], t[0]))                              # grouped by landmark, pose-sorted inside
    obs = np.array(obs, np.int32)
    n_obs = len(obs)
    stereo = rng.random(n_obs) < 0.6
    info = np.float32(1.0) / (np.float32(1.2) ** rng.integers(0, 6, n_obs)).astype(np.float32) ** 2
    return dict(n_pose=n_pose, n_point=n_point, cam=cam, upd=upd, fixed=fixed, X=X, obs=obs, stereo=stereo,
                info=info.astype(np.float32), rng=rng)


def make(path, seed=7):
    S = scenario(seed)
    rng = S["rng"]
    G = Graph()
    order = rng.permutation(S["n_pose"])                             # insertion order != id order
    for i in order:
        G.add_pose(int(i), S["upd"][i], bool(S["fixed"][i]))
    for j in rng.permutation(S["n_point"]):
        G.add_point(int(j), S["X"][j])
    pose0 = np.stack([G.pose_estimate(i) for i in range(S["n_pose"])])
    # measurements: projection at the current estimates + noise, some gross errors so that Huber's outlier branch is hit
    Rm = []
    for p in pose0:
        x, y, z, w = p[3:]
        Rm.append(np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]))
    meas = np.zeros((len(S["obs"]), 4), np.float32)
    for k, (i, j) in enumerate(S["obs"]):
        Xc = Rm[i] @ S["X"][j] + pose0[i, :3]
        u = S["cam"][0] * Xc[0] / Xc[2] + S["cam"][2] + rng.normal(0, 1.0) + (25.0 if k % 7 == 3 else 0.0)
        v = S["cam"][1] * Xc[1] / Xc[2] + S["cam"][3] + rng.normal(0, 1.0)
        ur = u - S["cam"][4] / Xc[2] + rng.normal(0, 1.0) if S["stereo"][k] else -1.0
        meas[k] = (u, v, ur, S["info"][k])
    d2, d3 = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815)))   # g2oOptimizer.cc:851-853
    for k, (i, j) in enumerate(S["obs"]):
        G.add_edge(int(i), int(j), meas[k].astype(np.float64), float(meas[k, 3]), S["cam"], d3 if S["stereo"][k] else d2)
    n_obs = len(S["obs"])
    # phase A: everything at level 0, Huber on (pass 1 of local BA)
    levA = np.zeros(n_obs, np.int32)
    nfreeA = int((S["fixed"] == 0).sum())
    updA = rng.normal(0, 1, 6 * nfreeA + 3 * S["n_point"]) * 0.01
    G.set_levels(levA, True)
    assert G.f["init"](G.opt.ctypes.data, 0)
    G.f["errors"](G.opt.ctypes.data)
    sysA, keep_jw = G.system(levA, S["fixed"], S["obs"])            # normal equations of pass 1 at the initial estimates
    A = G.phase(levA, True, updA, S["n_pose"], S["n_point"])
    # phase B: kernels off; some edges to level 1 -- among them EVERY edge of landmark 2 and every edge of free pose 4,
    # which therefore drop out of the index mapping (g2oOptimizer.cc:947-970 + sparse_optimizer.cpp:218-259)
    levB = (rng.random(n_obs) < 0.15).astype(np.int32)
    levB[S["obs"][:, 1] == 2] = 1
    levB[S["obs"][:, 0] == 4] = 1
    B0 = dict(pose_active=np.zeros(S["n_pose"], bool), point_active=np.zeros(S["n_point"], bool))
    for k, (i, j) in enumerate(S["obs"]):
        if levB[k] == 0:
            B0["pose_active"][i] = True
            B0["point_active"][j] = True
    nB = int((B0["pose_active"] & (S["fixed"] == 0)).sum()) * 6 + int(B0["point_active"].sum()) * 3
    updB = rng.normal(0, 1, nB) * 0.01
    B = G.phase(levB, False, updB, S["n_pose"], S["n_point"])
    out = dict(np.load(path)) if os.path.exists(path) else {}
    out.update(graph_pose=pose0, graph_fixed=S["fixed"], graph_X=S["X"], graph_obs=S["obs"], graph_meas=meas,
               graph_cam=S["cam"], graph_levA=levA, graph_updA=updA, graph_levB=levB, graph_updB=updB)
    for k, v in sysA.items():
        out[f"graph_sysA_{k}"] = v
    for tag, R in (("A", A), ("B", B)):
        for k, v in R.items():
            out[f"graph_{tag}_{k}"] = np.asarray(v)
    np.savez(path, **out)
    return out


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def lm_scenario(exp, seed, rs, ts, xs, outliers, n_pose=7, n_point=30):
    """A small BA problem that is far from its optimum (pose noise rs rad / ts m, landmark noise xs m, gross outliers):
    Levenberg needs rejected trials and gain ratios in the unclamped range, and stops by its own rule."""
    rng = np.random.default_rng(seed)
    cam = np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448]).astype(np.float32).astype(np.float64)
    upd_t = rng.normal(0, 1, (n_pose, 6)) * np.array([0.03, 0.03, 0.03, 0.8, 0.2, 0.8])
    fixed = np.zeros(n_pose, np.uint8)
    fixed[0] = fixed[3] = 1
    pose_t = np.stack([exp(u) for u in upd_t])
    upd0 = upd_t + rng.normal(0, 1, (n_pose, 6)) * np.array([rs, rs, rs, ts, ts, ts]) * (fixed == 0)[:, None]
    X = np.stack([rng.uniform(-8, 8, n_point), rng.uniform(-3, 3, n_point), rng.uniform(8, 35, n_point)], 1)
    X0 = X + rng.normal(0, xs, X.shape)
    obs = []
    for j in range(n_point):
        for i in rng.choice(n_pose, size=int(rng.integers(3, 6)), replace=False):
            obs.append((int(i), j))
    obs.sort(key=lambda t: (t[1], t[0]))
    obs = np.array(obs, np.int32)
    stereo = rng.random(len(obs)) < 0.6
    info = (np.float32(1.0) / (np.float32(1.2) ** rng.integers(0, 6, len(obs))).astype(np.float32) ** 2).astype(np.float32)
    meas = np.zeros((len(obs), 4), np.float32)
    for k, (i, j) in enumerate(obs):
        Xc = _rot(pose_t[i, 3:]) @ X[j] + pose_t[i, :3]
        out = rng.random() < outliers
        u = cam[0] * Xc[0] / Xc[2] + cam[2] + rng.normal(0, 1) + (rng.uniform(20, 60) if out else 0)
        v = cam[1] * Xc[1] / Xc[2] + cam[3] + rng.normal(0, 1)
        ur = u - cam[4] / Xc[2] + rng.normal(0, 1) if stereo[k] else -1.0
        meas[k] = (u, v, ur, info[k])
    return dict(cam=cam, upd0=upd0, fixed=fixed, X0=X0, obs=obs, meas=meas, n_pose=n_pose, n_point=n_point)


def make_lm(path, iters=30):
    """The reference binary's SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve on FakeSolver: records
    the lambda handed to every trial (a complete fingerprint of the gain ratios and of the accept / reject / stop
    decisions), the number of iterations optimize() returns and the final estimates, as `lm<k>_*` arrays."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    cases = [(17, 0.25, 1.5, 4.0, 0.15, True), (34, 0.08, 0.6, 1.5, 0.10, False)]
    keep = []
    for c, (seed, rs, ts, xs, of, robust) in enumerate(cases):
        G = Graph()
        S = lm_scenario(G.ed.g.se3_exp, seed, rs, ts, xs, of)
        for i in range(S["n_pose"]):
            G.add_pose(i, S["upd0"][i], bool(S["fixed"][i]))
        for j in range(S["n_point"]):
            G.add_point(j, S["X0"][j])
        pose0 = np.stack([G.pose_estimate(i) for i in range(S["n_pose"])])
        d2, d3 = float(np.float32(np.sqrt(5.99))), float(np.float32(np.sqrt(7.815)))   # g2oOptimizer.cc:163-164 (GBA)
        for k, (i, j) in enumerate(S["obs"]):
            m = S["meas"][k].astype(np.float64)
            G.add_edge(int(i), int(j), m, float(m[3]), S["cam"], d3 if not (m[2] < 0) else d2)
        lev = np.zeros(len(S["obs"]), np.int32)
        G.set_levels(lev, robust)
        o = G.opt.ctypes.data
        assert G.f["init"](o, 0)
        G.f["errors"](o)
        chi0 = G.f["rchi2"](o)
        G.system(lev, S["fixed"], S["obs"])                          # locates the vertices' _b
        n_it, solver, alg = run_lm(G, S["fixed"], S["obs"], lev, iters)
        G.f["errors"](o)
        chi1 = G.f["rchi2"](o)
        lam = np.array([v for k, v in solver.log if k == "lambda"])
        out.update({f"lm{c}_pose0": pose0, f"lm{c}_fixed": S["fixed"], f"lm{c}_X0": S["X0"], f"lm{c}_obs": S["obs"],
                    f"lm{c}_meas": S["meas"], f"lm{c}_cam": S["cam"], f"lm{c}_robust": np.array(int(robust)),
                    f"lm{c}_iters": np.array(iters), f"lm{c}_lambda": lam, f"lm{c}_n_iterations": np.array(n_it),
                    f"lm{c}_chi2": np.array([chi0, chi1]),
                    f"lm{c}_poses": np.stack([G.pose_estimate(i) for i in range(S["n_pose"])]),
                    f"lm{c}_points": np.stack([G.point_estimate(j) for j in range(S["n_point"])])})
        keep.append((G, solver, alg))
        print(f"lm case {c}: optimize -> {n_it} iterations, {len(lam)} trials, "
              f"{int((lam[1:] > lam[:-1]).sum())} rejected, chi2 {chi0:.1f} -> {chi1:.1f}")
    np.savez(path, **out)
    return out


def make_lba(path):
    """The two-pass schedule of g2oOptimizer::LocalBundleAdjustment (g2oOptimizer.cc:923-976, 1119-1142) driven over the
    binary's objects: optimize(5) with Huber kernels; every edge with chi2() > 5.991 / 7.815 or non-positive depth goes
    to level 1 and every kernel is dropped; initializeOptimization(0); optimize(10); the same test again gives the
    outlier flags.  chi2() and isDepthPositive() are inline in the reference (base_edge.h:58-61,
    types_six_dof_expmap.h:97-101,129-133) and evaluated here from the binary's stored `_error` / estimates, so the
    stale-_error rule (level-1 edges keep their pass-1 error and are always flagged) comes from the binary itself."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    G = Graph()
    S = lm_scenario(G.ed.g.se3_exp, 5, 0.02, 0.15, 0.4, 0.15, n_pose=8, n_point=40)
    for i in range(S["n_pose"]):
        G.add_pose(i, S["upd0"][i], bool(S["fixed"][i]))
    for j in range(S["n_point"]):
        G.add_point(j, S["X0"][j])
    pose0 = np.stack([G.pose_estimate(i) for i in range(S["n_pose"])])
    d2, d3 = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815)))   # g2oOptimizer.cc:851-853
    for k, (i, j) in enumerate(S["obs"]):
        m = S["meas"][k].astype(np.float64)
        G.add_edge(int(i), int(j), m, float(m[3]), S["cam"], d3 if not (m[2] < 0) else d2)
    n_obs = len(S["obs"])
    o = G.opt.ctypes.data

    def classify():
        err = G.errors()
        flags = np.zeros(n_obs, np.uint8)
        for k, (i, j) in enumerate(S["obs"]):
            stereo = G.edges[k][1]
            info = float(S["meas"][k, 3])
            d = 3 if stereo else 2
            chi = float(sum(err[k, c] * (info * err[k, c]) for c in range(d)))     # _error.dot(information() * _error)
            P7 = G.pose_estimate(int(i))
            z = (_rot(P7[3:]) @ G.point_estimate(int(j)) + P7[:3])[2]
            flags[k] = 1 if (chi > (7.815 if stereo else 5.991) or not (z > 0.0)) else 0
        return flags

    lev = np.zeros(n_obs, np.int32)
    G.set_levels(lev, True)
    assert G.f["init"](o, 0)
    G.f["errors"](o)
    G.system(lev, S["fixed"], S["obs"])
    n1, solver, alg = run_lm(G, S["fixed"], S["obs"], lev, 5)
    lam1 = [v for k, v in solver.log if k == "lambda"]
    lev2 = classify().astype(np.int32)                              # :947-970
    solver.levels[:] = lev2
    G.set_levels(lev2, False)
    assert G.f["init"](o, 0)
    optimize = G.ed.g.L._ZN3g2o15SparseOptimizer8optimizeEib
    n2 = optimize(o, 10, False)
    lam2 = [v for k, v in solver.log if k == "lambda"][len(lam1):]
    flags = classify()                                              # :1125-1142
    out.update(lba_pose0=pose0, lba_fixed=S["fixed"], lba_X0=S["X0"], lba_obs=S["obs"], lba_meas=S["meas"],
               lba_cam=S["cam"], lba_lambda1=np.array(lam1), lba_lambda2=np.array(lam2),
               lba_n_iterations=np.array([n1, n2]), lba_level2=lev2, lba_outlier=flags,
               lba_poses=np.stack([G.pose_estimate(i) for i in range(S["n_pose"])]),
               lba_points=np.stack([G.point_estimate(j) for j in range(S["n_point"])]))
    print(f"lba: pass 1 {n1} iterations / {len(lam1)} trials, {int(lev2.sum())} of {n_obs} edges to level 1, "
          f"pass 2 {n2} iterations / {len(lam2)} trials, {int(flags.sum())} outliers, "
          f"{int(((lev2 == 1) & (flags == 1)).sum())} of them frozen since pass 1")
    np.savez(path, **out)
    return out, (G, solver, alg)


# ---------------------------------------------------------------------------------------------------------------------
# pose-only optimisation: g2oOptimizer::PoseOptimization over the binary's objects
# ---------------------------------------------------------------------------------------------------------------------
class PoseOnly:
    """One VertexSE3Expmap + EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose edges of the binary
    (types_six_dof_expmap.h:143-202).  Their constructors are inline, so the objects are laid out here the way the
    inline constructors would (the exported vtable, a one-element `_vertices` vector, id -1, dimension D; everything
    else zero) at the member offsets checked for the binary edges; Xw / fx.. are located by probing cam_project."""
    U = {False: dict(vt="_ZTVN3g2o25EdgeSE3ProjectXYZOnlyPoseE", err="_ZN3g2o25EdgeSE3ProjectXYZOnlyPose12computeErrorEv",
                     cam="_ZNK3g2o25EdgeSE3ProjectXYZOnlyPose11cam_projectERKN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE",
                     lin="_ZN3g2o13BaseUnaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_15VertexSE3ExpmapEE14linearizeOplusERNS_17JacobianWorkspaceE",
                     quad="_ZN3g2o13BaseUnaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_15VertexSE3ExpmapEE22constructQuadraticFormEv"),
         True: dict(vt="_ZTVN3g2o31EdgeStereoSE3ProjectXYZOnlyPoseE", err="_ZN3g2o31EdgeStereoSE3ProjectXYZOnlyPose12computeErrorEv",
                    cam="_ZNK3g2o31EdgeStereoSE3ProjectXYZOnlyPose11cam_projectERKN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE",
                    lin="_ZN3g2o13BaseUnaryEdgeILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEENS_15VertexSE3ExpmapEE14linearizeOplusERNS_17JacobianWorkspaceE",
                    quad="_ZN3g2o13BaseUnaryEdgeILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEENS_15VertexSE3ExpmapEE22constructQuadraticFormEv")}

    def __init__(self):
        self.G = Graph()
        L = self.G.ed.g.L
        self.fn = {}
        for st, d in self.U.items():
            f = {}
            f["vt"] = C.addressof(C.c_char.in_dll(L, d["vt"])) + 16
            for k in ("err", "quad"):
                f[k] = getattr(L, d[k]); f[k].restype = None; f[k].argtypes = [C.c_void_p]
            f["lin"] = getattr(L, d["lin"]); f["lin"].restype = None; f["lin"].argtypes = [C.c_void_p, C.c_void_p]
            f["cam"] = getattr(L, d["cam"]); f["cam"].restype = C.c_void_p; f["cam"].argtypes = [C.c_void_p] * 3
            self.fn[st] = f
        self.cam_off = {st: self._probe_cam(st) for st in (False, True)}
        self.edges, self.keep = [], []

    def _probe_cam(self, stereo):
        found = {}
        xyz = P._aligned(4)
        xyz[:3] = [2.0, 3.0, 1.0]
        for i in range(20, 80):
            this = P._aligned(256)
            this[i] = 1.0
            out = P._aligned(4)
            self.fn[stereo]["cam"](out.ctypes.data, this.ctypes.data, xyz.ctypes.data)
            r = out[:3]
            if r[0] == 2.0 and r[1] == 0.0:
                found["fx"] = i
            elif r[0] == 1.0 and r[1] == 0.0:
                found["cx"] = i
            elif r[1] == 3.0 and r[0] == 0.0:
                found["fy"] = i
            elif r[1] == 1.0 and r[0] == 0.0:
                found["cy"] = i
            elif stereo and r[0] == 0.0 and r[1] == 0.0 and r[2] != 0.0:
                found["bf"] = i
        assert {"fx", "fy", "cx", "cy"} <= set(found) and (not stereo or "bf" in found), found
        assert found["fy"] == found["fx"] + 1 and found["cx"] == found["fx"] + 2 and found["cy"] == found["fx"] + 3
        return found

    def add_edge(self, Xw, meas, cam, delta):
        G = self.G
        stereo = not (meas[2] < 0)
        d = 3 if stereo else 2
        lay, off = G.lay[stereo], self.cam_off[stereo]
        e = P._aligned(PE.OBJ)
        u, i32 = e.view(np.uint64), e.view(np.int32)
        u[0] = self.fn[stereo]["vt"]
        slot = P._aligned(2)                                          # storage of the one-element _vertices vector
        slot.view(np.uint64)[0] = G.poses[0].ctypes.data
        u[1], u[2], u[3] = slot.ctypes.data, slot.ctypes.data + 8, slot.ctypes.data + 8
        i32[E_ID // 4], i32[E_DIM // 4], i32[E_LEVEL // 4] = -1, d, 0
        e[lay["meas"]:lay["meas"] + d] = meas[:d]
        io = lay["err"] - d * d
        e[io:io + d * d] = (np.eye(d) * float(meas[3])).ravel()
        e[off["fx"] - 3:off["fx"]] = Xw                               # Vector3d Xw sits right before fx
        for k, val in zip(("fx", "fy", "cx", "cy"), cam[:4]):
            e[off[k]] = val
        if stereo:
            e[off["bf"]] = cam[4]
        rk = G.f["huber_new"](None)
        G.ed.g.f["set_delta"](rk, float(delta))
        u[E_KERNEL // 8] = rk
        assert G.f["add_edge"](G.opt.ctypes.data, e.ctypes.data)
        jw = P._aligned(64)
        G.ed.f["jw_ctor"](jw.ctypes.data)
        G.ed.f["jw_size"](jw.ctypes.data, e.ctypes.data)
        assert G.ed.f["jw_alloc"](jw.ctypes.data)
        self.edges.append(dict(e=e, stereo=stereo, d=d, rk=rk, jw=jw, info=float(meas[3]), lay=lay))
        self.keep.append(slot)

    def error(self, k):
        ed = self.edges[k]
        return ed["e"][ed["lay"]["err"]:ed["lay"]["err"] + ed["d"]].copy()

    def chi2(self, k):
        ed, err = self.edges[k], self.error(k)
        return float(sum(err[c] * (ed["info"] * err[c]) for c in range(ed["d"])))

    def set_level(self, k, lv):
        self.edges[k]["e"].view(np.int32)[E_LEVEL // 4] = lv

    def level(self, k):
        return int(self.edges[k]["e"].view(np.int32)[E_LEVEL // 4])


class FakeSolverUnary(FakeSolver):
    """The dense 6x6 case (LinearSolverDense in the reference, linear_solver_dense.h:65-113): one pose, unary edges."""

    def __init__(self, po):
        self.po = po
        super().__init__(po.G, None, None, None)

    def _build_structure(self, this, zero):
        G = self.G
        assert G.poses[0].view(np.int32)[V_HIDX // 4] == 0
        self.n = 6
        self.H = P._aligned(36)
        G.f["pose_map"](G.poses[0].ctypes.data, self.H.ctypes.data)
        self.x, self.b = P._aligned(16), P._aligned(16)
        u = self.obj.view(np.uint64)
        u[2], u[3], u[4], u[5] = self.x.ctypes.data, self.b.ctypes.data, 6, 6
        return True

    def _build_system(self, this):
        G, po = self.G, self.po
        self.H[:] = 0.0
        G.f["pose_clear"](G.poses[0].ctypes.data)
        for k, ed in enumerate(po.edges):
            if po.level(k) != 0:
                continue
            po.fn[ed["stereo"]]["lin"](ed["e"].ctypes.data, ed["jw"].ctypes.data)
            po.fn[ed["stereo"]]["quad"](ed["e"].ctypes.data)
        self.b[:6] = G.poses[0][G.boff["p"]:G.boff["p"] + 6]
        return True

    def _diag(self):
        yield self.H, [r * 6 + r for r in range(6)]

    def _solve(self, this):
        x = np.linalg.solve(self.H[:36].reshape(6, 6).T, self.b[:6])
        self.x[:6] = x
        self.log.append(("solve", float(np.linalg.norm(x))))
        return True


def make_poseopt(path, seed=3, n=150):
    """g2oOptimizer::PoseOptimization (g2oOptimizer.cc:385-559, 655-690) over the binary: four rounds of
    setEstimate(initial) / initializeOptimization(0) / optimize(10), after each the chi2 re-classification
    (computeError() on the edges that were outliers, `const float chi2 = e->chi2()`, level 1 / 0, kernels dropped in the
    third round), then the final classification.  Records lambda per trial and round, flags, inlier count, pose."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    po = PoseOnly()
    G = po.G
    rng = np.random.default_rng(seed)
    cam = np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448]).astype(np.float32).astype(np.float64)
    upd_t = rng.normal(0, 1, 6) * np.array([0.05, 0.05, 0.05, 1.0, 0.3, 1.0])
    pose_t = G.ed.g.se3_exp(upd_t)
    upd0 = upd_t + rng.normal(0, 1, 6) * np.array([0.03, 0.03, 0.03, 0.3, 0.3, 0.3])
    Xc = np.stack([rng.uniform(-10, 10, n), rng.uniform(-4, 4, n), rng.uniform(5, 50, n)], 1)
    R = _rot(pose_t[3:])
    Xw = ((Xc - pose_t[:3]) @ R).astype(np.float32).astype(np.float64)     # map points are float in the reference
    meas = np.zeros((n, 4), np.float32)
    for k in range(n):
        c = R @ Xw[k] + pose_t[:3]
        bad = rng.random() < 0.2
        u = cam[0] * c[0] / c[2] + cam[2] + rng.normal(0, 1) + (rng.uniform(8, 40) * rng.choice([-1, 1]) if bad else 0)
        v = cam[1] * c[1] / c[2] + cam[3] + rng.normal(0, 1)
        ur = u - cam[4] / c[2] + rng.normal(0, 1) if rng.random() < 0.5 else -1.0
        meas[k] = (u, v, ur, np.float32(1.0) / np.float32(1.2) ** (2 * int(rng.integers(0, 5))))
    G.add_pose(0, upd0, False)
    pose0 = G.pose_estimate(0)
    est0 = G.poses[0][G.pose_est_off:G.pose_est_off + 7].copy()
    dm, ds = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815)))   # g2oOptimizer.cc:426-428
    for k in range(n):
        m = meas[k].astype(np.float64)
        po.add_edge(Xw[k], m, cam, ds if not (m[2] < 0) else dm)
    o = G.opt.ctypes.data
    # locate the pose vertex's _b (differential probe on a throw-away quadratic form)
    assert G.f["init"](o, 0)
    G.f["errors"](o)
    Hs = P._aligned(36)
    G.f["pose_map"](G.poses[0].ctypes.data, Hs.ctypes.data)
    G.f["pose_clear"](G.poses[0].ctypes.data)
    snap = G.poses[0].copy()
    ed0 = po.edges[0]
    po.fn[ed0["stereo"]]["lin"](ed0["e"].ctypes.data, ed0["jw"].ctypes.data)
    po.fn[ed0["stereo"]]["quad"](ed0["e"].ctypes.data)
    ch = np.flatnonzero(G.poses[0].view(np.uint64) != snap.view(np.uint64))
    assert 0 < len(ch) <= 6 and ch.max() - ch.min() < 6, ch
    G.boff = {"p": int(ch.min())}
    G.f["pose_clear"](G.poses[0].ctypes.data)

    L = G.ed.g.L
    lm_ctor = L._ZN3g2o30OptimizationAlgorithmLevenbergC1EPNS_6SolverE
    lm_ctor.restype, lm_ctor.argtypes = None, [C.c_void_p, C.c_void_p]
    set_alg = L._ZN3g2o15SparseOptimizer12setAlgorithmEPNS_21OptimizationAlgorithmE
    set_alg.restype, set_alg.argtypes = None, [C.c_void_p, C.c_void_p]
    optimize = L._ZN3g2o15SparseOptimizer8optimizeEib
    optimize.restype, optimize.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_bool]
    solver = FakeSolverUnary(po)
    alg = P._aligned(1024)
    lm_ctor(alg.ctypes.data, solver.obj.ctypes.data)
    set_alg(o, alg.ctypes.data)

    outlier = np.zeros(n, np.uint8)
    lam_rounds, n_its = [], []
    for it in range(4):
        G.poses[0][G.pose_est_off:G.pose_est_off + 7] = est0             # vSE3->setEstimate(initial), :510
        assert G.f["init"](o, 0)
        before = len(solver.log)
        n_its.append(optimize(o, 10, False))
        lam_rounds.append([v for k, v in solver.log[before:] if k == "lambda"])
        for k, ed in enumerate(po.edges):                                # :518-547
            if outlier[k]:
                po.fn[ed["stereo"]]["err"](ed["e"].ctypes.data)
            chi = np.float32(po.chi2(k))
            if chi > np.float32(7.815 if ed["stereo"] else 5.991):
                outlier[k] = 1
                po.set_level(k, 1)
            else:
                outlier[k] = 0
                po.set_level(k, 0)
            if it == 2:
                ed["e"].view(np.uint64)[E_KERNEL // 8] = 0
    for k, ed in enumerate(po.edges):                                    # final classification, :656-680
        if outlier[k]:
            po.fn[ed["stereo"]]["err"](ed["e"].ctypes.data)
        chi = np.float32(po.chi2(k))
        outlier[k] = 1 if float(chi) > (7.815 if ed["stereo"] else 5.991) else 0
    lam = np.concatenate([np.array(r) for r in lam_rounds])
    out.update(po_pose0=pose0, po_cam=cam, po_xyz=Xw, po_meas=meas, po_lambda=lam,
               po_round_trials=np.array([len(r) for r in lam_rounds]), po_n_iterations=np.array(n_its),
               po_outlier=outlier, po_inliers=np.array(n - int(outlier.sum())), po_pose=G.pose_estimate(0))
    print(f"pose-only: trials per round {[len(r) for r in lam_rounds]}, iterations {n_its}, {int(outlier.sum())} of {n} outliers")
    np.savez(path, **out)
    return out, (po, solver, alg)


# ---------------------------------------------------------------------------------------------------------------------
# essential-graph (Sim3 pose-graph) optimisation, SURVEY.md 8(f) N3: g2o::Sim3, VertexSim3Expmap, EdgeSim3 of the binary
# ---------------------------------------------------------------------------------------------------------------------
class Sim3Graph:
    """VertexSim3Expmap / EdgeSim3 objects (types_seven_dof_expmap.h:48-110) built with the binary's constructors; the
    estimate (Sim3: quaternion x y z w | t | s), `_fix_scale` and `_b` are located by probing, the measurement goes in
    through BaseEdge<7,Sim3>::setMeasurement, information / error through informationData() / errorData()."""
    S = {
        "v_ctor": ("_ZN3g2o16VertexSim3ExpmapC1Ev", None, 1), "v_origin": ("_ZN3g2o16VertexSim3Expmap15setToOriginImplEv", None, 1),
        "v_oplus": ("_ZN3g2o16VertexSim3Expmap9oplusImplEPKd", None, 2),
        "v_map": ("_ZN3g2o10BaseVertexILi7ENS_4Sim3EE16mapHessianMemoryEPd", None, 2),
        "v_clear": ("_ZN3g2o10BaseVertexILi7ENS_4Sim3EE18clearQuadraticFormEv", None, 1),
        "sim3_exp": ("_ZN3g2o4Sim3C1ERKN5Eigen6MatrixIdLi7ELi1ELi0ELi7ELi1EEE", None, 2),
        "e_ctor": ("_ZN3g2o8EdgeSim3C1Ev", None, 1), "e_err": ("_ZN3g2o8EdgeSim312computeErrorEv", None, 1),
        "e_meas": ("_ZN3g2o8BaseEdgeILi7ENS_4Sim3EE14setMeasurementERKS1_", None, 2),
        "e_info": ("_ZN3g2o8BaseEdgeILi7ENS_4Sim3EE15informationDataEv", C.c_void_p, 1),
        "e_errd": ("_ZN3g2o8BaseEdgeILi7ENS_4Sim3EE9errorDataEv", C.c_void_p, 1),
        "e_lin": ("_ZN3g2o14BaseBinaryEdgeILi7ENS_4Sim3ENS_16VertexSim3ExpmapES2_E14linearizeOplusERNS_17JacobianWorkspaceE", None, 2),
        "e_quad": ("_ZN3g2o14BaseBinaryEdgeILi7ENS_4Sim3ENS_16VertexSim3ExpmapES2_E22constructQuadraticFormEv", None, 1),
        "e_map": ("_ZN3g2o14BaseBinaryEdgeILi7ENS_4Sim3ENS_16VertexSim3ExpmapES2_E16mapHessianMemoryEPdiib", None, -1),
        "lam_init": ("_ZN3g2o30OptimizationAlgorithmLevenberg17setUserLambdaInitEd", None, -2),
    }

    def __init__(self):
        self.G = Graph()
        L = self.G.ed.g.L
        self.f = {}
        for k, (name, res, nargs) in self.S.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = ([C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_bool] if nargs == -1 else
                           [C.c_void_p, C.c_double] if nargs == -2 else [C.c_void_p] * nargs)
            self.f[k] = fn
        self.verts, self.edges = [], []
        v = self.new_vertex_block()
        # _estimate: identity quaternion, t = 0, s = 1 after setToOriginImpl
        self.est = None
        for i in range(0, 400, 2):
            if list(v[i:i + 8]) == [0, 0, 0, 1, 0, 0, 0, 1]:
                u = P._aligned(8)
                u[:7] = [0.1, 0.2, -0.1, 1, 2, 3, 0.3]
                before = v[i:i + 8].copy()
                self.f["v_oplus"](v.ctypes.data, u.ctypes.data)
                if not np.array_equal(before, v[i:i + 8]):
                    self.est = i
                    break
        assert self.est is not None, "VertexSim3Expmap::_estimate not found"
        # _fix_scale: the bool after the four Vector2d members that makes oplus ignore update[6]
        self.fix_off = None
        for byte in range((self.est + 8) * 8, (self.est + 8) * 8 + 256):
            w = self.new_vertex_block()
            if w.view(np.uint8)[byte] != 0:
                continue
            w.view(np.uint8)[byte] = 1
            u = P._aligned(8)
            u[:7] = [0, 0, 0, 0, 0, 0, 0.5]
            self.f["v_oplus"](w.ctypes.data, u.ctypes.data)
            if w[self.est + 7] == 1.0:
                self.fix_off = byte
                break
        assert self.fix_off is not None, "VertexSim3Expmap::_fix_scale not found"

    def new_vertex_block(self):
        v = P._aligned(PE.OBJ)
        self.f["v_ctor"](v.ctypes.data)
        self.f["v_origin"](v.ctypes.data)
        return v

    def exp(self, upd7):
        out, u = P._aligned(8), P._aligned(8)
        u[:7] = upd7
        self.f["sim3_exp"](out.ctypes.data, u.ctypes.data)
        return out.copy()

    def oplus(self, s8, upd7, fix_scale):
        v = self.new_vertex_block()
        v[self.est:self.est + 8] = s8
        v.view(np.uint8)[self.fix_off] = 1 if fix_scale else 0
        u = P._aligned(8)
        u[:7] = upd7
        self.f["v_oplus"](v.ctypes.data, u.ctypes.data)
        return v[self.est:self.est + 8].copy()

    def add_vertex(self, idx, s8, fixed, fix_scale):
        v = self.new_vertex_block()
        i32 = v.view(np.int32)
        assert i32[V_ID // 4] == -1 and i32[V_HIDX // 4] == -1 and i32[V_DIM // 4] == 7
        v[self.est:self.est + 8] = s8
        i32[V_ID // 4] = idx
        v.view(np.uint8)[V_FIXED] = 1 if fixed else 0
        v.view(np.uint8)[self.fix_off] = 1 if fix_scale else 0
        assert self.G.f["add_vertex"](self.G.opt.ctypes.data, v.ctypes.data, None)
        self.verts.append(v)

    def estimate(self, i):
        return self.verts[i][self.est:self.est + 8].copy()

    def new_edge(self, vi, vj, meas8):
        e = P._aligned(PE.OBJ)
        self.f["e_ctor"](e.ctypes.data)
        vec = PE._vector_slots(e, 16)
        assert vec and vec[0][0] == 1
        ptrs = (C.c_uint64 * 2).from_address(vec[0][1])
        ptrs[0], ptrs[1] = vi.ctypes.data, vj.ctypes.data
        m = P._aligned(8)
        m[:] = meas8
        self.f["e_meas"](e.ctypes.data, m.ctypes.data)
        info = np.ctypeslib.as_array((C.c_double * 49).from_address(self.f["e_info"](e.ctypes.data)))
        info[:] = np.eye(7).ravel()
        return e

    def edge_error(self, meas8, a8, b8):
        va, vb = self.new_vertex_block(), self.new_vertex_block()
        va[self.est:self.est + 8] = a8
        vb[self.est:self.est + 8] = b8
        e = self.new_edge(va, vb, meas8)
        self.f["e_err"](e.ctypes.data)
        return np.ctypeslib.as_array((C.c_double * 7).from_address(self.f["e_errd"](e.ctypes.data))).copy()

    def add_edge(self, i, j, meas8):
        e = self.new_edge(self.verts[i], self.verts[j], meas8)
        assert self.G.f["add_edge"](self.G.opt.ctypes.data, e.ctypes.data)
        jw = P._aligned(64)
        self.G.ed.f["jw_ctor"](jw.ctypes.data)
        self.G.ed.f["jw_size"](jw.ctypes.data, e.ctypes.data)
        assert self.G.ed.f["jw_alloc"](jw.ctypes.data)
        self.edges.append(dict(e=e, i=i, j=j, jw=jw))


class FakeSolverSim3(FakeSolver):
    """BlockSolver_7_3 without marginalised vertices (g2oOptimizer.cc:1220-1230): every free vertex in the one system."""

    def __init__(self, sg):
        self.sg = sg
        super().__init__(sg.G, None, None, None)

    def _build_structure(self, this, zero):
        sg = self.sg
        ent = sorted((int(v.view(np.int32)[V_HIDX // 4]), i) for i, v in enumerate(sg.verts))
        self.off = {i: 7 * k for k, (h, i) in enumerate(e for e in ent if e[0] >= 0)}
        self.n = 7 * len(self.off)
        self.Hv = {i: P._aligned(52) for i in self.off}
        for i, h in self.Hv.items():
            sg.f["v_map"](sg.verts[i].ctypes.data, h.ctypes.data)
        self.He = {}
        for k, ed in enumerate(sg.edges):
            if ed["i"] in self.off and ed["j"] in self.off:
                self.He[k] = P._aligned(52)
                sg.f["e_map"](ed["e"].ctypes.data, self.He[k].ctypes.data, 0, 1, False)
        self.x, self.b = P._aligned(self.n + 8), P._aligned(self.n + 8)
        u = self.obj.view(np.uint64)
        u[2], u[3], u[4], u[5] = self.x.ctypes.data, self.b.ctypes.data, self.n, self.n
        return True

    def _build_system(self, this):
        sg = self.sg
        for h in list(self.Hv.values()) + list(self.He.values()):
            h[:] = 0.0
        for i in self.off:
            sg.f["v_clear"](sg.verts[i].ctypes.data)
        for ed in sg.edges:
            sg.f["e_lin"](ed["e"].ctypes.data, ed["jw"].ctypes.data)
            sg.f["e_quad"](ed["e"].ctypes.data)
        for i, o in self.off.items():
            self.b[o:o + 7] = sg.verts[i][sg.b_off:sg.b_off + 7]
        return True

    def _diag(self):
        for i, h in self.Hv.items():
            yield h, [r * 7 + r for r in range(7)]

    def _solve(self, this):
        H = np.zeros((self.n, self.n))
        for i, h in self.Hv.items():
            o = self.off[i]
            H[o:o + 7, o:o + 7] = h[:49].reshape(7, 7).T
        for k, h in self.He.items():
            oi, oj = self.off[self.sg.edges[k]["i"]], self.off[self.sg.edges[k]["j"]]
            blk = h[:49].reshape(7, 7).T                                  # Ji^T Jj
            H[oi:oi + 7, oj:oj + 7] += blk
            H[oj:oj + 7, oi:oi + 7] += blk.T
        x = np.linalg.solve(H, self.b[:self.n])
        self.x[:self.n] = x
        self.log.append(("solve", float(np.linalg.norm(x))))
        return True


def _sim3_mul(a, b):
    """numpy helper for building scenarios only (not a pinned quantity): Sim3 product of 8-vectors."""
    def qm(p, q):
        x1, y1, z1, w1 = p
        x2, y2, z2, w2 = q
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                         w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    return np.concatenate([qm(a[:4], b[:4]), a[7] * (_rot(a[:4]) @ b[4:7]) + a[4:7], [a[7] * b[7]]])


def _sim3_inv(a):
    qi = np.array([-a[0], -a[1], -a[2], a[3]])
    return np.concatenate([qi, _rot(qi) @ (-a[4:7] / a[7]), [1.0 / a[7]]])


def make_sim3(path, n=200, seed=11):
    """Sim3 arithmetic of the binary: exp (Sim3(Vector7d)), oplusImpl with and without _fix_scale, EdgeSim3 error
    (= log of C * v1 * v2^-1, which exercises operator*, inverse and log)."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    sg = Sim3Graph()
    rng = np.random.default_rng(seed)
    upd = rng.normal(0, 1, (n, 7)) * np.array([0.4, 0.4, 0.4, 2, 2, 2, 0.3])
    upd[:10, :3] *= 1e-7          # theta < eps
    upd[5:20, 6] *= 1e-7          # |sigma| < eps (rows 5-9 have both small)
    ex = np.stack([sg.exp(u) for u in upd])
    upd2 = rng.normal(0, 1, (n, 7)) * np.array([0.05, 0.05, 0.05, 0.3, 0.3, 0.3, 0.05])
    op_free = np.stack([sg.oplus(ex[k], upd2[k], False) for k in range(n)])
    op_fix = np.stack([sg.oplus(ex[k], upd2[k], True) for k in range(n)])
    # edge error: v2 close to meas * v1 so that the log stays in its principal range, plus small-angle / small-scale cases
    v1 = ex
    pert = rng.normal(0, 1, (n, 7)) * np.array([0.2, 0.2, 0.2, 1, 1, 1, 0.2])
    pert[:10, :3] *= 1e-7
    pert[5:20, 6] *= 1e-8
    meas = np.stack([sg.exp(rng.normal(0, 1, 7) * np.array([0.3, 0.3, 0.3, 1, 1, 1, 0.2])) for _ in range(n)])
    v2 = np.stack([_sim3_mul(_sim3_inv(sg.exp(pert[k])), _sim3_mul(meas[k], v1[k])) for k in range(n)])
    err = np.stack([sg.edge_error(meas[k], v1[k], v2[k]) for k in range(n)])
    out.update(sim3_upd=upd, sim3_exp=ex, sim3_upd2=upd2, sim3_oplus_free=op_free, sim3_oplus_fix=op_fix,
               sim3_meas=meas, sim3_v1=v1, sim3_v2=v2, sim3_err=err)
    np.savez(path, **out)
    print(f"sim3: estimate at double {sg.est}, _fix_scale at byte {sg.fix_off}, max |edge error| {np.abs(err).max():.3f}")
    return out


def make_posegraph(path):
    """The optimisation of g2oOptimizer::OptimizeEssentialGraph (g2oOptimizer.cc:1212-1232, 1472-1478) on the binary:
    VertexSim3Expmap per keyframe (one fixed), EdgeSim3 with identity information, Levenberg with
    setUserLambdaInit(1e-16), optimize(20) -- on a drifting loop with a closing edge, once with fixed scale (stereo)
    and once with free scale (monocular)."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    keep = []
    for case, fix_scale in enumerate((True, False)):
        sg = Sim3Graph()
        G = sg.G
        rng = np.random.default_rng(21 + case)
        n = 24
        # ground truth: a circle; odometry = true relative motion + noise (+ scale drift when the scale is free)
        truth = []
        for k in range(n):
            ang = 2 * np.pi * k / n
            q = np.array([0, np.sin(ang / 2), 0, np.cos(ang / 2)])
            c = np.array([10 * np.sin(ang), 0.0, 10 * (1 - np.cos(ang))])
            truth.append(np.concatenate([q, -_rot(q) @ c, [1.0]]))      # S_kw
        edges, meas = [], []
        est = [truth[0].copy()]
        for k in range(1, n):
            rel = _sim3_mul(truth[k], _sim3_inv(truth[k - 1]))          # S_k,k-1
            noise = sg.exp(rng.normal(0, 1, 7) * np.array([0.01, 0.01, 0.01, 0.05, 0.02, 0.05, 0.0 if fix_scale else 0.01]))
            m = _sim3_mul(noise, rel)
            edges.append((k - 1, k))                                     # vertex 0 = i, vertex 1 = j, measurement S_ji
            meas.append(m)
            est.append(_sim3_mul(m, est[k - 1]))                         # dead reckoning: drifts
        edges.append((n - 1, 0))                                         # the loop closes
        meas.append(_sim3_mul(truth[0], _sim3_inv(truth[n - 1])))
        for k in range(2, n, 5):                                         # a few covisibility edges
            edges.append((k - 2, k))
            meas.append(_sim3_mul(truth[k], _sim3_inv(truth[k - 2])))
        est, meas, edges = np.stack(est), np.stack(meas), np.array(edges, np.int32)
        fixed = np.zeros(n, np.uint8)
        fixed[0] = 1                                                     # pLoopKF
        for k in range(n):
            sg.add_vertex(k, est[k], bool(fixed[k]), fix_scale)
        for (i, j), m in zip(edges, meas):
            sg.add_edge(int(i), int(j), m)
        o = G.opt.ctypes.data
        assert G.f["init"](o, 0)
        G.f["errors"](o)
        chi0 = G.f["chi2"](o)
        # locate _b of the Sim3 vertex
        Hs = P._aligned(52)
        sg.f["v_map"](sg.verts[1].ctypes.data, Hs.ctypes.data)
        sg.f["v_clear"](sg.verts[1].ctypes.data)
        snap = sg.verts[1].copy()
        e0 = sg.edges[0]
        He = P._aligned(52)
        sg.f["e_map"](e0["e"].ctypes.data, He.ctypes.data, 0, 1, False)
        Hz = P._aligned(52)
        sg.f["v_map"](sg.verts[0].ctypes.data, Hz.ctypes.data)
        sg.f["e_lin"](e0["e"].ctypes.data, e0["jw"].ctypes.data)
        sg.f["e_quad"](e0["e"].ctypes.data)
        ch = np.flatnonzero(sg.verts[1].view(np.uint64) != snap.view(np.uint64))
        # _b is the first run of changed doubles (7 long; its scale component stays 0 when the scale is fixed); further
        # up the object the backup stack's bookkeeping moves too (push / pop of the numeric Jacobian)
        assert len(ch) >= 3 and (ch < ch.min() + 7).sum() >= 3 and ch.min() > V_DIM // 8, ch
        sg.b_off = int(ch.min())
        sg.f["v_clear"](sg.verts[1].ctypes.data)
        # the numeric Jacobians of every edge at the initial estimates (BaseBinaryEdge::linearizeOplus, central
        # differences through oplusImpl): read from the JacobianWorkspace, 7x7 column-major per vertex
        e_all, Ji_all, Jj_all = [], [], []
        for ed in sg.edges:
            sg.f["e_err"](ed["e"].ctypes.data)
            sg.f["e_lin"](ed["e"].ctypes.data, ed["jw"].ctypes.data)
            ws = PE._vector_slots(ed["jw"], 32)
            assert ws, "JacobianWorkspace::_workspace not found"
            pw = (C.c_uint64 * 4).from_address(ws[0][1])
            Ji = np.ctypeslib.as_array((C.c_double * 49).from_address(pw[0])).copy().reshape(7, 7).T
            Jj = np.ctypeslib.as_array((C.c_double * 49).from_address(pw[2])).copy().reshape(7, 7).T
            e_all.append(np.ctypeslib.as_array((C.c_double * 7).from_address(sg.f["e_errd"](ed["e"].ctypes.data))).copy())
            Ji_all.append(Ji if not fixed[ed["i"]] else np.zeros((7, 7)))   # a fixed vertex's Jacobian is never computed
            Jj_all.append(Jj if not fixed[ed["j"]] else np.zeros((7, 7)))
        out.update({f"pg{case}_err0": np.stack(e_all), f"pg{case}_Ji0": np.stack(Ji_all), f"pg{case}_Jj0": np.stack(Jj_all)})
        L = G.ed.g.L
        lm_ctor = L._ZN3g2o30OptimizationAlgorithmLevenbergC1EPNS_6SolverE
        lm_ctor.restype, lm_ctor.argtypes = None, [C.c_void_p, C.c_void_p]
        set_alg = L._ZN3g2o15SparseOptimizer12setAlgorithmEPNS_21OptimizationAlgorithmE
        set_alg.restype, set_alg.argtypes = None, [C.c_void_p, C.c_void_p]
        optimize = L._ZN3g2o15SparseOptimizer8optimizeEib
        optimize.restype, optimize.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_bool]
        solver = FakeSolverSim3(sg)
        alg = P._aligned(1024)
        lm_ctor(alg.ctypes.data, solver.obj.ctypes.data)
        sg.f["lam_init"](alg.ctypes.data, 1e-16)                         # solver->setUserLambdaInit(1e-16), :1230
        set_alg(o, alg.ctypes.data)
        n_it = optimize(o, 20, False)
        G.f["errors"](o)
        chi1 = G.f["chi2"](o)
        lam = np.array([v for k, v in solver.log if k == "lambda"])
        out.update({f"pg{case}_vert0": est, f"pg{case}_fixed": fixed, f"pg{case}_fix_scale": np.array(int(fix_scale)),
                    f"pg{case}_edges": edges, f"pg{case}_meas": meas, f"pg{case}_lambda": lam,
                    f"pg{case}_n_iterations": np.array(n_it), f"pg{case}_chi2": np.array([chi0, chi1]),
                    f"pg{case}_vert": np.stack([sg.estimate(k) for k in range(n)])})
        keep.append((sg, solver, alg, Hs, He, Hz))
        print(f"pose graph (fix_scale={fix_scale}): optimize -> {n_it} iterations, {len(lam)} trials, chi2 {chi0:.4f} -> {chi1:.6f}")
    np.savez(path, **out)
    return out, keep


class FakeSolverSim3Dense(FakeSolver):
    """The one-vertex dense case of OptimizeSim3 (BlockSolverX + LinearSolverDense, g2oOptimizer.cc:1565-1573): the
    7x7 block of the Sim3 vertex, the fixed points contribute no unknowns."""

    def __init__(self, so):
        self.so = so
        super().__init__(so.sg.G, None, None, None)

    def _build_structure(self, this, zero):
        so = self.so
        assert so.vs.view(np.int32)[V_HIDX // 4] == 0
        self.n = 7
        self.H = P._aligned(52)
        so.sg.f["v_map"](so.vs.ctypes.data, self.H.ctypes.data)
        self.x, self.b = P._aligned(16), P._aligned(16)
        u = self.obj.view(np.uint64)
        u[2], u[3], u[4], u[5] = self.x.ctypes.data, self.b.ctypes.data, 7, 7
        return True

    def _build_system(self, this):
        so = self.so
        self.H[:] = 0.0
        so.sg.f["v_clear"](so.vs.ctypes.data)
        for ed in so.edges:
            if ed["e"].view(np.int32)[E_LEVEL // 4] != 0:
                continue
            so.f["lin"](ed["e"].ctypes.data, ed["jw"].ctypes.data)
            so.f["quad"](ed["e"].ctypes.data)
        self.b[:7] = so.vs[so.b_off:so.b_off + 7]
        return True

    def _diag(self):
        yield self.H, [r * 7 + r for r in range(7)]

    def _solve(self, this):
        x = np.linalg.solve(self.H[:49].reshape(7, 7).T, self.b[:7])
        self.x[:7] = x
        self.log.append(("solve", float(np.linalg.norm(x))))
        return True


class Sim3Opt:
    """The graph of g2oOptimizer::OptimizeSim3 (g2oOptimizer.cc:1560-1796) out of the binary's objects: one
    VertexSim3Expmap (with both cameras' intrinsics -- the four Vector2d members in front of `_fix_scale`,
    types_seven_dof_expmap.h:71-72), fixed VertexSBAPointXYZ, EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ with Huber
    kernels.  optimizer.removeEdge is replaced by setLevel(1): either way the edge leaves the active set of the next
    initializeOptimization(0) and the order of the others is kept."""
    B2 = "_ZN3g2o14BaseBinaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_17VertexSBAPointXYZENS_16VertexSim3ExpmapEE"
    E2 = "_ZN3g2o8BaseEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEEE"
    S = {
        "e12_ctor": ("_ZN3g2o18EdgeSim3ProjectXYZC1Ev", None, 1), "e21_ctor": ("_ZN3g2o25EdgeInverseSim3ProjectXYZC1Ev", None, 1),
        "e12_err": ("_ZN3g2o18EdgeSim3ProjectXYZ12computeErrorEv", None, 1),
        "e21_err": ("_ZN3g2o25EdgeInverseSim3ProjectXYZ12computeErrorEv", None, 1),
        "lin": (B2 + "14linearizeOplusERNS_17JacobianWorkspaceE", None, 2), "quad": (B2 + "22constructQuadraticFormEv", None, 1),
        "meas": (E2 + "14setMeasurementERKS3_", None, 2), "info": (E2 + "15informationDataEv", C.c_void_p, 1),
        "errd": (E2 + "9errorDataEv", C.c_void_p, 1), "chi2": ("_ZNK3g2o8BaseEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEEE4chi2Ev", C.c_double, 1),
    }

    def __init__(self, s12, cam8, fix_scale):
        self.sg = Sim3Graph()
        L = self.sg.G.ed.g.L
        self.f = {}
        for k, (name, res, nargs) in self.S.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, [C.c_void_p] * nargs
            self.f[k] = fn
        sg = self.sg
        assert sg.fix_off % 8 == 0
        self.cam_off = sg.fix_off // 8 - 8                 # _principle_point1, _principle_point2, _focal_length1, _focal_length2
        sg.add_vertex(0, s12, False, fix_scale)
        self.vs = sg.verts[0]
        fx1, fy1, cx1, cy1, fx2, fy2, cx2, cy2 = cam8
        self.vs[self.cam_off:self.cam_off + 8] = [cx1, cy1, cx2, cy2, fx1, fy1, fx2, fy2]
        self.edges, self.points = [], []

    def _point(self, vid, X):
        G = self.sg.G
        v = G.ed._point_vertex(np.asarray(X, float))
        i32 = v.view(np.int32)
        assert i32[V_ID // 4] == -1 and i32[V_DIM // 4] == 3 and np.array_equal(v[19:22], X)
        i32[V_ID // 4] = vid
        v.view(np.uint8)[V_FIXED] = 1
        assert G.f["add_vertex"](G.opt.ctypes.data, v.ctypes.data, None)
        self.points.append(v)
        return v

    def _edge(self, kind, vp, obs, info, delta):
        G = self.sg.G
        e = P._aligned(PE.OBJ)
        self.f[kind + "_ctor"](e.ctypes.data)
        i32 = e.view(np.int32)
        assert i32[E_ID // 4] == -1 and i32[E_DIM // 4] == 2 and i32[E_LEVEL // 4] == 0 and e.view(np.uint64)[E_KERNEL // 8] == 0
        vec = PE._vector_slots(e, 16)
        assert vec
        ptrs = (C.c_uint64 * 2).from_address(vec[0][1])
        ptrs[0], ptrs[1] = vp.ctypes.data, self.vs.ctypes.data           # vertex 0 = point, vertex 1 = the Sim3
        m = P._aligned(2)
        m[:] = obs
        self.f["meas"](e.ctypes.data, m.ctypes.data)
        inf = np.ctypeslib.as_array((C.c_double * 4).from_address(self.f["info"](e.ctypes.data)))
        inf[:] = (np.eye(2) * info).ravel()
        rk = G.f["huber_new"](None)
        G.ed.g.f["set_delta"](rk, float(delta))
        e.view(np.uint64)[E_KERNEL // 8] = rk
        assert G.f["add_edge"](G.opt.ctypes.data, e.ctypes.data)
        jw = P._aligned(64)
        G.ed.f["jw_ctor"](jw.ctypes.data)
        G.ed.f["jw_size"](jw.ctypes.data, e.ctypes.data)
        assert G.ed.f["jw_alloc"](jw.ctypes.data)
        ed = dict(e=e, kind=kind, jw=jw, rk=rk)
        self.edges.append(ed)
        return ed

    def add_match(self, i, P1c, P2c, meas6, delta):
        v1, v2 = self._point(2 * i + 1, P1c), self._point(2 * (i + 1), P2c)
        e12 = self._edge("e12", v2, meas6[0:2], meas6[2], delta)       # x1 = S12 * X2
        e21 = self._edge("e21", v1, meas6[3:5], meas6[5], delta)       # x2 = S21 * X1
        return e12, e21

    def error(self, ed):
        self.f[ed["kind"] + "_err"](ed["e"].ctypes.data)
        return self.stored_error(ed)

    def stored_error(self, ed):
        return np.ctypeslib.as_array((C.c_double * 2).from_address(self.f["errd"](ed["e"].ctypes.data))).copy()

    def jacobian(self, ed):
        """numeric 2x7 Jacobian w.r.t. the Sim3 (vertex 1) out of the JacobianWorkspace (column-major)"""
        self.f["lin"](ed["e"].ctypes.data, ed["jw"].ctypes.data)
        ws = PE._vector_slots(ed["jw"], 32)
        assert ws, "JacobianWorkspace::_workspace not found"
        pw = (C.c_uint64 * 4).from_address(ws[0][1])
        return np.ctypeslib.as_array((C.c_double * 14).from_address(pw[2])).copy().reshape(7, 2).T

    def chi2(self, ed):
        return self.f["chi2"](ed["e"].ctypes.data)

    def set_level(self, ed, lvl):
        ed["e"].view(np.int32)[E_LEVEL // 4] = lvl


def make_sim3opt(path):
    """g2oOptimizer::OptimizeSim3 over the binary (fixed and free scale): errors and numeric Jacobians of both edge types
    at the initial estimate, then the whole schedule -- optimize(5), chi2 test on the stored errors, optimize(10 / 5),
    final count -- with the binary's Levenberg; records lambda per trial, kept matches, nIn and the final S12."""
    out = dict(np.load(path)) if os.path.exists(path) else {}
    keep_alive = []
    for case, fix_scale in enumerate((True, False)):
        rng = np.random.default_rng(40 + case)
        n = 48
        cam8 = np.array([718.856, 718.856, 607.1928, 185.2157, 707.0912, 707.0912, 601.8873, 183.1104]).astype(np.float32).astype(np.float64)
        sg0 = Sim3Graph()
        S12 = sg0.exp(rng.normal(0, 1, 7) * np.array([0.05, 0.05, 0.05, 0.6, 0.2, 0.6, 0.0 if fix_scale else 0.1]))
        S21 = _sim3_inv(S12)
        p1 = np.stack([rng.uniform(-8, 8, n), rng.uniform(-3, 3, n), rng.uniform(5, 40, n)], 1)
        p2 = np.stack([S21[7] * (_rot(S21[:4]) @ x) + S21[4:7] for x in p1]) * (1 + rng.normal(0, 0.01, (n, 1)))
        p1, p2 = p1.astype(np.float32).astype(np.float64), p2.astype(np.float32).astype(np.float64)
        meas = np.zeros((n, 6), np.float32)
        for k in range(n):
            bad = rng.random() < 0.2
            jump = rng.uniform(8, 40) * rng.choice([-1, 1]) if bad else 0.0
            meas[k, 0] = cam8[0] * p1[k, 0] / p1[k, 2] + cam8[2] + rng.normal(0, 1) + (jump if k % 2 == 0 else 0)
            meas[k, 1] = cam8[1] * p1[k, 1] / p1[k, 2] + cam8[3] + rng.normal(0, 1)
            meas[k, 2] = np.float32(1.0) / np.float32(1.2) ** (2 * int(rng.integers(0, 5)))
            meas[k, 3] = cam8[4] * p2[k, 0] / p2[k, 2] + cam8[6] + rng.normal(0, 1) + (jump if k % 2 == 1 else 0)
            meas[k, 4] = cam8[5] * p2[k, 1] / p2[k, 2] + cam8[7] + rng.normal(0, 1)
            meas[k, 5] = np.float32(1.0) / np.float32(1.2) ** (2 * int(rng.integers(0, 5)))
        noise = sg0.exp(rng.normal(0, 1, 7) * np.array([0.01, 0.01, 0.01, 0.05, 0.05, 0.05, 0.0 if fix_scale else 0.02]))
        s0 = _sim3_mul(noise, S12)
        th2 = np.float32(10.0)                                           # LoopClosing.cc: OptimizeSim3(..., 10, mbFixScale)
        delta = float(np.float32(np.sqrt(th2)))                          # const float deltaHuber = sqrt(th2), :1619
        so = Sim3Opt(s0, cam8, fix_scale)
        sg, G = so.sg, so.sg.G
        pairs = [so.add_match(k, p1[k], p2[k], meas[k].astype(np.float64), delta) for k in range(n)]
        o = G.opt.ctypes.data
        assert G.f["init"](o, 0)
        # locate _b of the Sim3 vertex (differential probe on a throw-away quadratic form)
        Hs = P._aligned(52)
        sg.f["v_map"](so.vs.ctypes.data, Hs.ctypes.data)
        sg.f["v_clear"](so.vs.ctypes.data)
        snap = so.vs.copy()
        so.error(pairs[0][0])
        so.f["lin"](pairs[0][0]["e"].ctypes.data, pairs[0][0]["jw"].ctypes.data)
        so.f["quad"](pairs[0][0]["e"].ctypes.data)
        ch = np.flatnonzero(so.vs.view(np.uint64) != snap.view(np.uint64))
        assert len(ch) >= 3 and (ch < ch.min() + 7).sum() >= 3 and ch.min() > V_DIM // 8, ch
        so.b_off = int(ch.min())
        sg.f["v_clear"](so.vs.ctypes.data)
        e12_0 = np.stack([so.error(a) for a, _ in pairs])
        e21_0 = np.stack([so.error(b) for _, b in pairs])
        J12_0 = np.stack([so.jacobian(a) for a, _ in pairs])
        J21_0 = np.stack([so.jacobian(b) for _, b in pairs])
        L = G.ed.g.L
        lm_ctor = L._ZN3g2o30OptimizationAlgorithmLevenbergC1EPNS_6SolverE
        lm_ctor.restype, lm_ctor.argtypes = None, [C.c_void_p, C.c_void_p]
        set_alg = L._ZN3g2o15SparseOptimizer12setAlgorithmEPNS_21OptimizationAlgorithmE
        set_alg.restype, set_alg.argtypes = None, [C.c_void_p, C.c_void_p]
        optimize = L._ZN3g2o15SparseOptimizer8optimizeEib
        optimize.restype, optimize.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_bool]
        solver = FakeSolverSim3Dense(so)
        alg = P._aligned(1024)
        lm_ctor(alg.ctypes.data, solver.obj.ctypes.data)
        set_alg(o, alg.ctypes.data)
        assert G.f["init"](o, 0)
        it1 = optimize(o, 5, False)
        n_l1 = sum(1 for k, _ in solver.log if k == "lambda")
        keep = np.ones(n, np.uint8)
        nBad = 0
        for k, (a, b) in enumerate(pairs):
            if so.chi2(a) > th2 or so.chi2(b) > th2:                      # stored _error, float th2 promoted
                keep[k] = 0
                so.set_level(a, 1)
                so.set_level(b, 1)
                nBad += 1
        more = 10 if nBad > 0 else 5
        assert n - nBad >= 10
        assert G.f["init"](o, 0)
        it2 = optimize(o, more, False)
        nIn = 0
        for k, (a, b) in enumerate(pairs):
            if not keep[k]:
                continue
            if so.chi2(a) > th2 or so.chi2(b) > th2:
                keep[k] = 0
            else:
                nIn += 1
        lam = np.array([v for k, v in solver.log if k == "lambda"])
        pre = f"s3o{case}_"
        out.update({pre + "s0": s0, pre + "cam8": cam8, pre + "p1c": p1, pre + "p2c": p2, pre + "meas6": meas,
                    pre + "fix_scale": np.array(int(fix_scale)), pre + "th2": np.array(th2), pre + "e12_0": e12_0,
                    pre + "e21_0": e21_0, pre + "J12_0": J12_0, pre + "J21_0": J21_0, pre + "lambda": lam,
                    pre + "n_trials_pass0": np.array(n_l1), pre + "iterations": np.array([it1, it2]),
                    pre + "keep": keep, pre + "nIn": np.array(nIn), pre + "nBad": np.array(nBad),
                    pre + "s12": sg.estimate(0)})
        keep_alive.append((so, solver, alg, Hs))
        print(f"OptimizeSim3 (fix_scale={fix_scale}): {it1}+{it2} iterations, {len(lam)} trials, nBad {nBad}, nIn {nIn} of {n}")
    np.savez(path, **out)
    return out, keep_alive


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "tests", "golden", "libg2o_vectors.npz")
    o = make(p)
    make_lm(p)
    _keep = make_lba(p)
    _keep2 = make_poseopt(p)
    make_sim3(p)
    _keep3 = make_posegraph(p)
    _keep4 = make_sim3opt(p)
    print("wrote", p, "| phase A index:", o["graph_A_pose_index"], o["graph_A_point_index"], "| phase B index:",
          o["graph_B_pose_index"], o["graph_B_point_index"], "| chi2", o["graph_A_chi2"], o["graph_A_robust_chi2"],
          o["graph_B_chi2"])
    sys.stdout.flush()
    os._exit(0)
