"""Secondary oracle: the reference's OWN Levenberg-Marquardt back-end ("mybackend"),
restated in NumPy.  TEST INFRASTRUCTURE ONLY -- nothing under sqrtlm-slam_b200/ or include/ may import this file.

The reference ships two optimisers behind `Optimizer::LocalBundleAdjustment`: vendored g2o (the default,
src/backend/Optimizer.cc:26) and the author's hand-written `myslam::backend::Problem` (Optimizer.cc:27, disabled).  Both
solve the same damped Schur system with the same Nielsen lambda policy, but they were written independently and follow
different conventions, which is exactly what makes the second one useful as a cross-check of the first oracle
(oracle/refba.cpp): the two restatements follow different reference files and must agree wherever the mathematics says
they have to (tests/test_mybackend_oracle.py).

What is restated, with the lines it follows (relative to /root/reference):
  * pose update  T <- exp([upsilon, omega]) * T, TRANSLATION FIRST (Sophus::SE3d)      src/backend/mybackend/vertex_pose.cc:7-14
  * mono reprojection edge: residual obs - proj, J_point (2x3), J_pose (2x6, translation columns first)
                                                                                       src/backend/mybackend/edge_reprojection.cc:33-130
  * chi2 = r^T Omega r, robust chi2, robust information with Triggs' second-order term when rho' + 2 rho'' chi2 > 0
                                                                                       src/backend/mybackend/edge.cc:35-77
  * Huber and Cauchy losses (rho, rho', rho'')                                         src/backend/mybackend/loss_function.cc:9-31
  * MakeHessian (DENSE_MODE blocks, only edges of the optimised level, b -= rho' J^T Omega r)
                                                                                       src/backend/mybackend/problem.cc:330-423
  * lambda0 = 1e-5 * min(5e10, max diag), chi = 1/2 sum of ROBUST chi2 over ALL edges (every level)
                                                                                       src/backend/mybackend/problem.cc:591-630
  * Schur complement with lambda on BOTH diagonals, LDLT, back-substitution            src/backend/mybackend/problem.cc:430-557, 632-676
  * gain ratio with scale 1/2 d^T (lambda d + b) + 1e-6, lambda policy, <= 10 retries, stop when the cost decreased by
    less than 1e-5 over an outer iteration                                             src/backend/mybackend/problem.cc:92-167, 679-713
  * the two-pass local BA around it: Cauchy(sqrt 5.991), Solve(5), chi2 > 5.991 or negative depth -> level 1, loss off,
    Solve(10), final classification                                                    src/backend/myOptimizer.cc:395-403, 440-442, 464-533

Parity status: UNPINNED upstream -- mybackend needs Eigen + Sophus, neither is in this image, and the reference has no
test or golden vector for it.  It is pinned against the PRIMARY oracle instead (which is pinned against the reference's
g2o binary): same Jacobians up to the tangent permutation, same first LM step, same minimiser.

Dense O((6 Np)^3 + Np^2 Nl) arithmetic: small windows only.
"""
from __future__ import annotations

import numpy as np


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def se3_exp_translation_first(delta):
    """Sophus::SE3d::exp: delta = [upsilon (3), omega (3)] -> (R, t) with t = V upsilon."""
    ups, om = delta[:3], delta[3:]
    th2 = float(om @ om)
    th = np.sqrt(th2)
    Om = hat(om)
    if th < 1e-10:  # Sophus switches to the Taylor series near zero
        R = np.eye(3) + Om + 0.5 * Om @ Om
        V = np.eye(3) + 0.5 * Om + Om @ Om / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th2 * Om + (th - np.sin(th)) / (th2 * th) * Om @ Om
    return R, V @ ups


class Loss:
    """rho(e2), rho', rho'' (loss_function.cc)."""

    def __init__(self, kind, delta):
        self.kind, self.delta = kind, float(delta)

    def compute(self, e2):
        e2 = np.asarray(e2, float)
        d2 = self.delta * self.delta
        if self.kind == "cauchy":
            aux = e2 / d2 + 1.0
            r1 = 1.0 / aux
            return d2 * np.log(aux), r1, -(1.0 / d2) * r1 ** 2
        if self.kind == "huber":
            inl = e2 <= d2
            s = np.sqrt(np.where(inl, 1.0, e2))
            r1 = np.where(inl, 1.0, self.delta / s)
            return (np.where(inl, e2, 2 * s * self.delta - d2), r1,
                    np.where(inl, 0.0, -0.5 * r1 / np.where(inl, 1.0, e2)))
        raise ValueError(self.kind)


class MyBackendBA:
    """`myslam::backend::Problem` over a synth.Problem with monocular edges (EdgeReprojectionXYZ)."""

    def __init__(self, prob):
        self.p = prob
        n = prob.n_pose
        self.R = np.zeros((n, 3, 3))
        self.t = prob.pose_qt[:, :3].copy()
        for i in range(n):
            q = prob.pose_qt[i, 3:]
            self.R[i] = quat_to_R(q / np.linalg.norm(q))
        self.X = prob.point_xyz.copy()
        self.free = np.where(prob.pose_fixed == 0)[0]
        self.slot = -np.ones(n, int)
        self.slot[self.free] = np.arange(len(self.free))
        self.level = np.zeros(prob.n_obs, int)
        self.loss = None                     # one loss object shared by the edges that still carry it
        self.has_loss = np.zeros(prob.n_obs, bool)
        self.info = prob.obs_meas[:, 3].astype(np.float64)   # Omega = invSigma2 * I
        self.obs = prob.obs_meas[:, :2].astype(np.float64)
        self.res = np.zeros((prob.n_obs, 2))  # Edge::residual_: whatever ComputeResidual left there
        self.trace = []                       # (outer iteration, retry, lambda, chi before, chi trial, rho, accepted)
        self.compute_residuals(np.arange(prob.n_obs))

    # ---- edges
    def cam_points(self):
        p = self.p
        return np.einsum("kij,kj->ki", self.R[p.obs_pose], self.X[p.obs_point]) + self.t[p.obs_pose]

    def compute_residuals(self, idx):
        p = self.p
        Xc = self.cam_points()[idx]
        fx, fy, cx, cy = (p.cam[p.obs_pose[idx], c] for c in range(4))
        proj = np.stack([Xc[:, 0] / Xc[:, 2] * fx + cx, Xc[:, 1] / Xc[:, 2] * fy + cy], 1)
        self.res[idx] = self.obs[idx] - proj

    def chi2(self):
        return self.info * np.einsum("ki,ki->k", self.res, self.res)

    def robust_chi2(self):
        c = self.chi2()
        if self.loss is None:
            return c
        return np.where(self.has_loss, self.loss.compute(c)[0], c)

    def depth_positive(self):
        return self.cam_points()[:, 2] > 0.0

    def jacobians(self, idx):
        """(J_point (m,2,3), J_pose (m,2,6)) -- pose columns: translation 0-2, rotation 3-5 (edge_reprojection.cc:98-111)."""
        p = self.p
        Xc = self.cam_points()[idx]
        x, y, iz = Xc[:, 0], Xc[:, 1], 1.0 / Xc[:, 2]
        iz2 = iz * iz
        fx, fy = p.cam[p.obs_pose[idx], 0], p.cam[p.obs_pose[idx], 1]
        tmp = np.zeros((len(idx), 2, 3))
        tmp[:, 0, 0] = fx
        tmp[:, 0, 2] = -x * iz * fx
        tmp[:, 1, 1] = fy
        tmp[:, 1, 2] = -y * iz * fy
        Jl = -iz[:, None, None] * np.einsum("kij,kjl->kil", tmp, self.R[p.obs_pose[idx]])
        Jp = np.zeros((len(idx), 2, 6))
        Jp[:, 0, 0] = -iz * fx
        Jp[:, 0, 2] = x * iz2 * fx
        Jp[:, 0, 3] = x * y * iz2 * fx
        Jp[:, 0, 4] = -(1 + x * x * iz2) * fx
        Jp[:, 0, 5] = y * iz * fx
        Jp[:, 1, 1] = -iz * fy
        Jp[:, 1, 2] = y * iz2 * fy
        Jp[:, 1, 3] = (1 + y * y * iz2) * fy
        Jp[:, 1, 4] = -x * y * iz2 * fy
        Jp[:, 1, 5] = -x * iz * fy
        return Jl, Jp

    def robust_info(self, idx):
        """drho (m), robust information (m,2,2) -- Edge::RobustInfo."""
        m = len(idx)
        info = self.info[idx]
        W = info[:, None, None] * np.eye(2)[None]
        drho = np.ones(m)
        if self.loss is not None:
            hl = self.has_loss[idx]
            e2 = info * np.einsum("ki,ki->k", self.res[idx], self.res[idx])
            _, r1, r2 = self.loss.compute(e2)
            we = info[:, None] * self.res[idx]
            Wr = r1[:, None, None] * W
            trig = (r1 + 2 * r2 * e2) > 0.0
            Wr = Wr + np.where(trig, 2 * r2, 0.0)[:, None, None] * np.einsum("ki,kj->kij", we, we)
            W = np.where(hl[:, None, None], Wr, W)
            drho = np.where(hl, r1, 1.0)
        return drho, W

    # ---- Problem::MakeHessian (DENSE_MODE)
    def make_hessian(self, level=0):
        p = self.p
        act = np.where(self.level == level)[0]
        self.compute_residuals(act)
        Jl, Jp = self.jacobians(act)
        drho, W = self.robust_info(act)
        Np, Nl = len(self.free), p.n_point
        Hpp = np.zeros((6 * Np, 6 * Np))
        Hpl = np.zeros((6 * Np, 3 * Nl))
        Hll = np.zeros((Nl, 3, 3))
        b = np.zeros(6 * Np + 3 * Nl)
        s = self.slot[p.obs_pose[act]]
        l = p.obs_point[act]
        np.add.at(Hll, l, np.einsum("kji,kjm,kmn->kin", Jl, W, Jl))
        we = self.info[act][:, None] * self.res[act]
        gl = -drho[:, None] * np.einsum("kji,kj->ki", Jl, we)
        np.add.at(b, (6 * Np + 3 * l[:, None] + np.arange(3)[None]), gl)
        fr = s >= 0
        Hp = np.einsum("kji,kjm,kmn->kin", Jp[fr], W[fr], Jp[fr])
        Hx = np.einsum("kji,kjm,kmn->kin", Jp[fr], W[fr], Jl[fr])  # pose x landmark
        gp = -drho[fr][:, None] * np.einsum("kji,kj->ki", Jp[fr], we[fr])
        for k, (sk, lk) in enumerate(zip(s[fr], l[fr])):
            Hpp[6 * sk:6 * sk + 6, 6 * sk:6 * sk + 6] += Hp[k]
            Hpl[6 * sk:6 * sk + 6, 3 * lk:3 * lk + 3] += Hx[k]
            b[6 * sk:6 * sk + 6] += gp[k]
        self.Hpp, self.Hpl, self.Hll, self.b = Hpp, Hpl, Hll, b

    def lambda_init(self):
        self.ni = 2.0
        self.chi = 0.5 * float(self.robust_chi2().sum())
        md = max(float(np.abs(np.diag(self.Hpp)).max(initial=0.0)),
                 float(np.abs(self.Hll[:, [0, 1, 2], [0, 1, 2]]).max(initial=0.0)))
        self.lam = 1e-5 * min(5e10, md)

    def solve_linear(self):
        Np, Nl = len(self.free), self.p.n_point
        lam = self.lam
        Hpp = self.Hpp + lam * np.eye(6 * Np)
        Hmm_inv = np.linalg.inv(self.Hll + lam * np.eye(3)[None])
        bp, bm = self.b[:6 * Np], self.b[6 * Np:]
        tempH = np.einsum("plk,lkm->plm", self.Hpl.reshape(6 * Np, Nl, 3), Hmm_inv).reshape(6 * Np, 3 * Nl)
        S = Hpp - tempH @ self.Hpl.T
        bs = bp - tempH @ bm
        dp = np.linalg.solve(S, bs)
        ind = (bm - self.Hpl.T @ dp).reshape(Nl, 3)
        dl = np.einsum("lij,lj->li", Hmm_inv, ind).ravel()
        self.delta = np.concatenate([dp, dl])
        self.S, self.bs = S, bs

    def update_states(self):
        self._bak = (self.R.copy(), self.t.copy(), self.X.copy())
        Np = len(self.free)
        for k, i in enumerate(self.free):
            Rd, td = se3_exp_translation_first(self.delta[6 * k:6 * k + 6])
            self.R[i], self.t[i] = Rd @ self.R[i], Rd @ self.t[i] + td
        self.X = self.X + self.delta[6 * Np:].reshape(-1, 3)

    def rollback(self):
        self.R, self.t, self.X = self._bak

    def is_good_step(self):
        scale = 0.5 * float(self.delta @ (self.lam * self.delta + self.b)) + 1e-6
        self.compute_residuals(np.arange(self.p.n_obs))     # every edge, whatever its level
        temp = 0.5 * float(self.robust_chi2().sum())
        rho = (self.chi - temp) / scale
        if rho > 0 and np.isfinite(temp):
            alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
            self.lam *= max(1.0 / 3.0, alpha)
            self.ni = 2.0
            out = (self.chi, temp, rho, True)
            self.chi = temp
            return out
        out = (self.chi, temp, rho, False)
        self.lam *= self.ni
        self.ni *= 2
        return out

    def solve(self, iterations, level=0):
        """Problem::Solve."""
        self.make_hessian(level)
        self.lambda_init()
        stop, it, last = False, 0, 1e20
        while not stop and it < iterations:
            ok, fails = False, 0
            while not ok and fails < 10:
                lam = self.lam
                self.solve_linear()
                self.update_states()
                chi0, chi1, rho, ok = self.is_good_step()
                self.trace.append((it, fails, lam, chi0, chi1, rho, float(ok)))
                if ok:
                    self.make_hessian(level)
                    fails = 0
                else:
                    fails += 1
                    self.rollback()
            it += 1
            if last - self.chi < 1e-5:
                stop = True
            last = self.chi
        return it

    # ---- myOptimizer::LocalBundleAdjustment around it
    def local_ba(self):
        self.loss = Loss("cauchy", np.sqrt(5.991))
        self.has_loss[:] = True
        self.solve(5)
        bad = (self.chi2() > 5.991) | ~self.depth_positive()
        self.level[bad] = 1
        self.has_loss[:] = False
        self.loss = None
        self.solve(10)
        return (self.chi2() > 5.991) | ~self.depth_positive()

    def poses_qt(self):
        """(n_pose, 7) t, q(x, y, z, w) -- for comparisons with the other oracles."""
        from scipy.spatial.transform import Rotation
        q = Rotation.from_matrix(self.R).as_quat()
        q = np.where(q[:, 3:4] < 0, -q, q)
        return np.concatenate([self.t, q], 1)
