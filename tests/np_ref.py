"""Independent NumPy/SciPy restatement of the reference BA (second opinion on oracle/refba.cpp).

Deliberately written differently from the C++ oracle: rotation matrices + scipy.linalg.expm instead of
quaternions, ONE dense (H + lambda I) x = b solve over all variables instead of Schur + back-substitution.
Agreement of the two (tests/test_oracle.py) is the mitigation for "parity unpinned upstream" (SURVEY.md §8(c)).
Follows the same reference lines: types_six_dof_expmap.cpp:103-157,188-234 (edges),
robust_kernel_impl.cpp:78-91 (Huber), optimization_algorithm_levenberg.cpp:61-189 (LM),
g2oOptimizer.cc:704-976,1119-1142 (two-pass local BA).  Small problems only (dense O(n^3)).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm, solve


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def hat6(xi):
    w, v = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


class NpBA:
    def __init__(self, prob):
        self.p = prob
        self.T = np.zeros((prob.n_pose, 4, 4))
        for i in range(prob.n_pose):
            q = prob.pose_qt[i, 3:] / np.linalg.norm(prob.pose_qt[i, 3:])
            self.T[i] = np.eye(4)
            self.T[i, :3, :3] = quat_to_R(q)
            self.T[i, :3, 3] = prob.pose_qt[i, :3]
        self.X = prob.point_xyz.copy()
        self.stereo = ~(prob.obs_meas[:, 2] < 0)
        self.level = np.zeros(prob.n_obs, int)
        self.robust = True
        self.delta = np.where(self.stereo, float(np.float32(np.sqrt(7.815))), float(np.float32(np.sqrt(5.991))))
        self.err = np.zeros((prob.n_obs, 3))  # stored edge errors (g2o _error semantics)
        self.trace = []

    # residual of one observation at the current state
    def _resid(self, k):
        p = self.p
        T = self.T[p.obs_pose[k]]
        Xc = T[:3, :3] @ self.X[p.obs_point[k]] + T[:3, 3]
        fx, fy, cx, cy, bf = p.cam[p.obs_pose[k]]
        m = p.obs_meas[k].astype(np.float64)
        if self.stereo[k]:
            invz = np.float32(1.0 / Xc[2])  # float32 inverse depth (types_six_dof_expmap.cpp:151)
            u = Xc[0] * float(invz) * fx + cx
            v = Xc[1] * float(invz) * fy + cy
            ur = u - float(np.float32(bf) * invz)
            return np.array([m[0] - u, m[1] - v, m[2] - ur]), Xc
        return np.array([m[0] - (Xc[0] / Xc[2] * fx + cx), m[1] - (Xc[1] / Xc[2] * fy + cy), 0.0]), Xc

    def compute_errors(self, active):
        for k in active:
            self.err[k], _ = self._resid(k)

    def chi2(self, k):
        return float(self.p.obs_meas[k, 3]) * float(self.err[k] @ self.err[k])

    def rho(self, k):
        c = self.chi2(k)
        d = self.delta[k]
        dsqr = float(np.float32(d * d))   # RobustKernelHuber keeps delta^2 in a float member (robust_kernel_impl.h:84)
        if not self.robust or c <= dsqr:
            return c, 1.0
        s = np.sqrt(c)
        return 2 * s * d - dsqr, d / s

    def robust_chi2(self, active):
        return sum(self.rho(k)[0] for k in active)

    def jac(self, k):
        p = self.p
        T = self.T[p.obs_pose[k]]
        R = T[:3, :3]
        x, y, z = R @ self.X[p.obs_point[k]] + T[:3, 3]
        fx, fy, cx, cy, bf = p.cam[p.obs_pose[k]]
        # d(proj)/d(Xc)
        dp = np.array([[fx / z, 0, -fx * x / z ** 2], [0, fy / z, -fy * y / z ** 2], [0, 0, 0]])
        if self.stereo[k]:
            dp[2] = dp[0] + np.array([0, 0, bf / z ** 2])
        # d(Xc)/d(xi) for T <- exp(xi) T, xi = (omega, upsilon): [-[Xc]x, I]
        dX = np.zeros((3, 6))
        dX[:, :3] = -np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        dX[:, 3:] = np.eye(3)
        return -dp @ dX, -dp @ R  # Jp (3x6), Jl (3x3)

    def optimize(self, iters, pass_id):
        p = self.p
        active = [k for k in range(p.n_obs) if self.level[k] == 0]
        pose_act = sorted({int(p.obs_pose[k]) for k in active if not p.pose_fixed[p.obs_pose[k]]})
        pt_act = sorted({int(p.obs_point[k]) for k in active})
        ps = {i: s for s, i in enumerate(pose_act)}
        ls = {i: s for s, i in enumerate(pt_act)}
        npd, n = 6 * len(pose_act), 6 * len(pose_act) + 3 * len(pt_act)
        lam, ni, nbad = None, 2.0, 0
        for it in range(iters):
            self.compute_errors(active)
            cur = self.robust_chi2(active)
            ini = cur
            H = np.zeros((n, n))
            b = np.zeros(n)
            for k in active:
                Jp, Jl = self.jac(k)
                w = self.rho(k)[1] * float(p.obs_meas[k, 3])
                e = self.err[k]
                cols = []
                J = np.zeros((3, 0))
                if int(p.obs_pose[k]) in ps:
                    s = ps[int(p.obs_pose[k])]
                    cols += list(range(6 * s, 6 * s + 6))
                    J = np.hstack([J, Jp])
                s = ls[int(p.obs_point[k])]
                cols += list(range(npd + 3 * s, npd + 3 * s + 3))
                J = np.hstack([J, Jl])
                H[np.ix_(cols, cols)] += w * J.T @ J
                b[cols] += -w * J.T @ e
            if it == 0:
                lam, ni, nbad = 1e-5 * np.abs(np.diag(H)).max(), 2.0, 0
            q, rho = 0, 0.0
            while True:
                bakT, bakX = self.T.copy(), self.X.copy()
                x = solve(H + lam * np.eye(n), b, assume_a="pos")
                for i, s in ps.items():
                    self.T[i] = expm(hat6(x[6 * s:6 * s + 6])) @ self.T[i]
                for i, s in ls.items():
                    self.X[i] += x[npd + 3 * s:npd + 3 * s + 3]
                self.compute_errors(active)
                tmp = self.robust_chi2(active)
                rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
                acc = rho > 0 and np.isfinite(tmp)
                self.trace.append((pass_id, it, q, lam, cur, tmp, rho, float(acc)))
                if acc:
                    lam *= max(1. / 3., min(1. - (2 * rho - 1) ** 3, 2. / 3.))
                    ni = 2.0
                    cur = tmp
                else:
                    lam *= ni
                    ni *= 2
                    self.T, self.X = bakT, bakX
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break

    def depth_positive(self, k):
        return self._resid(k)[1][2] > 0.0

    def solve_local(self):
        p = self.p
        self.robust = True
        self.optimize(5, 0)
        for k in range(p.n_obs):
            thr = 7.815 if self.stereo[k] else 5.991
            if self.chi2(k) > thr or not self.depth_positive(k):
                self.level[k] = 1
        self.robust = False
        self.optimize(10, 1)
        out = np.zeros(p.n_obs, np.uint8)
        for k in range(p.n_obs):
            thr = 7.815 if self.stereo[k] else 5.991
            out[k] = self.chi2(k) > thr or not self.depth_positive(k)
        return out
