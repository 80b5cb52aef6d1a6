"""NumPy model of the square-root (QR / null-space) LM step that the CUDA kernels implement.

Test infrastructure: proves on the CPU that the sqrt formulation -- per-landmark Householder QR of
[sqrt(lambda) I3 ; J_l], projector I - Q1 Q1^T, PCG with 6x6 block-Jacobi on the implicit reduced system,
back-substitution through R -- reproduces g2o's damped Schur step (block_solver.hpp:354-486) from the
oracle.  Mirrors the data flow of csrc/ (same quantities, same names) so kernel bugs can be bisected."""
import numpy as np


def householder_qr_damped(Jl, lam):
    """QR of A = [sqrt(lam) I3 ; Jl] with the damping rows first.  Returns R (3x3 upper) and the obs rows of Q1."""
    m = Jl.shape[0]
    a = Jl.copy()
    R = np.zeros((3, 3))
    V = np.zeros((m, 3))
    v0 = np.zeros(3)
    beta = np.zeros(3)
    sl = np.sqrt(lam)
    for j in range(3):
        sigma = a[:, j] @ a[:, j]
        norm = np.sqrt(lam + sigma)
        v0[j] = sl + norm
        R[j, j] = -norm
        beta[j] = 1.0 / (norm * v0[j])
        V[:, j] = a[:, j]
        for c in range(j + 1, 3):
            w = beta[j] * (a[:, j] @ a[:, c])      # damping row j of column c is zero
            R[j, c] = -w * v0[j]
            a[:, c] -= w * a[:, j]
    # compact WY: Q = I - V T V^T, T upper triangular (larft, forward/columnwise)
    G = V.T @ V
    T = np.zeros((3, 3))
    for j in range(3):
        T[j, j] = beta[j]
        if j:
            # full reflector j has v0[j] on damping row j only, so V_full^T v_full = V^T v (obs part) for i<j
            T[:j, j] = -beta[j] * T[:j, :j] @ G[:j, j]
    M = T @ np.diag(v0)                            # Q1 obs rows = -V M^T? derive: Q[:,0:3] = E - V_full T V_full^T E
    Q1 = -V @ (T @ np.diag(v0))                    # V_full^T E = diag(v0) (damping rows of V_full)
    return R, Q1


def sqrt_step(prob, lin, lam, pose_slot, pcg_tol=1e-12, max_iter=500):
    """One damped LM step in square-root form. lin = oracle.linearize_all() dict (unweighted J, err, w)."""
    Np = int(pose_slot.max()) + 1
    lm_ptr = prob.lm_ptr()
    sw = np.sqrt(lin["w"])[:, None]
    stereo = ~(prob.obs_meas[:, 2] < 0)
    d = np.where(stereo, 3, 2)
    Jp = lin["Jp"] * sw[:, :, None]
    Jl = lin["Jl"] * sw[:, :, None]
    r = lin["err"] * sw
    for k in range(prob.n_obs):
        Jp[k, d[k]:] = 0; Jl[k, d[k]:] = 0; r[k, d[k]:] = 0
    slot = pose_slot[prob.obs_pose]
    Q1 = np.zeros_like(Jl)
    Rl = np.zeros((prob.n_point, 3, 3))
    tl = np.zeros((prob.n_point, 3))
    for l in range(prob.n_point):
        a, b = lm_ptr[l], lm_ptr[l + 1]
        A = Jl[a:b].reshape(-1, 3)
        R, Q = householder_qr_damped(A, lam)
        Rl[l] = R
        Q1[a:b] = Q.reshape(-1, 3, 3)
        tl[l] = np.einsum("kri,kr->i", Q1[a:b], r[a:b])          # Q1^T c  (damping rows of c are zero)

    def project_scatter(v):
        """y = sum_o Jp_o^T (v_o - Q1_o s_l), s_l = sum_o Q1_o^T v_o"""
        y = np.zeros((Np, 6))
        for l in range(prob.n_point):
            a, b = lm_ptr[l], lm_ptr[l + 1]
            s = np.einsum("kri,kr->i", Q1[a:b], v[a:b])
            u = v[a:b] - np.einsum("kri,i->kr", Q1[a:b], s)
            for k in range(a, b):
                if slot[k] >= 0:
                    y[slot[k]] += Jp[k].T @ u[k - a]
        return y

    def matvec(p):
        v = np.zeros((prob.n_obs, 3))
        m = slot >= 0
        v[m] = np.einsum("kri,ki->kr", Jp[m], p[slot[m]])
        return project_scatter(v) + lam * p

    bs = -project_scatter(r)
    # block-Jacobi preconditioner: diagonal blocks of the reduced matrix
    D = np.tile(lam * np.eye(6), (Np, 1, 1))
    for k in range(prob.n_obs):
        if slot[k] >= 0:
            G = Jp[k].T @ Q1[k]
            D[slot[k]] += Jp[k].T @ Jp[k] - G @ G.T
    Dinv = np.linalg.inv(D)
    x = np.zeros((Np, 6)); res = bs.copy(); z = np.einsum("nij,nj->ni", Dinv, res); p = z.copy()
    rz = (res * z).sum(); rz0 = rz; its = 0
    while its < max_iter and rz > pcg_tol ** 2 * rz0:
        q = matvec(p); alpha = rz / (p * q).sum()
        x += alpha * p; res -= alpha * q
        z = np.einsum("nij,nj->ni", Dinv, res); rz_new = (res * z).sum()
        p = z + (rz_new / rz) * p; rz = rz_new; its += 1
    # back-substitution: dl = -R^{-1} (t_l + sum_o Q1_o^T Jp_o dp)
    dl = np.zeros((prob.n_point, 3))
    for l in range(prob.n_point):
        a, b = lm_ptr[l], lm_ptr[l + 1]
        g = tl[l].copy()
        for k in range(a, b):
            if slot[k] >= 0:
                g += Q1[k].T @ (Jp[k] @ x[slot[k]])
        dl[l] = -np.linalg.solve(Rl[l], g)
    # gradient pieces for the gain ratio
    bp = np.zeros((Np, 6)); bl = np.zeros((prob.n_point, 3))
    for k in range(prob.n_obs):
        if slot[k] >= 0:
            bp[slot[k]] -= Jp[k].T @ r[k]
        bl[prob.obs_point[k]] -= Jl[k].T @ r[k]
    return dict(dp=x, dl=dl, bs=bs, bp=bp, bl=bl, iters=its, matvec=matvec, Dinv=Dinv)
